// N-dimensional grid on one flat vector, x fastest.
// Interface mirrored: reference include/emcGrid.hpp (layout :242-254, iteration
// helpers :122-138, neighbour access :76-92, box fill :102-118, bounds errors
// :266-272).  This layout is the contract of every grid that crosses the C ABI.
#ifndef EMC_GRID_HPP
#define EMC_GRID_HPP

#include <algorithm>
#include <array>
#include <numeric>
#include <string>
#include <vector>

#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <typename T, SizeType Dim> class emcGrid {
public:
  typedef std::array<SizeType, Dim> CoordVec;

private:
  std::vector<T> data;
  CoordVec extent;

  static SizeType volume(const CoordVec &e) {
    SizeType n = 1;
    for (auto x : e)
      n *= x;
    return n;
  }
  SizeType flat(const CoordVec &c) const {
    SizeType idx = 0;
    for (SizeType d = Dim; d-- > 0;)
      idx = idx * extent[d] + c[d];
    return idx;
  }
  void requireValid(const CoordVec &c) const {
    if (!isValid(c)) {
      std::string s;
      for (auto x : c)
        s += std::to_string(x) + " ";
      emcMessage::getInstance().addError("Grid coordinate { " + s + "} is out of bounds.").print();
    }
  }
  void requireDim(SizeType d) const {
    if (d >= Dim)
      emcMessage::getInstance().addError("Grid dimension index is out of bounds.").print();
  }

public:
  emcGrid() = delete;
  explicit emcGrid(const CoordVec &inExtent) : data(volume(inExtent)), extent(inExtent) {}
  emcGrid(const CoordVec &inExtent, T value) : data(volume(inExtent), value), extent(inExtent) {}
  emcGrid(const emcGrid &) = default;
  emcGrid &operator=(const emcGrid &other) {
    data = other.data;
    extent = other.extent;
    return *this;
  }

  SizeType getSize(SizeType idxDim) const {
    requireDim(idxDim);
    return extent[idxDim];
  }
  SizeType getSize() const { return data.size(); }
  CoordVec getExtent() const { return extent; }
  // raw x-fastest storage (what the C ABI uploads)
  const T *raw() const { return data.data(); }
  T *raw() { return data.data(); }

  T &operator[](const CoordVec &c) {
    requireValid(c);
    return data[flat(c)];
  }
  const T &operator[](const CoordVec &c) const {
    requireValid(c);
    return data[flat(c)];
  }
  const T &getNextValue(const CoordVec &c, int idxDim) const {
    CoordVec n = c;
    n[idxDim]++;
    if (!isValid(n))
      emcMessage::getInstance().addError("Grid neighbour does not exist.").print();
    return data[flat(n)];
  }
  const T &getPrevValue(const CoordVec &c, int idxDim) const {
    if (c[idxDim] == 0)
      emcMessage::getInstance().addError("Grid neighbour does not exist.").print();
    CoordVec n = c;
    n[idxDim]--;
    return data[flat(n)];
  }

  void fill(const T value) { std::fill(data.begin(), data.end(), value); }
  // closed box [minCoord, maxCoord], clipped to the grid
  void fill(const T value, const CoordVec &minCoord, const CoordVec &maxCoord) {
    CoordVec c;
    for (c.fill(0); !isEndCoord(c); advanceCoord(c)) {
      bool inside = true;
      for (SizeType d = 0; d < Dim; d++)
        inside = inside && c[d] >= minCoord[d] && c[d] <= maxCoord[d];
      if (inside)
        data[flat(c)] = value;
    }
  }
  void iota(const T value) { std::iota(data.begin(), data.end(), value); }

  // iterate coordinates in storage order: for (c.fill(0); !isEndCoord(c); advanceCoord(c))
  void advanceCoord(CoordVec &c) const {
    c[0]++;
    for (SizeType d = 0; d + 1 < Dim; d++) {
      if (c[d] >= extent[d]) {
        c[d + 1]++;
        for (SizeType r = 0; r <= d; r++)
          c[r] = 0;
      }
    }
  }
  bool isEndCoord(const CoordVec &c) const { return c[Dim - 1] == extent[Dim - 1]; }

  typename std::vector<T>::const_iterator begin() const { return data.begin(); }
  typename std::vector<T>::const_iterator end() const { return data.end(); }
  typename std::vector<T>::iterator begin() { return data.begin(); }
  typename std::vector<T>::iterator end() { return data.end(); }

  bool onBoundary(const CoordVec &c) const {
    for (SizeType d = 0; d < Dim; d++)
      if (c[d] == 0 || c[d] == extent[d] - 1)
        return true;
    return false;
  }
  bool hasSameExtent(const emcGrid &o) const { return extent == o.extent; }
  bool isValid(const CoordVec &c) const {
    for (SizeType d = 0; d < Dim; d++)
      if (c[d] >= extent[d])
        return false;
    return true;
  }

  emcGrid &operator+=(const emcGrid &rhs) {
    if (!hasSameExtent(rhs))
      emcMessage::getInstance().addError("Grids of different extent cannot be added.").print();
    for (SizeType i = 0; i < data.size(); i++)
      data[i] += rhs.data[i];
    return *this;
  }
  emcGrid &operator-=(const emcGrid &rhs) {
    if (!hasSameExtent(rhs))
      emcMessage::getInstance().addError("Grids of different extent cannot be subtracted.").print();
    for (SizeType i = 0; i < data.size(); i++)
      data[i] -= rhs.data[i];
    return *this;
  }
  // extent line, then rows of x-values, a blank line between z-planes (the layout of the reference's emcGrid::print,
  // include/emcGrid.hpp:143-165, which helper/emcPlottingFiles/readResultFile.py:readGridFile reads)
  void print(std::ostream &out = std::cout) const {
    out << extent << "\n";
    CoordVec c;
    for (c.fill(0); !isEndCoord(c); advanceCoord(c)) {
      if (c[0] == 0) {
        bool first = true;
        for (SizeType d = 1; d < Dim; d++)
          first = first && c[d] == 0;
        if (!first) {
          out << "\n";
          if (Dim > 2 && c[1] == 0)
            out << "\n";
        }
      }
      out << data[flat(c)];
      if (c[0] + 1 != extent[0])
        out << " ";
    }
    out << "\n";
  }
};

template <typename T, SizeType Dim> emcGrid<T, Dim> operator+(emcGrid<T, Dim> lhs, const emcGrid<T, Dim> &rhs) {
  lhs += rhs;
  return lhs;
}
template <typename T, SizeType Dim> emcGrid<T, Dim> operator-(emcGrid<T, Dim> lhs, const emcGrid<T, Dim> &rhs) {
  lhs -= rhs;
  return lhs;
}

#endif
