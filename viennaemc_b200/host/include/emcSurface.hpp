// The 2*Dim faces of the box and the contacts on them: per face one grid of contact indices
// (-1 = artificial boundary).
// Interface mirrored: reference include/emcSurface.hpp -- contact queries :58-118, addContact :123-146
// (a reservoir contact -- ohmic / Schottky -- also takes over its cells on the other faces they lie on,
// :351-381), cell classification :148-178, coordinate mapping :302-338, getBoundaryPos :340-349 (first
// face in the order XMIN XMAX YMIN YMAX ZMIN ZMAX), face iteration :383-405.
#ifndef EMC_SURFACE_HPP
#define EMC_SURFACE_HPP

#include <memory>
#include <vector>

#include <emcBoundaryPos.hpp>
#include <emcContact.hpp>
#include <emcGrid.hpp>
#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <class T, SizeType DimSurface> class emcSurface {
  static_assert(DimSurface == 1 || DimSurface == 2, "Wrong Dimension for Surface, possible dimensions are: {1,2}.");

public:
  static const SizeType DimDevice = DimSurface + 1;
  typedef std::array<SizeType, DimSurface> SizeVecSurface;
  typedef std::array<T, DimSurface> ValueVecSurface;
  typedef std::array<SizeType, DimDevice> SizeVecDevice;

private:
  SizeVecDevice deviceExtent;
  std::vector<std::unique_ptr<emcContact<T>>> contacts;
  std::vector<emcGrid<int, DimSurface>> contactOfFace; // [2*DimDevice]
  std::vector<emcBoundaryPos> faces;
  T voltageNormalization;

  void requireContact(SizeType idx) const {
    if (idx >= contacts.size())
      emcMessage::getInstance().addError("Index for Contact is out of bounds.").print();
  }
  void requireFace(emcBoundaryPos face) const {
    if (toUnderlying(face) >= 2 * DimDevice)
      emcMessage::getInstance().addError("Index for Boundary is out of bounds.").print();
  }

public:
  emcSurface() = delete;
  explicit emcSurface(const SizeVecDevice &deviceGridExtent, T inVoltageNormalizationParam = 1.)
      : deviceExtent(deviceGridExtent), voltageNormalization(inVoltageNormalizationParam) {
    for (SizeType f = 0; f < 2 * DimDevice; f++) {
      faces.push_back(static_cast<emcBoundaryPos>(f));
      contactOfFace.emplace_back(getCoordBoundary(deviceGridExtent, faces.back()), -1);
    }
  }

  SizeType getNrContacts() const { return contacts.size(); }
  const std::vector<emcBoundaryPos> &getAllBoundaryPos() const { return faces; }

  // ---- geometry -------------------------------------------------------------
  bool isOnBoundary(const SizeVecDevice &c, emcBoundaryPos face) const {
    const SizeType d = toUnderlying(face) / 2;
    return toUnderlying(face) % 2 ? c[d] == deviceExtent[d] - 1 : c[d] == 0;
  }
  bool isOnBoundary(const SizeVecDevice &c) const { return getBoundaryPos(c) != emcBoundaryPos::INVALID; }
  emcBoundaryPos getBoundaryPos(const SizeVecDevice &c) const {
    for (auto face : faces)
      if (isOnBoundary(c, face))
        return face;
    return emcBoundaryPos::INVALID;
  }
  // drop the coordinate normal to the face
  SizeVecSurface getCoordBoundary(const SizeVecDevice &c, emcBoundaryPos face) const {
    requireFace(face);
    const SizeType fixed = toUnderlying(face) / 2;
    SizeVecSurface out;
    for (SizeType d = 0, o = 0; d < DimDevice; d++)
      if (d != fixed)
        out[o++] = c[d];
    return out;
  }
  SizeVecDevice getCoordDevice(const SizeVecSurface &s, emcBoundaryPos face) const {
    requireFace(face);
    const SizeType fixed = toUnderlying(face) / 2;
    SizeVecDevice out;
    for (SizeType d = 0, i = 0; d < DimDevice; d++)
      out[d] = d == fixed ? (toUnderlying(face) % 2 ? deviceExtent[d] - 1 : 0) : s[i++];
    return out;
  }

  // ---- contacts -------------------------------------------------------------
  int getContactIdx(const SizeVecSurface &s, emcBoundaryPos face) const { return contactOfFace[toUnderlying(face)][s]; }
  int getContactIdx(const SizeVecDevice &c) const {
    const auto face = getBoundaryPos(c);
    return face == emcBoundaryPos::INVALID ? -1 : getContactIdx(getCoordBoundary(c, face), face);
  }
  int getOhmicContactIdx(const SizeVecDevice &c) const { return getContactIdx(c); }

  T getContactVoltage(SizeType idx, bool normalized = false) const {
    requireContact(idx);
    return normalized ? contacts[idx]->getAppliedVoltage() / voltageNormalization : contacts[idx]->getAppliedVoltage();
  }
  T getContactVoltage(const SizeVecSurface &s, emcBoundaryPos face, bool normalized = false) const {
    return getContactVoltage(getContactIdx(s, face), normalized);
  }
  emcContactType getContactType(SizeType idx) const {
    requireContact(idx);
    return contacts[idx]->getType();
  }
  emcContactType getContactType(const SizeVecSurface &s, emcBoundaryPos face) const {
    return getContactType(getContactIdx(s, face));
  }
  T getContactFurtherParameter(SizeType idx, SizeType idxInformation) const {
    requireContact(idx);
    return contacts[idx]->getFurtherParameter(idxInformation);
  }
  T getContactFurtherParameter(const SizeVecSurface &s, emcBoundaryPos face, const SizeType idxInformation) const {
    return getContactFurtherParameter(getContactIdx(s, face), idxInformation);
  }

  void addContact(emcBoundaryPos face, emcContactType type, T appliedVoltage, const SizeVecSurface &minCoord,
                  const SizeVecSurface &maxCoord, T epsOxide = 0, T thickness = 0, T barrierHeight = 0) {
    auto &grid = contactOfFace[toUnderlying(face)];
    if (!grid.isValid(minCoord))
      emcMessage::getInstance().addError("MinPos for Contact is out of bounds.").print();
    if (!grid.isValid(maxCoord))
      emcMessage::getInstance().addError("MaxCoord for Contact is out of bounds.").print();
    for (SizeType d = 0; d < DimSurface; d++)
      if (maxCoord[d] < minCoord[d])
        emcMessage::getInstance().addError("MaxPos of an added contact is smaller than its MinPos in one direction.").print();
    const int idx = static_cast<int>(contacts.size());
    grid.fill(idx, minCoord, maxCoord);
    if (type != emcContactType::GATE) {
      // a reservoir contact owns its cells on every face they touch (corners / edges)
      SizeVecSurface s;
      for (s.fill(0); !grid.isEndCoord(s); grid.advanceCoord(s)) {
        bool inside = true;
        for (SizeType d = 0; d < DimSurface; d++)
          inside = inside && s[d] >= minCoord[d] && s[d] <= maxCoord[d];
        if (!inside)
          continue;
        const auto cell = getCoordDevice(s, face);
        for (auto other : faces)
          if (isOnBoundary(cell, other))
            contactOfFace[toUnderlying(other)][getCoordBoundary(cell, other)] = idx;
      }
    }
    switch (type) {
    case emcContactType::GATE:
      contacts.push_back(std::make_unique<emcGateContact<T>>(appliedVoltage, epsOxide, thickness, barrierHeight));
      break;
    case emcContactType::OHMIC:
      contacts.push_back(std::make_unique<emcOhmicContact<T>>(appliedVoltage));
      break;
    case emcContactType::SCHOTTKY:
      contacts.push_back(std::make_unique<emcSchottkyContact<T>>(appliedVoltage, barrierHeight));
      break;
    }
  }

  bool isArtificialBoundary(const SizeVecSurface &s, emcBoundaryPos face) const { return getContactIdx(s, face) == -1; }
  bool isContactType(const SizeVecSurface &s, emcBoundaryPos face, emcContactType type) const {
    const int idx = getContactIdx(s, face);
    return idx != -1 && getContactType(idx) == type;
  }
  bool isOhmicContact(const SizeVecDevice &c) const { return cellHasType(c, emcContactType::OHMIC); }
  bool isSchottkyContact(const SizeVecDevice &c) const { return cellHasType(c, emcContactType::SCHOTTKY); }
  bool isReservoirContact(const SizeVecDevice &c) const { return isOhmicContact(c) || isSchottkyContact(c); }

  // ---- iteration over all face cells: for (initCoord(f, s); !isEndCoord(f, s); advanceCoord(f, s)) -------------
  void initCoord(emcBoundaryPos &face, SizeVecSurface &s) const {
    face = emcBoundaryPos::XMIN;
    s.fill(0);
  }
  void advanceCoord(emcBoundaryPos &face, SizeVecSurface &s) const {
    const auto &grid = contactOfFace[toUnderlying(face)];
    grid.advanceCoord(s);
    if (grid.isEndCoord(s) && face != faces.back()) {
      s.fill(0);
      face = static_cast<emcBoundaryPos>(toUnderlying(face) + 1);
    }
  }
  bool isEndCoord(emcBoundaryPos &face, SizeVecSurface &s) const {
    return face == faces.back() && contactOfFace.back().isEndCoord(s);
  }

  void print(std::ostream &out = std::cout) const {
    for (const auto &g : contactOfFace) {
      g.print(out);
      out << "\n";
    }
  }
  void print(emcBoundaryPos face, std::ostream &out = std::cout) const {
    if (toUnderlying(face) < 2 * DimDevice)
      contactOfFace[toUnderlying(face)].print(out);
  }

private:
  bool cellHasType(const SizeVecDevice &c, emcContactType type) const {
    const auto face = getBoundaryPos(c);
    return face != emcBoundaryPos::INVALID && isContactType(getCoordBoundary(c, face), face, type);
  }
};

#endif
