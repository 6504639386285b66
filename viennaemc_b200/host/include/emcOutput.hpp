// Text writers of the result files.  Interface mirrored: reference include/emcOutput.hpp:14-72 (the
// formats helper/emcPlottingFiles/readResultFile.py reads): a grid as rows of x-values; a grid with a
// leading extent line and a per-value conversion; the per-step contact counters and currents.
#ifndef EMC_OUTPUT_HPP
#define EMC_OUTPUT_HPP

#include <fstream>
#include <string>
#include <vector>

#include <emcGrid.hpp>
#include <emcUtil.hpp>

template <class T, SizeType Dim> void writeToFile(const emcGrid<T, Dim> &grid, std::string fileName) {
  std::ofstream out(fileName + ".txt");
  grid.print(out);
}

// extent line, then rows of x-values, a blank line between z-planes; every value passed through func
template <class T, class DeviceType, SizeType Dim>
void writeToFile(const emcGrid<T, Dim> &grid, std::string fileName, T (*func)(const T &, const DeviceType &device),
                 const DeviceType &device) {
  std::ofstream out(fileName + ".txt");
  const auto extent = grid.getExtent();
  out << extent << "\n";
  typename emcGrid<T, Dim>::CoordVec c;
  for (c.fill(0); !grid.isEndCoord(c); grid.advanceCoord(c)) {
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
    out << func(grid[c], device);
    if (c[0] != extent[0] - 1)
      out << " ";
  }
  out << "\n";
}

// per non-transient step: time, net particle count per contact, current per contact
template <class T>
void writeToFile(const std::vector<std::vector<std::vector<int>>> &nettoNrPart,
                 const std::vector<std::vector<std::vector<T>>> &current, SizeType idxType, T stepTime, T startTime,
                 std::string fileName) {
  std::ofstream out(fileName + ".txt");
  for (SizeType step = 1; step <= current.size(); step++) {
    out << startTime + step * stepTime;
    for (auto n : nettoNrPart[step - 1][idxType])
      out << " " << n;
    for (auto c : current[step - 1][idxType])
      out << " " << c;
    out << "\n";
  }
}

template <class T, class DeviceType> T undoNormalizationPotential(const T &val, const DeviceType &device) {
  return device.undoNormalizeVoltage(val);
}
template <class T, class DeviceType> T undoNormalizationConcentration(const T &val, const DeviceType &device) {
  return device.undoNormalizeDoping(val);
}

#endif
