// Doping concentration and doping-region index per grid point.
// Interface mirrored: reference include/emcDopingProfile.hpp (region -1 = undoped
// background at Ni; regions are numbered in the order they are added; later
// regions overwrite earlier ones where they overlap).
#ifndef EMC_DOPING_PROFILE_HPP
#define EMC_DOPING_PROFILE_HPP

#include <map>

#include <emcGrid.hpp>
#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <class T, SizeType Dim> class emcDopingProfile {
  typedef std::array<SizeType, Dim> SizeVec;

  SizeType nrRegions = 0;
  emcGrid<int, Dim> idxDopingRegions;
  emcGrid<T, Dim> doping;
  std::map<int, T> regionDoping;
  T normalization;

  void clipAndOrder(SizeVec &lo, SizeVec &hi) {
    const auto ext = doping.getExtent();
    bool clipped = false, swapped = false;
    for (SizeType d = 0; d < Dim; d++) {
      if (std::min(lo[d], hi[d]) > ext[d])
        emcMessage::getInstance().addError("Added Region is completely out of bounds.").print();
      if (lo[d] > ext[d] || hi[d] > ext[d]) {
        clipped = true;
        lo[d] = std::min(lo[d], ext[d]);
        hi[d] = std::min(hi[d], ext[d]);
      }
      if (hi[d] < lo[d]) {
        swapped = true;
        std::swap(lo[d], hi[d]);
      }
    }
    if (clipped)
      emcMessage::getInstance().addWarning("Doping region was partially out of bounds and has been clipped.").print();
    if (swapped)
      emcMessage::getInstance().addWarning("Doping region had min > max in one direction; the two were swapped.").print();
  }

public:
  emcDopingProfile() = delete;
  emcDopingProfile(const SizeVec &gridExtent, T Ni) : emcDopingProfile(gridExtent, Ni, 1.) {}
  emcDopingProfile(const SizeVec &gridExtent, T Ni, T inNormalization)
      : idxDopingRegions(gridExtent, -1), doping(gridExtent, Ni), normalization(inNormalization) {
    regionDoping[-1] = Ni;
  }

  void addConstantDopingRegion(SizeVec minCoord, SizeVec maxCoord, T inDoping) {
    clipAndOrder(minCoord, maxCoord);
    idxDopingRegions.fill(static_cast<int>(nrRegions), minCoord, maxCoord);
    doping.fill(inDoping, minCoord, maxCoord);
    regionDoping[static_cast<int>(nrRegions)] = inDoping;
    nrRegions++;
  }

  const emcGrid<int, Dim> &getDopingRegionIdx() const { return idxDopingRegions; }
  int getDopingRegionIdx(const SizeVec &coord) const { return idxDopingRegions[coord]; }
  emcGrid<T, Dim> getDoping(bool normalized = false) const {
    emcGrid<T, Dim> out = doping;
    if (normalized)
      for (auto &x : out)
        x /= normalization;
    return out;
  }
  T getDoping(const SizeVec &coord, bool normalized = false) const {
    return normalized ? doping[coord] / normalization : doping[coord];
  }
  T getDoping(int idxRegion, bool normalized = false) const {
    if (idxRegion < -1 || idxRegion >= static_cast<int>(nrRegions))
      emcMessage::getInstance().addError("Idx for Doping Region is out of bounds.").print();
    const T d = regionDoping.find(idxRegion)->second;
    return normalized ? d / normalization : d;
  }
  T getDoping(SizeType idxRegion, bool normalized = false) const {
    return getDoping(static_cast<int>(idxRegion), normalized);
  }
  SizeType getNrDopingRegions() const { return nrRegions; }

  template <class, SizeType> friend class emcDevice;
};

#endif
