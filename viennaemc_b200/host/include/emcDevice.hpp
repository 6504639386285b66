// Simulation box: geometry, material, lattice temperature, doping regions and the
// normalisations derived from them.
// Interface mirrored: reference include/emcDevice.hpp (ctor :78-100, getters :113-160,
// addConstantDopingRegion :166-171, normalisation helpers :236-262, coordinate
// helpers :265-291, cell volume :333-343).
// Contacts: addOhmicContact / addGateContact / addSchottkyContact (:178-215), positions along a
// face given in the face's own axes (posToCoord(boundPos, position) :293-330).
#ifndef EMC_DEVICE_HPP
#define EMC_DEVICE_HPP

#include <cmath>
#include <stdexcept>

#include <emcBoundaryPos.hpp>
#include <emcConstants.hpp>
#include <emcDopingProfile.hpp>
#include <emcGrid.hpp>
#include <emcMaterial.hpp>
#include <emcMessage.hpp>
#include <emcSurface.hpp>
#include <emcUtil.hpp>

template <class T, SizeType Dim> class emcDevice {
  static_assert(Dim == 2 || Dim == 3, "Wrong Dimension for Device, possible dimensions are: {2,3}.");

public:
  static const SizeType Dimension = Dim;
  typedef T ValueType;
  typedef emcMaterial<T> MaterialType;
  typedef emcDopingProfile<T, Dim> DopingProfileType;
  typedef std::array<T, Dim> ValueVec;
  typedef std::array<SizeType, Dim> SizeVec;
  typedef emcSurface<T, Dim - 1> SurfaceType;
  typedef std::array<SizeType, Dim - 1> SizeVecSurface;
  typedef std::array<T, Dim - 1> ValueVecSurface;

private:
  MaterialType material;
  ValueVec maxPos, spacing;
  T temperature;
  T deviceWidth{1e-6}; // third dimension of a 2-D device [m]
  T cellVolume, normedCellVolume;
  T thermalVoltage, debyeLength;
  DopingProfileType dopingProfile;
  SurfaceType surface;

  void updateCellVolume() {
    cellVolume = 1.;
    normedCellVolume = 1.;
    for (SizeType d = 0; d < Dim; d++) {
      cellVolume *= spacing[d];
      normedCellVolume *= spacing[d] / debyeLength;
    }
    if (Dim == 2) {
      cellVolume *= deviceWidth;
      normedCellVolume *= deviceWidth / debyeLength;
    }
  }

public:
  emcDevice() = delete;
  emcDevice(MaterialType inMaterial, ValueVec inMaxPos, ValueVec inSpacing, T inTemperature = 300)
      : material(inMaterial), maxPos(inMaxPos), spacing(inSpacing), temperature(inTemperature),
        thermalVoltage(constants::kB / constants::q * temperature),
        debyeLength(std::sqrt(constants::eps0 * material.getEpsR() * thermalVoltage / constants::q / material.getNi())),
        dopingProfile(maxPosToExtent(inMaxPos, inSpacing), material.getNi(), material.getNi()),
        surface(maxPosToExtent(inMaxPos, inSpacing), thermalVoltage) {
    if (!(debyeLength > 0) || !std::isfinite(debyeLength))
      throw std::domain_error("emcDevice: computed Debye length is non-positive or non-finite. Check the material's "
                              "intrinsic carrier concentration (Ni) and permittivity.");
    updateCellVolume();
  }

  template <SizeType D = Dim> typename std::enable_if<(D == 2)>::type setDeviceWidth(T inDeviceWidth) {
    deviceWidth = inDeviceWidth;
    updateCellVolume();
  }

  const MaterialType &getMaterial() const { return material; }
  const DopingProfileType &getDopingProfile() const { return dopingProfile; }
  const SurfaceType &getSurface() const { return surface; }
  T getTemperature() const { return temperature; }
  T getThermalVoltage() const { return thermalVoltage; }
  T getDebyeLength() const { return debyeLength; }
  ValueVec getSpacing(bool normalized = false) const { return normalized ? normalizeLength(spacing) : spacing; }
  ValueVec getMaxPos(bool normalized = false) const { return normalized ? normalizeLength(maxPos) : maxPos; }
  T getCellVolume(bool normalized = false) const { return normalized ? normedCellVolume : cellVolume; }
  SizeVec getGridExtent() const { return dopingProfile.doping.getExtent(); }

  void addConstantDopingRegion(ValueVec inMinPos, ValueVec inMaxPos, T inDoping) {
    dopingProfile.addConstantDopingRegion(posToCoord(inMinPos), posToCoord(inMaxPos), inDoping);
  }

  // minPos / maxPos: extent of the contact along the face, in the face's own coordinates [m]
  void addOhmicContact(emcBoundaryPos boundaryPos, T voltage, ValueVecSurface minPos, ValueVecSurface maxPos) {
    surface.addContact(boundaryPos, emcContactType::OHMIC, voltage, posToCoord(boundaryPos, minPos),
                       posToCoord(boundaryPos, maxPos));
  }
  void addGateContact(emcBoundaryPos boundaryPos, T voltage, ValueVecSurface minPos, ValueVecSurface maxPos, T epsRoxide,
                      T thickness, T barrierHeight) {
    surface.addContact(boundaryPos, emcContactType::GATE, voltage, posToCoord(boundaryPos, minPos),
                       posToCoord(boundaryPos, maxPos), epsRoxide, thickness, barrierHeight);
  }
  void addSchottkyContact(emcBoundaryPos boundaryPos, T voltage, ValueVecSurface minPos, ValueVecSurface maxPos,
                          T barrierHeight) {
    surface.addContact(boundaryPos, emcContactType::SCHOTTKY, voltage, posToCoord(boundaryPos, minPos),
                       posToCoord(boundaryPos, maxPos), 0, 0, barrierHeight);
  }

  bool isOutOfBounds(ValueVec &position) const {
    for (SizeType d = 0; d < Dim; d++)
      if (position[d] < 0 || position[d] > maxPos[d])
        return true;
    return false;
  }

  T normalizeLength(T length) const { return length / debyeLength; }
  ValueVec normalizeLength(const ValueVec &length) const {
    ValueVec out = length;
    for (auto &x : out)
      x /= debyeLength;
    return out;
  }
  T normalizeDoping(T doping) const { return doping / material.getNi(); }
  T normalizeVoltage(T voltage) const { return voltage / thermalVoltage; }
  T undoNormalizeLength(T normLength) const { return normLength * debyeLength; }
  T undoNormalizeDoping(T normDoping) const { return normDoping * material.getNi(); }
  T undoNormalizeVoltage(T normVoltage) const { return normVoltage * thermalVoltage; }

  void advanceCoord(SizeVec &coord) const { dopingProfile.doping.advanceCoord(coord); }
  bool isEndCoord(const SizeVec &coord) const { return dopingProfile.doping.isEndCoord(coord); }
  ValueVec coordToPos(const SizeVec &coord) const {
    ValueVec pos;
    for (SizeType d = 0; d < Dim; d++)
      pos[d] = coord[d] * spacing[d];
    return pos;
  }
  SizeVec posToCoord(const ValueVec &position) const {
    SizeVec coord;
    for (SizeType d = 0; d < Dim; d++)
      coord[d] = static_cast<SizeType>(std::round(position[d] / spacing[d]));
    return coord;
  }
  // position along a face -> face coordinates (the axes of the face keep their device order)
  SizeVecSurface posToCoord(emcBoundaryPos boundPos, const ValueVecSurface &position) const {
    const SizeType fixed = toUnderlying(boundPos) / 2;
    SizeVecSurface coord;
    for (SizeType d = 0, o = 0; d < Dim; d++)
      if (d != fixed) {
        coord[o] = static_cast<SizeType>(std::round(position[o] / spacing[d]));
        o++;
      }
    return coord;
  }
};

#endif
