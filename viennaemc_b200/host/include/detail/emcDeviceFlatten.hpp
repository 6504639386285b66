// emcDevice + emcDopingProfile + emcSurface as the flat arrays of emcgpu_device_t (include/emcgpu.h).
#ifndef EMC_DETAIL_DEVICE_FLATTEN_HPP
#define EMC_DETAIL_DEVICE_FLATTEN_HPP

#include <cstdint>
#include <vector>

#include <emcgpu.h>

#include <emcDevice.hpp>

namespace emcdetail {

template <class T, SizeType Dim> struct FlatDevice {
  emcgpu_device_t desc{};
  std::vector<int32_t> contactType, region;
  std::vector<double> contactVoltage, gateEpsOx, gateThickness, gateBarrier, doping;
  std::vector<int8_t> faceContact;

  explicit FlatDevice(const emcDevice<T, Dim> &device) {
    const auto extent = device.getGridExtent();
    const auto &surface = device.getSurface();
    const auto spacing = device.getSpacing();
    const auto maxPos = device.getMaxPos();
    desc.dim = Dim;
    desc.nContacts = static_cast<int32_t>(surface.getNrContacts());
    SizeType cells = 1;
    for (SizeType d = 0; d < 3; d++) {
      desc.extent[d] = d < Dim ? static_cast<int32_t>(extent[d]) : 1;
      desc.spacing[d] = d < Dim ? spacing[d] : 1.;
      desc.maxPos[d] = d < Dim ? maxPos[d] : 0.;
      if (d < Dim)
        cells *= extent[d];
    }
    desc.thermalVoltage = device.getThermalVoltage();
    desc.debyeLength = device.getDebyeLength();
    desc.ni = device.getMaterial().getNi();
    desc.cellVolume = device.getCellVolume();
    desc.epsR = device.getMaterial().getEpsR();
    for (SizeType c = 0; c < surface.getNrContacts(); c++) {
      const auto type = surface.getContactType(c);
      contactType.push_back(type == emcContactType::OHMIC ? EMCGPU_CONTACT_OHMIC
                            : type == emcContactType::SCHOTTKY ? EMCGPU_CONTACT_SCHOTTKY
                                                               : EMCGPU_CONTACT_GATE);
      contactVoltage.push_back(surface.getContactVoltage(c));
      const bool gate = type == emcContactType::GATE;
      gateEpsOx.push_back(gate ? surface.getContactFurtherParameter(c, 0) : 0.);
      gateThickness.push_back(gate ? surface.getContactFurtherParameter(c, 1) : 0.);
      gateBarrier.push_back(gate ? surface.getContactFurtherParameter(c, 2) : 0.);
    }
    region.reserve(cells);
    doping.reserve(cells);
    faceContact.reserve(cells * 2 * Dim);
    typename emcDevice<T, Dim>::SizeVec coord;
    for (coord.fill(0); !device.isEndCoord(coord); device.advanceCoord(coord)) {
      region.push_back(device.getDopingProfile().getDopingRegionIdx(coord));
      doping.push_back(device.getDopingProfile().getDoping(coord));
      for (SizeType f = 0; f < 2 * Dim; f++) {
        const auto face = static_cast<emcBoundaryPos>(f);
        int8_t v = -2;
        if (surface.isOnBoundary(coord, face))
          v = static_cast<int8_t>(surface.getContactIdx(surface.getCoordBoundary(coord, face), face));
        faceContact.push_back(v);
      }
    }
    desc.contactType = contactType.data();
    desc.contactVoltage = contactVoltage.data();
    desc.gateEpsOx = gateEpsOx.data();
    desc.gateThickness = gateThickness.data();
    desc.gateBarrier = gateBarrier.data();
    desc.region = region.data();
    desc.faceContact = faceContact.data();
    desc.doping = doping.data();
  }
};

} // namespace emcdetail

#endif
