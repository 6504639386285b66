// Contacts on the faces of the box.  Interface mirrored: reference include/emcContact.hpp
// (emcContactType, emcContact and its three subclasses with getFurtherParameter).
#ifndef EMC_CONTACT_HPP
#define EMC_CONTACT_HPP

#include <emcMessage.hpp>
#include <emcUtil.hpp>

enum struct emcContactType : SizeType { OHMIC, SCHOTTKY, GATE };

template <class T> class emcContact {
protected:
  T appliedVoltage; // [V]

public:
  explicit emcContact(T inAppliedVoltage) : appliedVoltage(inAppliedVoltage) {}
  virtual ~emcContact() = default;
  T getAppliedVoltage() const { return appliedVoltage; }
  virtual T getFurtherParameter(SizeType idxParameter) const = 0;
  virtual emcContactType getType() const = 0;
};

template <class T> class emcOhmicContact : public emcContact<T> {
public:
  explicit emcOhmicContact(T inAppliedVoltage) : emcContact<T>(inAppliedVoltage) {}
  T getFurtherParameter(SizeType) const override {
    emcMessage::getInstance().addError("Ohmic Contact has no further parameter.").print();
    return 0;
  }
  emcContactType getType() const override { return emcContactType::OHMIC; }
};

// parameters: 0 relative permittivity of the oxide, 1 oxide thickness [m], 2 barrier height [V]
template <class T> class emcGateContact : public emcContact<T> {
  T parameter[3];

public:
  emcGateContact(T inAppliedVoltage, T inEpsR, T inThickness, T inBarrierHeight)
      : emcContact<T>(inAppliedVoltage), parameter{inEpsR, inThickness, inBarrierHeight} {}
  T getFurtherParameter(SizeType idxParameter) const override {
    if (idxParameter > 2) {
      emcMessage::getInstance().addError("Gate Contact only has 3 further parameter.").print();
      return 0;
    }
    return parameter[idxParameter];
  }
  emcContactType getType() const override { return emcContactType::GATE; }
};

// parameter 0: barrier height [V]
template <class T> class emcSchottkyContact : public emcContact<T> {
  T barrierHeight;

public:
  emcSchottkyContact(T inAppliedVoltage, T inBarrierHeight) : emcContact<T>(inAppliedVoltage), barrierHeight(inBarrierHeight) {}
  T getFurtherParameter(SizeType idxParameter) const override {
    if (idxParameter != 0) {
      emcMessage::getInstance().addError("Schottky Contact only has 1 further parameter.").print();
      return 0;
    }
    return barrierHeight;
  }
  emcContactType getType() const override { return emcContactType::SCHOTTKY; }
};

#endif
