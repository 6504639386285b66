// Pauli blocking of final states on a k-space occupancy grid.
// Interface mirrored: reference include/emcPauliExclusion.hpp (ctor: grid spacing dk, kMax, simulated volume; counters
// nRejected / nScattered) -- as far as drivers of the bulk handler need it to compile.
//
// Band filling makes the particle loop SEQUENTIAL in the reference (every accepted event changes the occupancy the next
// particle sees, basicBulkParticleHandler.hpp:417-418); it is outside the data-parallel path this library accelerates.
// The GPU bulk handler rejects moveParticleTypeWithBandFilling() with an error instead of running anything on the CPU.
#ifndef EMC_PAULI_EXCLUSION_HPP
#define EMC_PAULI_EXCLUSION_HPP

#include <emcUtil.hpp>

template <class T> class emcPauliExclusion {
public:
  SizeType nRejected = 0;
  SizeType nScattered = 0;
  T dk, kMax, Vsim;
  emcPauliExclusion() = delete;
  emcPauliExclusion(T inDk, T inKMax, T inVsim) : dk(inDk), kMax(inKMax), Vsim(inVsim) {}
  // fraction of the attempted events that were blocked (reference :140-144); nothing is ever attempted here
  T getRejectionRate() const { return nScattered ? T(nRejected) / T(nScattered) : T(0); }
};

#endif
