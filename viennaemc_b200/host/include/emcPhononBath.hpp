// Non-equilibrium population of one polar (LO) phonon mode, resolved in |q| bins, with optional acoustic (Klemens) and
// TO (Ridley) reservoirs behind it.
// Interface mirrored: reference include/emcPhononBath.hpp -- public members nrBins, dq, tauLO, tauLOProfile, N0, Vsim,
// Nq, nEm, nAbs, cumW, cumWN, qs2Screen; ctor :161-196; setScreeningQ2 :200-205; recordEmission / recordAbsorption
// :237-253; update :264-358; getNqInWindow :378-394; sampleQ :423-458; getMeanNq :461-470 and the small getters.
//
// Division of labour on the GPU path: the particle kernels count emission / absorption events per bin on the device
// (emcgpu_get_phonon_counts); the GPU particle handler adds them to nEm / nAbs after every step; the relaxation
// (update) and everything that feeds the HOST-side rate tables stays here, 300 bins of arithmetic per step.  The
// q-resolved polar angle is sampled on the device from copies of cumW / cumWN.  Expressions are evaluated in the
// reference's order: a seeded run reproduces the reference's occupations bit for bit.
#ifndef EMC_PHONON_BATH_HPP
#define EMC_PHONON_BATH_HPP

#include <algorithm>
#include <cmath>
#include <vector>

#include <emcConstants.hpp>
#include <emcUtil.hpp>

template <class T> class emcPhononBath {
public:
  SizeType nrBins;
  T dq;    // bin width [1/m]
  T tauLO; // LO -> reservoir decay time [s]
  std::vector<T> tauLOProfile; // optional per-bin decay times
  T N0;    // Bose-Einstein occupation of the mode at the lattice temperature
  T Vsim;  // simulated volume [m^3]
  std::vector<T> Nq, nEm, nAbs;
  std::vector<T> cumW, cumWN; // prefix sums of the coupling weight q / (q^2 + qs^2) and of weight * Nq
  T qs2Screen = T(0);

private:
  struct Reservoir { // a phonon population that relaxes to the lattice with its own time constant
    bool on = false;
    T energyEV = T(0), tau = T(0), N = T(0), Neq = T(0);
    T temperature(T lattice) const { return (!on || N <= Neq) ? lattice : occupationToTemp(energyEV, N); }
    void relax(T dt, T feed) {
      N += feed - dt * (N - Neq) / tau;
      if (N < Neq)
        N = Neq; // cannot cool below the lattice
    }
  };
  T loEnergyEV, latticeTempK;
  Reservoir acoustic, transverse;
  T wRidley = T(0);
  // getMeanNq() is asked once per energy level and mechanism at every table rebuild (thousands of times per time step)
  // for an occupation that only changes in update(): the value is kept, guarded by three sample bins in case Nq (a
  // public member, as in the reference) was written directly
  mutable T cachedMean = T(0);
  mutable T cacheGuard[3] = {T(-1), T(-1), T(-1)};
  mutable bool cacheValid = false;

  static T occupation(T energyEV, T tempK) {
    const T x = constants::q * energyEV / (constants::kB * tempK);
    return T(1) / (std::exp(x) - T(1));
  }
  static T occupationToTemp(T energyEV, T N) {
    return N <= T(0) ? T(0) : constants::q * energyEV / (constants::kB * std::log(T(1) + T(1) / N));
  }
  // LO occupation in equilibrium with a (possibly heated) reservoir
  T loOccupationFedBy(const Reservoir &r) const {
    return r.N > r.Neq ? occupation(loEnergyEV, occupationToTemp(r.energyEV, r.N)) : N0;
  }
  void rebuildWindowSums() {
    cumW.assign(nrBins + 1, T(0));
    cumWN.assign(nrBins + 1, T(0));
    for (SizeType i = 0; i < nrBins; i++) {
      const T q = qCentre(i);
      const T denom = q * q + qs2Screen;
      const T w = denom > T(0) ? q / denom : T(0);
      cumW[i + 1] = cumW[i] + w;
      cumWN[i + 1] = cumWN[i] + w * Nq[i];
    }
  }
  // bin range [lo, hi) that covers [qMin, qMax]; false if it is empty
  bool window(T qMin, T qMax, long &lo, long &hi) const {
    lo = std::max(0L, static_cast<long>(std::floor(qMin / dq)));
    hi = std::min(static_cast<long>(nrBins), static_cast<long>(std::ceil(qMax / dq)));
    return hi - lo >= 1;
  }

public:
  emcPhononBath() = delete;
  emcPhononBath(SizeType inNrBins, T inDq, T inTauLO, T phononEnergy, T latticeTemp, T inVsim, bool inEnableAcoustic = false,
                T inAcPhononEnergy = T(0), T inTauAcoustic = T(0), T inWRidley = T(0), T inToPhononEnergy = T(0),
                T inTauTO = T(0))
      : nrBins(inNrBins), dq(inDq), tauLO(inTauLO), N0(occupation(phononEnergy, latticeTemp)), Vsim(inVsim), Nq(inNrBins, N0),
        nEm(inNrBins, T(0)), nAbs(inNrBins, T(0)), loEnergyEV(phononEnergy), latticeTempK(latticeTemp) {
    if (inEnableAcoustic) {
      acoustic.on = true;
      acoustic.energyEV = inAcPhononEnergy;
      acoustic.tau = inTauAcoustic;
      acoustic.N = acoustic.Neq = occupation(inAcPhononEnergy, latticeTemp);
      if (inWRidley > T(0) && inToPhononEnergy > T(0)) {
        transverse.on = true;
        wRidley = std::min(inWRidley, T(1));
        transverse.energyEV = inToPhononEnergy;
        transverse.tau = inTauTO;
        transverse.N = transverse.Neq = occupation(inToPhononEnergy, latticeTemp);
      }
    }
    rebuildWindowSums();
  }

  void setScreeningQ2(T inQs2) {
    if (inQs2 != qs2Screen) {
      qs2Screen = inQs2;
      rebuildWindowSums();
    }
  }
  T getScreeningQ2() const { return qs2Screen; }
  void setTauLOProfile(std::vector<T> inProfile) { tauLOProfile = std::move(inProfile); }

  T qCentre(SizeType i) const { return (T(i) + T(0.5)) * dq; }
  SizeType binOf(T q) const { return std::min<SizeType>(static_cast<SizeType>(q / dq), nrBins - 1); }
  T getNq(T q) const { return Nq[binOf(q)]; }
  void recordEmission(T q) { nEm[binOf(q)] += T(1); }
  void recordAbsorption(T q) { nAbs[binOf(q)] += T(1); }

  // one time step: net generation from the event counters, decay towards the occupation the reservoirs dictate,
  // reservoirs heated by the decayed LO phonons
  void update(T dt) {
    T target = N0;
    if (acoustic.on) {
      const T viaKlemens = loOccupationFedBy(acoustic);
      target = transverse.on ? (T(1) - wRidley) * viaKlemens + wRidley * loOccupationFedBy(transverse) : viaKlemens;
    }
    const bool profile = !tauLOProfile.empty();
    T sumW = T(0), sumExcess = T(0), sumExcessRate = T(0);
    for (SizeType i = 0; i < nrBins; i++) {
      const T q = qCentre(i);
      const T modes = q * q * dq * Vsim / (T(2) * constants::pi * constants::pi); // phonon modes of the bin
      const T g = modes > T(0) ? (nEm[i] - nAbs[i]) / (modes * dt) : T(0);
      const T tau = profile ? tauLOProfile[i] : tauLO;
      if (acoustic.on) {
        const T w = q * q;
        sumW += w;
        if (profile)
          sumExcessRate += w * (Nq[i] - target) / tau;
        else
          sumExcess += w * (Nq[i] - target);
      }
      Nq[i] += g * dt - (dt / tau) * (Nq[i] - target);
      if (Nq[i] < T(0))
        Nq[i] = T(0);
      nEm[i] = T(0);
      nAbs[i] = T(0);
    }
    cacheValid = false;
    if (acoustic.on) {
      const T toKlemens = transverse.on ? (T(1) - wRidley) : T(1);
      if (!profile) {
        const T meanExcess = sumW > T(0) ? sumExcess / sumW : T(0);
        acoustic.relax(dt, dt * toKlemens * meanExcess / tauLO);
        if (transverse.on)
          transverse.relax(dt, dt * wRidley * meanExcess / tauLO);
      } else {
        const T decayFlux = sumW > T(0) ? sumExcessRate / sumW : T(0);
        acoustic.relax(dt, dt * toKlemens * decayFlux);
        if (transverse.on)
          transverse.relax(dt, dt * wRidley * decayFlux);
      }
    }
    rebuildWindowSums();
  }

  // coupling-weighted mean occupation of the phonons a transition |k - k'| <= q <= k + k' can exchange
  T getNqInWindow(T qMin, T qMax) const {
    long lo, hi;
    if (nrBins < 2 || qMax <= qMin || !window(qMin, qMax, lo, hi))
      return getMeanNq();
    const T wSum = cumW[hi] - cumW[lo];
    return wSum <= T(0) ? getMeanNq() : (cumWN[hi] - cumWN[lo]) / wSum;
  }

  // |q| from the occupation-weighted coupling in the window (emission: N + 1, absorption: N), inverse-CDF with r in [0,1)
  T sampleQ(T qMin, T qMax, bool emission, T r) const {
    long lo, hi;
    if (nrBins < 2 || qMax <= qMin || !window(qMin, qMax, lo, hi))
      return qMin;
    const auto S = [&](long i) { return emission ? (cumWN[i] + cumW[i]) : cumWN[i]; };
    const T sLo = S(lo), span = S(hi) - sLo;
    if (!(span > T(0)))
      return T(0.5) * (qMin + qMax);
    const T goal = sLo + r * span;
    long a = lo, b = hi;
    while (b - a > 1) {
      const long mid = (a + b) / 2;
      (S(mid) <= goal ? a : b) = mid;
    }
    const T sA = S(a), sB = S(a + 1);
    const T frac = sB > sA ? (goal - sA) / (sB - sA) : T(0.5);
    return std::max(qMin, std::min(qMax, (T(a) + frac) * dq));
  }

  // density-of-states (q^2) weighted mean occupation
  T getMeanNq() const {
    const T guard[3] = {Nq.front(), Nq[nrBins / 2], Nq.back()};
    if (cacheValid && guard[0] == cacheGuard[0] && guard[1] == cacheGuard[1] && guard[2] == cacheGuard[2])
      return cachedMean;
    T sumW = T(0), sumWN = T(0);
    for (SizeType i = 0; i < nrBins; i++) {
      const T q = qCentre(i), w = q * q;
      sumW += w;
      sumWN += w * Nq[i];
    }
    cachedMean = sumW > T(0) ? sumWN / sumW : N0;
    for (int i = 0; i < 3; i++)
      cacheGuard[i] = guard[i];
    cacheValid = true;
    return cachedMean;
  }

  T getMeanNac() const { return acoustic.N; }
  T getAcousticN0() const { return acoustic.Neq; }
  T getAcousticTemp() const { return acoustic.temperature(latticeTempK); }
  T getMeanNTO() const { return transverse.N; }
  T getTOTemp() const { return transverse.temperature(latticeTempK); }
  bool acousticBathEnabled() const { return acoustic.on; }
  bool ridleyEnabled() const { return transverse.on; }
  T getRidleyBranching() const { return wRidley; }
  T getN0() const { return N0; }
  T getTauLO() const { return tauLO; }
  T getLOEnergy() const { return loEnergyEV; }
};

#endif
