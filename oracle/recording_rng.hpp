// TEST INFRASTRUCTURE (oracle/). Not product code.
// Drop-in replacement for the reference's `emcRNG` typedef (reference:
// include/emcUtil.hpp:15, `typedef std::mt19937_64 emcRNG`).  It forwards to a
// std::mt19937_64 and appends every raw 64-bit output to a global log so that
// the unmodified reference step loop can be replayed draw-for-draw elsewhere.
#pragma once
#include <cstdint>
#include <random>
#include <vector>

struct RecordingRNG {
  using result_type = std::uint64_t;
  std::mt19937_64 engine;
  static std::vector<std::uint64_t> *&sink() {
    static std::vector<std::uint64_t> *s = nullptr;
    return s;
  }
  RecordingRNG() : engine() {}
  explicit RecordingRNG(result_type s) : engine(s) {}
  void seed(result_type s) { engine.seed(s); }
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return ~result_type(0); }
  result_type operator()() {
    result_type x = engine();
    if (sink())
      sink()->push_back(x);
    return x;
  }
};
