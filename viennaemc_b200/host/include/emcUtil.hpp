// Small vector / RNG helpers of the drop-in host API.
// Interface mirrored: reference include/emcUtil.hpp (SizeType, emcRNG, square, norm,
// normalize, innerProduct, scale, add, subtract, stream operators, maxPosToExtent,
// initRandomDirection, initRandomDirectionWithRespectToCurrentK).
// Summation / multiplication orders follow the reference exactly because the host
// side builds the rate tables and the initial ensemble that the GPU path consumes.
#ifndef EMC_UTIL_HPP
#define EMC_UTIL_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <random>
#include <type_traits>
#include <vector>

typedef size_t SizeType;
typedef std::mt19937_64 emcRNG; // reference include/emcUtil.hpp:15

template <typename Enum> constexpr auto toUnderlying(Enum e) {
  return static_cast<typename std::underlying_type<Enum>::type>(e);
}

template <class T, SizeType Dim> T square(const std::array<T, Dim> &v) {
  T acc = 0;
  for (SizeType i = 0; i < Dim; i++)
    acc += v[i] * v[i];
  return acc;
}
template <class T, SizeType Dim> T norm(const std::array<T, Dim> &v) { return std::sqrt(square(v)); }
template <class T, SizeType Dim> void normalize(std::array<T, Dim> &v) {
  const T len = norm(v);
  if (len == T(0))
    return;
  for (auto &x : v)
    x /= len;
}
template <class T> T innerProduct(const std::array<T, 3> &a, const std::array<T, 3> &b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
template <class T, SizeType Dim> std::array<T, Dim> scale(const std::array<T, Dim> &v, T f) {
  std::array<T, Dim> r;
  for (SizeType i = 0; i < Dim; i++)
    r[i] = v[i] * f;
  return r;
}
template <class T, SizeType Dim> std::array<T, Dim> add(const std::array<T, Dim> &a, const std::array<T, Dim> &b) {
  std::array<T, Dim> r;
  for (SizeType i = 0; i < Dim; i++)
    r[i] = a[i] + b[i];
  return r;
}
template <class T, SizeType Dim>
std::array<T, Dim> subtract(const std::array<T, Dim> &a, const std::array<T, Dim> &b) {
  std::array<T, Dim> r;
  for (SizeType i = 0; i < Dim; i++)
    r[i] = a[i] - b[i];
  return r;
}

// "a b c" (single blanks, no trailing blank): the format of every result file
template <class T, SizeType Dim> std::ostream &operator<<(std::ostream &os, const std::array<T, Dim> &v) {
  for (SizeType i = 0; i < Dim; i++)
    os << (i ? " " : "") << v[i];
  return os;
}
template <class T> std::ostream &operator<<(std::ostream &os, const std::vector<T> &v) {
  for (SizeType i = 0; i < v.size(); i++)
    os << (i ? " " : "") << v[i];
  return os;
}

// number of grid points per dimension of a box [0, maxPos] with the given spacing
template <class T, SizeType Dim>
std::array<SizeType, Dim> maxPosToExtent(const std::array<T, Dim> &maxPos, const std::array<T, Dim> &spacing) {
  std::array<SizeType, Dim> extent;
  for (SizeType i = 0; i < Dim; i++)
    extent[i] = static_cast<SizeType>(std::round(maxPos[i] / spacing[i]) + 1);
  return extent;
}

template <class T> constexpr T emcPi() { return T(3.14159265358979323846L); }

// isotropic direction: phi = 2 pi rand1, cos(theta) = 1 - 2 rand2
template <class T> std::array<T, 3> initRandomDirection(T length, T rand1, T rand2) {
  const T phi = 2 * emcPi<T>() * rand1;
  const T c = 1 - 2 * rand2;
  return {length * std::sqrt(1 - c * c) * std::cos(phi), length * std::sqrt(1 - c * c) * std::sin(phi), length * c};
}

// new direction with polar cosine cosTheta about the current k and azimuth 2 pi rand; |k| is kept
template <class T>
std::array<T, 3> initRandomDirectionWithRespectToCurrentK(const std::array<T, 3> &k, T cosTheta, T rand) {
  const T kxy = std::sqrt(k[0] * k[0] + k[1] * k[1]);
  const T kn = std::sqrt(kxy * kxy + k[2] * k[2]);
  if (kn == T(0))
    return {T(0), T(0), T(0)};
  const T ct0 = k[2] / kn, st0 = kxy / kn;
  const T cf0 = kxy > T(0) ? k[0] / kxy : T(1);
  const T sf0 = kxy > T(0) ? k[1] / kxy : T(0);
  const T st = std::sqrt(1.0 - cosTheta * cosTheta);
  const T phi = 2.0 * emcPi<T>() * rand;
  const T xp = kn * st * std::cos(phi), yp = kn * st * std::sin(phi), zp = kn * cosTheta;
  return {xp * cf0 * ct0 - yp * sf0 + zp * cf0 * st0, xp * sf0 * ct0 + yp * cf0 + zp * sf0 * st0,
          -xp * st0 + zp * ct0};
}

#endif
