// Double-precision twin of fft_reg.cuh (generated from it by type substitution: float2 -> double2, names get a _d
// suffix).  Used where a decision must be taken from spectra that agree with the reference's complex-double FFT
// (phase masks of phase.cpp / phasempf.cpp); B200 issues DFMA at half the FFMA rate.
#pragma once
#include "fft_reg.cuh"

namespace bf {

BF_HD double2 cadd_d(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
BF_HD double2 csub_d(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
BF_HD double2 cmul_d(double2 a, double2 b) { return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y)); }
BF_HD double2 cmulc_d(double2 a, double2 b) {   // a * conj(b)
  return make_double2(fma(a.y, b.y, a.x * b.x), fma(a.y, b.x, -a.x * b.y));
}

// One DIT butterfly on (a, b) with constant twiddle W = exp(DIR * 2*pi*i * K64/64).
// DIR = -1: forward (e^{-i...}), +1: backward.
template <int K64, int DIR>
BF_HD void bfly_d(double2& a, double2& b) {
  constexpr int k = ((K64 % 64) + 64) % 64;
  if constexpr (k == 0) {
    double2 t = b;
    b = csub_d(a, t);
    a = cadd_d(a, t);
  } else if constexpr (k == 16) {   // W = DIR * i  ->  w*b = DIR*(-b.y, b.x)
    double2 t = (DIR > 0) ? make_double2(-b.y, b.x) : make_double2(b.y, -b.x);
    b = csub_d(a, t);
    a = cadd_d(a, t);
  } else if constexpr (k == 32) {
    double2 t = b;
    b = cadd_d(a, t);
    a = csub_d(a, t);
  } else if constexpr (k == 48) {
    double2 t = (DIR > 0) ? make_double2(b.y, -b.x) : make_double2(-b.y, b.x);
    b = csub_d(a, t);
    a = cadd_d(a, t);
  } else {
    constexpr double wr = (double)cos64(k);
    constexpr double wi = (double)(DIR * sin64(k));
    double pr = fma(b.x, wr, a.x);
    double pi = fma(b.x, wi, a.y);
    pr = fma(-b.y, wi, pr);
    pi = fma(b.y, wr, pi);
    b = make_double2(fma(a.x, 2.0, -pr), fma(a.y, 2.0, -pi));
    a = make_double2(pr, pi);
  }
}


template <int R, int DIR, int LEN, int S, int K>
struct StageK_d {
  // butterflies k = K.. of one DIT stage with span LEN inside block starting at S
  static BF_HD void run(double2* v) {
    if constexpr (K < LEN / 2) {
      bfly_d<(64 / LEN) * K, DIR>(v[S + K], v[S + K + LEN / 2]);
      StageK_d<R, DIR, LEN, S, K + 1>::run(v);
    }
  }
};
template <int R, int DIR, int LEN, int S>
struct StageS_d {
  static BF_HD void run(double2* v) {
    if constexpr (S < R) {
      StageK_d<R, DIR, LEN, S, 0>::run(v);
      StageS_d<R, DIR, LEN, S + LEN>::run(v);
    }
  }
};
template <int R, int DIR, int LEN>
struct Stages_d {
  static BF_HD void run(double2* v) {
    if constexpr (LEN <= R) {
      StageS_d<R, DIR, LEN, 0>::run(v);
      Stages_d<R, DIR, LEN * 2>::run(v);
    }
  }
};

// In-place radix-2 DIT FFT of size R over v[0..R).  INPUT must be supplied in bit-reversed order
// (v[brev(n)] = x[n]); OUTPUT is in natural order (v[k] = X[k]).  Callers do the bit reversal for
// free by choosing which register each loaded sample lands in (all indices are compile-time).
template <int R, int DIR>
BF_HD void fft_dit_d(double2* v) {
  static_assert(R >= 2 && R <= 64 && (R & (R - 1)) == 0, "size");
  Stages_d<R, DIR, 2>::run(v);
}

}   // namespace bf
