// In-register complex FFT building blocks (sizes 2..32), fully unrolled at compile time.
//
// Every twiddle inside a register FFT is a compile-time constant, so the DIT butterfly
//     out+ = a + w*b ,  out- = 2a - out+
// compiles to 6 FFMA whose multiplier is an immediate (full FP32 issue rate on sm_100a; a
// three-register FFMA is register-port limited to 2/3 rate — profiles/r01_ubench2_pipes.txt).
// The header is host/device so tests can run the identical template code on the CPU.
#pragma once
#if defined(__CUDACC__)
#define BF_HD __host__ __device__ __forceinline__
#else
#define BF_HD inline
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace bf {

// constexpr sin/cos(2*pi*k/n) for n a power of two <= 64, evaluated by the host compiler.
// Table of cos(2*pi*k/64), k=0..16 (first quadrant, 20 significant digits).
constexpr double kCos64[17] = {1.0,
                               0.99518472667219688624,
                               0.98078528040323044913,
                               0.95694033573220886494,
                               0.92387953251128675613,
                               0.88192126434835502971,
                               0.83146961230254523708,
                               0.77301045336273696081,
                               0.70710678118654752440,
                               0.63439328416364549822,
                               0.55557023301960222474,
                               0.47139673682599764856,
                               0.38268343236508977173,
                               0.29028467725446236764,
                               0.19509032201612826785,
                               0.09801714032956060199,
                               0.0};
constexpr double cos64(int k) {   // cos(2*pi*k/64), any integer k
  k = ((k % 64) + 64) % 64;
  return k <= 16 ? kCos64[k] : k <= 32 ? -kCos64[32 - k] : k <= 48 ? -kCos64[k - 32] : kCos64[64 - k];
}
constexpr double sin64(int k) { return cos64(k - 16); }

BF_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
BF_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
BF_HD float2 cmul(float2 a, float2 b) { return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y)); }
BF_HD float2 cmulc(float2 a, float2 b) {   // a * conj(b)
  return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(a.y, b.x, -a.x * b.y));
}

// One DIT butterfly on (a, b) with constant twiddle W = exp(DIR * 2*pi*i * K64/64).
// DIR = -1: forward (e^{-i...}), +1: backward.
template <int K64, int DIR>
BF_HD void bfly(float2& a, float2& b) {
  constexpr int k = ((K64 % 64) + 64) % 64;
  if constexpr (k == 0) {
    float2 t = b;
    b = csub(a, t);
    a = cadd(a, t);
  } else if constexpr (k == 16) {   // W = DIR * i  ->  w*b = DIR*(-b.y, b.x)
    float2 t = (DIR > 0) ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
    b = csub(a, t);
    a = cadd(a, t);
  } else if constexpr (k == 32) {
    float2 t = b;
    b = cadd(a, t);
    a = csub(a, t);
  } else if constexpr (k == 48) {
    float2 t = (DIR > 0) ? make_float2(b.y, -b.x) : make_float2(-b.y, b.x);
    b = csub(a, t);
    a = cadd(a, t);
  } else {
    constexpr float wr = (float)cos64(k);
    constexpr float wi = (float)(DIR * sin64(k));
    float pr = fmaf(b.x, wr, a.x);
    float pi = fmaf(b.x, wi, a.y);
    pr = fmaf(-b.y, wi, pr);
    pi = fmaf(b.y, wr, pi);
    b = make_float2(fmaf(a.x, 2.0f, -pr), fmaf(a.y, 2.0f, -pi));
    a = make_float2(pr, pi);
  }
}

// bit reversal of i over LOG bits
constexpr int brev(int i, int LOG) {
  int r = 0;
  for (int b = 0; b < LOG; b++)
    if (i & (1 << b)) r |= 1 << (LOG - 1 - b);
  return r;
}
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

template <int R, int DIR, int LEN, int S, int K>
struct StageK {
  // butterflies k = K.. of one DIT stage with span LEN inside block starting at S
  static BF_HD void run(float2* v) {
    if constexpr (K < LEN / 2) {
      bfly<(64 / LEN) * K, DIR>(v[S + K], v[S + K + LEN / 2]);
      StageK<R, DIR, LEN, S, K + 1>::run(v);
    }
  }
};
template <int R, int DIR, int LEN, int S>
struct StageS {
  static BF_HD void run(float2* v) {
    if constexpr (S < R) {
      StageK<R, DIR, LEN, S, 0>::run(v);
      StageS<R, DIR, LEN, S + LEN>::run(v);
    }
  }
};
template <int R, int DIR, int LEN>
struct Stages {
  static BF_HD void run(float2* v) {
    if constexpr (LEN <= R) {
      StageS<R, DIR, LEN, 0>::run(v);
      Stages<R, DIR, LEN * 2>::run(v);
    }
  }
};

// In-place radix-2 DIT FFT of size R over v[0..R).  INPUT must be supplied in bit-reversed order
// (v[brev(n)] = x[n]); OUTPUT is in natural order (v[k] = X[k]).  Callers do the bit reversal for
// free by choosing which register each loaded sample lands in (all indices are compile-time).
template <int R, int DIR>
BF_HD void fft_dit(float2* v) {
  static_assert(R >= 2 && R <= 64 && (R & (R - 1)) == 0, "size");
  Stages<R, DIR, 2>::run(v);
}

}   // namespace bf
