// TEST INFRASTRUCTURE: the slice of the FFTW3 API the reference nodes call (das.cpp:122-128,53,66):
// fftw_malloc, fftw_plan_dft_1d (c2c, double), fftw_execute.  FFTW's published definition is restated:
// unnormalised Y[k] = sum_n X[n] e^{sign * 2 pi i n k / N}, FFTW_FORWARD = -1, FFTW_BACKWARD = +1.
// The transform is a Stockham autosort radix-2 FFT with long-double twiddles (deliberately NOT the
// algorithm of oracle/bf_oracle.hpp, so the two check each other); non-power-of-two N falls back to the
// O(N^2) definition.  Like FFTW_MEASURE planning, creating a plan clobbers (zero-fills) in/out.
#pragma once
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef double fftw_complex[2];
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
struct bfshim_fftw_plan_s {
  int n, sign;
  std::complex<double>*in, *out;
  std::vector<std::complex<double> > tw, scratch, work;
};
typedef bfshim_fftw_plan_s* fftw_plan;
static inline void* fftw_malloc(size_t n) { return calloc(n + 64, 1); }   // +64: phasempf.cpp:274 writes y_fft[fft_win]
static inline void fftw_free(void* p) { free(p); }
static inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned) {
  fftw_plan p = new bfshim_fftw_plan_s();
  p->n = n; p->sign = sign;
  p->in = reinterpret_cast<std::complex<double>*>(in);
  p->out = reinterpret_cast<std::complex<double>*>(out);
  p->tw.resize(n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; k++) {
    long double a = (long double)sign * two_pi * (long double)k / (long double)n;
    p->tw[k] = std::complex<double>((double)cosl(a), (double)sinl(a));
  }
  p->scratch.resize(n);
  memset((void*)in, 0, sizeof(fftw_complex) * n);
  memset((void*)out, 0, sizeof(fftw_complex) * n);
  return p;
}
static inline void fftw_execute(const fftw_plan p) {
  const int n = p->n;
  typedef std::complex<double> cd;
  if (n & (n - 1)) {
    for (int k = 0; k < n; k++) {
      cd acc(0, 0);
      for (int j = 0; j < n; j++) acc += p->in[j] * p->tw[(int)(((long long)j * k) % n)];
      p->scratch[k] = acc;
    }
    for (int k = 0; k < n; k++) p->out[k] = p->scratch[k];
    return;
  }
  // Stockham: ping-pong between two plan-owned buffers (no allocation per call); stage with l butterfly groups of stride m
  if ((int)p->work.size() != n) p->work.resize(n);
  for (int k = 0; k < n; k++) p->work[k] = p->in[k];
  cd* x = p->work.data();
  cd* y = p->scratch.data();
  for (int l = n / 2, m = 1; l >= 1; l >>= 1, m <<= 1) {
    for (int j = 0; j < l; j++) {
      const cd w = p->tw[j * m];
      for (int k = 0; k < m; k++) {
        const cd c0 = x[k + j * m], c1 = x[k + j * m + l * m];
        y[k + 2 * j * m] = c0 + c1;
        y[k + 2 * j * m + m] = w * (c0 - c1);
      }
    }
    cd* t = x; x = y; y = t;
  }
  for (int k = 0; k < n; k++) p->out[k] = x[k];
}
static inline void fftw_destroy_plan(fftw_plan p) { delete p; }
