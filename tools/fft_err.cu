// Measures the absolute error of the FP32 shared-memory transform of the CTA-per-stream kernels against a double DFT:
//   max_k |Z32[k] - Z[k]| / ||z||_2   over random and coloured inputs, per frame size.
// The phase-mask kernel (phase_n_kernel.cu) bounds the error of a transform output by kFftErr ||z||_2 + kFftErrMax max_k |Z[k]|_1 + kFftErrRel |Z[k]|;
// the second line printed per frame size is the largest observed fraction of that bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I beamform_b200/csrc -o /tmp/fft_err tools/fft_err.cu && /tmp/fft_err
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "block_fft.cuh"

using namespace bf;

template <int NN>
__global__ void __launch_bounds__(kGenThreads, 2) fft_kernel(const float2* in, float2* out, const float2* tw) {
  float2* z = reinterpret_cast<float2*>(gen_smem_raw);
  const float2* src = in + (size_t)blockIdx.x * NN;
  for (int n = threadIdx.x; n < NN; n += kGenThreads) z[swz(n)] = src[n];
  __syncthreads();
  block_fft_fn<NN, -1, float2>(0u, 1, tw, threadIdx.x);
  for (int n = threadIdx.x; n < NN; n += kGenThreads) out[(size_t)blockIdx.x * NN + n] = z[swz(n)];
}

template <int NN>
static void run(int trials) {
  std::vector<float2> tw(NN), in((size_t)trials * NN), out((size_t)trials * NN);
  for (int k = 0; k < NN; k++) tw[k] = make_float2((float)cos(-2.0 * M_PI * k / NN), (float)sin(-2.0 * M_PI * k / NN));
  srand(1234 + NN);
  auto rnd = [] { return (double)rand() / RAND_MAX * 2.0 - 1.0; };
  for (int t = 0; t < trials; t++) {
    const int kind = t % 4;
    const double f0 = 3.0 + 40.0 * rnd() * rnd(), f1 = 11.0 + 70.0 * fabs(rnd());
    for (int n = 0; n < NN; n++) {
      const double w = 0.5 * sin(M_PI * n / NN);
      double a, b;
      if (kind == 0) { a = rnd(); b = rnd(); }                                        // white, both frames
      else if (kind == 1) { a = sin(2 * M_PI * f0 * n / NN) + 1e-3 * rnd(); b = 0.5 * sin(2 * M_PI * f1 * n / NN + 1.0) + 1e-3 * rnd(); }   // tones + floor
      else if (kind == 2) { a = 1e-4 * rnd(); b = sin(2 * M_PI * f0 * n / NN) + 0.3 * rnd(); }   // quiet frame packed with a loud one
      else { a = (n % 97 == 0) ? 1.0 : 1e-3 * rnd(); b = (n % 89 == 0) ? -1.0 : 0.0; }           // impulsive
      in[(size_t)t * NN + n] = make_float2((float)(a * w), (float)(b * w));
    }
  }
  float2 *d_in, *d_out, *d_tw;
  cudaMalloc(&d_in, sizeof(float2) * in.size()); cudaMalloc(&d_out, sizeof(float2) * in.size()); cudaMalloc(&d_tw, sizeof(float2) * NN);
  cudaMemcpy(d_in, in.data(), sizeof(float2) * in.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_tw, tw.data(), sizeof(float2) * NN, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(fft_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float2) * NN));
  fft_kernel<NN><<<trials, kGenThreads, sizeof(float2) * NN>>>(d_in, d_out, d_tw);
  cudaMemcpy(out.data(), d_out, sizeof(float2) * in.size(), cudaMemcpyDeviceToHost);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("N=%d: CUDA error\n", NN); return; }
  std::vector<double> c(NN), sn(NN);
  for (int k = 0; k < NN; k++) { c[k] = cos(2.0 * M_PI * k / NN); sn[k] = -sin(2.0 * M_PI * k / NN); }
  const double ka = 1.0e-6, km = 1.5e-7, kb = 4.0e-7;   // kFftErr, kFftErrMax, kFftErrRel of phase_n_kernel.cu
  double worst[4] = {0, 0, 0, 0}, rms[4] = {0, 0, 0, 0}, util[4] = {0, 0, 0, 0};
  for (int t = 0; t < trials; t++) {
    double nrm = 0.0;
    for (int n = 0; n < NN; n++) { const float2 v = in[(size_t)t * NN + n]; nrm += (double)v.x * v.x + (double)v.y * v.y; }
    nrm = sqrt(nrm);
    double mx = 0.0, acc = 0.0, zmax = 0.0;
    std::vector<double> zr(NN), zi(NN);
    for (int k = 0; k < NN; k++) {
      double re = 0.0, im = 0.0;
      for (int n = 0; n < NN; n++) {
        const float2 v = in[(size_t)t * NN + n];
        const int q = (int)(((long long)k * n) & (NN - 1));
        re += v.x * c[q] - v.y * sn[q];
        im += v.x * sn[q] + v.y * c[q];
      }
      zr[k] = re; zi[k] = im;
      zmax = fmax(zmax, fabs(re) + fabs(im));
    }
    for (int k = 0; k < NN; k++) {
      const double re = zr[k], im = zi[k];
      const float2 o = out[(size_t)t * NN + k];
      const double e = hypot(o.x - re, o.y - im);
      mx = fmax(mx, e); acc += e * e;
      util[t % 4] = fmax(util[t % 4], e / (ka * nrm + km * zmax + kb * hypot(re, im)));
    }
    worst[t % 4] = fmax(worst[t % 4], mx / nrm);
    rms[t % 4] = fmax(rms[t % 4], sqrt(acc / NN) / nrm);
  }
  printf("N=%4d  %3d trials  max|err|/||z||: white %.3g  tones %.3g  quiet+loud %.3g  impulsive %.3g   (rms: %.3g %.3g %.3g %.3g)\n", NN, trials,
         worst[0], worst[1], worst[2], worst[3], rms[0], rms[1], rms[2], rms[3]);
  printf("        max |err[k]| / (%.1e ||z|| + %.1e max|Z|_1 + %.1e |Z[k]|) (must stay well below 1): white %.3f  tones %.3f  quiet+loud %.3f  impulsive %.3f\n", ka, km, kb,
         util[0], util[1], util[2], util[3]);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_tw);
}

int main() {
  run<512>(64);
  run<1024>(64);
  run<2048>(48);
  run<4096>(32);
  return 0;
}
