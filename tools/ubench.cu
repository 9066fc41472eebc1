// Pipe-throughput microbenchmark for B200 (sm_100a): decides the FFT kernel's
// instruction mix (scalar FP32 vs packed f32x2, shuffle vs shared-memory exchange).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Output: warp-instructions per clock per SM for each op class.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 4096

enum Op { FADD, FMUL, FFMA, FFMA_IMM, FADD2, FMUL2, FFMA2, DFMA, SHFL, LDS32, LDS64, LDS128, STS64, MUFU_RSQ, NOPS };
static const char* names[] = {"FADD", "FMUL", "FFMA(3reg)", "FFMA(imm)", "FADD2", "FMUL2", "FFMA2", "DFMA", "SHFL", "LDS.32", "LDS.64", "LDS.128", "STS.64", "MUFU.RSQ"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float seed) {
  extern __shared__ float4 sm4[];
  float* sm = (float*)sm4;
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[8], b[8], c[8];
  float2 p[8], q[8], r[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = seed + i + tid; b[i] = seed * 0.5f + i; c[i] = seed * 0.25f - i;
    p[i] = make_float2(a[i], b[i]); q[i] = make_float2(b[i], c[i]); r[i] = make_float2(c[i], a[i]);
    d[i] = a[i];
  }
  float4 v4 = make_float4(0, 0, 0, 0);
  float2 v2 = make_float2(0, 0);
  int idx = tid;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == FADD) a[i] = a[i] + b[i];
      if (OP == FMUL) a[i] = a[i] * b[i];
      if (OP == FFMA) a[i] = fmaf(a[i], b[i], c[i]);
      if (OP == FFMA_IMM) a[i] = fmaf(a[i], 1.0001f, b[i]);
      if (OP == FADD2) p[i] = __fadd2_rn(p[i], q[i]);
      if (OP == FMUL2) p[i] = __fmul2_rn(p[i], q[i]);
      if (OP == FFMA2) p[i] = __ffma2_rn(p[i], q[i], r[i]);
      if (OP == DFMA) d[i] = fma(d[i], 1.0000001, 0.5);
      if (OP == SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + i);
      if (OP == LDS32) { a[i] += sm[(idx + i * 32) & 8191]; }
      if (OP == LDS64) { float2 t = ((float2*)sm)[(idx + i * 32) & 4095]; v2.x += t.x; v2.y += t.y; }
      if (OP == LDS128) { float4 t = sm4[(idx + i * 32) & 2047]; v4.x += t.x; v4.y += t.y; v4.z += t.z; v4.w += t.w; }
      if (OP == STS64) { ((float2*)sm)[(idx + i * 32) & 4095] = make_float2(a[i], b[i]); }
      if (OP == MUFU_RSQ) a[i] = rsqrtf(a[i]);
    }
    if (OP == LDS32 || OP == LDS64 || OP == LDS128) idx = (idx + 7) & 8191;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i] + p[i].x + p[i].y + (float)d[i];
  s += v4.x + v4.y + v4.z + v4.w + v2.x + v2.y + sm[tid];
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int nsm, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * nsm * threads);
  cudaMalloc(&cyc, sizeof(long long) * nsm);
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  k<OP><<<nsm, threads, 32768>>>(out, cyc, 1.0f);
  k<OP><<<nsm, threads, 32768>>>(out, cyc, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", names[OP], cudaGetErrorString(e)); return; }
  long long* h = (long long*)malloc(sizeof(long long) * nsm);
  cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
  double winstr = (double)ITERS * 8 * (threads / 32);
  // LDS variants carry one extra FADD per loaded word; report the primary op only.
  printf("%-12s threads=%4d  cycles=%9.0f  warp-instr/clk/SM=%6.3f  (lane-ops/clk/SM=%7.1f)\n", names[OP], threads, avg,
         winstr / avg, winstr * 32 / avg);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int nsm = pr.multiProcessorCount;
  printf("device %s SMs=%d clock=%d kHz\n", pr.name, nsm, pr.clockRate);
  for (int threads : {256, 512, 1024}) {
    run<FADD>(nsm, threads); run<FMUL>(nsm, threads); run<FFMA>(nsm, threads); run<FFMA_IMM>(nsm, threads);
    run<FADD2>(nsm, threads); run<FMUL2>(nsm, threads); run<FFMA2>(nsm, threads); run<DFMA>(nsm, threads);
    run<SHFL>(nsm, threads); run<LDS32>(nsm, threads); run<LDS64>(nsm, threads); run<LDS128>(nsm, threads);
    run<STS64>(nsm, threads); run<MUFU_RSQ>(nsm, threads);
  }
  return 0;
}
