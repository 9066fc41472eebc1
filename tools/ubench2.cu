// Second-pass pipe microbenchmark: low loop overhead (64 ops / iteration), operand-reuse
// variants of FFMA / FFMA2, SHFL + LDS concurrency, complex multiply in scalar vs f32x2 form.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 1024
#define NACC 16
#define REP 4

enum Op { FADD, FFMA_D, FFMA_R, FADD2, FFMA2_D, FFMA2_R, FFMA2_ACC, CMUL_S, CMUL_P, SHFL_LDS, SHFL_ONLY, LDS_ONLY, FADD_LDS, NOPS };
static const char* names[] = {"FADD", "FFMA distinct", "FFMA a=a*b+c (b,c shared)", "FADD2", "FFMA2 distinct", "FFMA2 (b,c shared)",
                              "FFMA2 p=q*r+p", "CMUL scalar (4 instr)", "CMUL packed (2 instr)", "SHFL+LDS.128 mix", "SHFL only", "LDS.128 only", "FADD+LDS.128 mix"};

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed) {
  extern __shared__ float4 sm4[];
  float* sm = (float*)sm4;
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[NACC], b[NACC], c[NACC];
  float2 p[NACC], q[NACC], r[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    a[i] = seed + i + tid; b[i] = seed * 0.5f + i; c[i] = seed * 0.25f - i;
    p[i] = make_float2(a[i], b[i]); q[i] = make_float2(b[i] * 1.5f, c[i]); r[i] = make_float2(c[i] + 3.f, a[i] - 7.f);
  }
  float4 v4 = make_float4(0, 0, 0, 0);
  const float bs = seed * 1.25f, cs = seed * 0.75f;
  const float2 qs = make_float2(bs, cs), rs = make_float2(cs, bs);
  int idx = tid;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int rep = 0; rep < REP; rep++) {
#pragma unroll
      for (int i = 0; i < NACC; i++) {
        if (OP == FADD) a[i] = a[i] + b[i];
        if (OP == FFMA_D) a[i] = fmaf(a[i], b[i], c[i]);
        if (OP == FFMA_R) a[i] = fmaf(a[i], bs, cs);
        if (OP == FADD2) p[i] = __fadd2_rn(p[i], q[i]);
        if (OP == FFMA2_D) p[i] = __ffma2_rn(p[i], q[i], r[i]);
        if (OP == FFMA2_R) p[i] = __ffma2_rn(p[i], qs, rs);
        if (OP == FFMA2_ACC) p[i] = __ffma2_rn(q[i], r[i], p[i]);
        if (OP == CMUL_S) {   // p *= q (complex), scalar: 2 FMUL + 2 FFMA
          float re = p[i].x * q[i].x; float im = p[i].x * q[i].y;
          re = fmaf(-p[i].y, q[i].y, re); im = fmaf(p[i].y, q[i].x, im);
          p[i] = make_float2(re, im);
        }
        if (OP == CMUL_P) {   // p *= q (complex), packed: FMUL2 + FFMA2 with operand half-swizzles
          float2 t = __fmul2_rn(make_float2(p[i].y, p[i].y), make_float2(-q[i].y, q[i].x));
          p[i] = __ffma2_rn(make_float2(p[i].x, p[i].x), q[i], t);
        }
        if (OP == SHFL_LDS) {
          if (i & 1) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
          else { float4 t = sm4[(idx + i * 32) & 2047]; v4.x += t.x; v4.y += t.y; v4.z += t.z; v4.w += t.w; }
        }
        if (OP == SHFL_ONLY) { if (i & 1) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15)); }
        if (OP == LDS_ONLY) { if (!(i & 1)) { float4 t = sm4[(idx + i * 32) & 2047]; v4.x += t.x; v4.y += t.y; v4.z += t.z; v4.w += t.w; } }
        if (OP == FADD_LDS) {
          if (i == 0) { float4 t = sm4[(idx + i * 32) & 2047]; v4.x += t.x; v4.y += t.y; v4.z += t.z; v4.w += t.w; }
          else a[i] = a[i] + b[i];
        }
      }
    }
    if (OP == SHFL_LDS || OP == LDS_ONLY || OP == FADD_LDS) idx = (idx + 7) & 2047;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += a[i] + p[i].x + p[i].y;
  s += v4.x + v4.y + v4.z + v4.w + sm[tid];
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
  double groups = (double)ITERS * REP * NACC * (threads / 32);   // "op groups" per SM
  printf("%-28s threads=%4d cycles=%9.0f  clk per op-group per SM = %6.3f\n", names[OP], threads, avg, avg / groups);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int nsm = pr.multiProcessorCount;
  printf("device %s SMs=%d. op-group = one warp executing the listed op once (CMUL = one complex multiply per lane;\n"
         "SHFL+LDS mix: half the groups are SHFL, half LDS.128; *_only run just their half)\n", pr.name, nsm);
  for (int threads : {256, 512}) {
    run<FADD>(nsm, threads); run<FFMA_D>(nsm, threads); run<FFMA_R>(nsm, threads); run<FADD2>(nsm, threads);
    run<FFMA2_D>(nsm, threads); run<FFMA2_R>(nsm, threads); run<FFMA2_ACC>(nsm, threads);
    run<CMUL_S>(nsm, threads); run<CMUL_P>(nsm, threads);
    run<SHFL_LDS>(nsm, threads); run<SHFL_ONLY>(nsm, threads); run<LDS_ONLY>(nsm, threads); run<FADD_LDS>(nsm, threads);
  }
  return 0;
}
