// Microbenchmark: FP32 scalar vs packed f32x2 pipe throughput on sm_100a, plus LDS co-issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_pipe fp_pipe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
typedef unsigned long long u64;
constexpr int NACC = 16;
constexpr int ITER = 4096;

template<int MODE>
__global__ void __launch_bounds__(512) kern(float* out, long long* cyc, float seed) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[NACC]; u64 p[NACC];
  float b = seed + 1.0f, c = seed * 0.5f;
  u64 B, C;
  asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(c), "f"(c));
#pragma unroll
  for (int i = 0; i < NACC; i++) { a[i] = seed * (threadIdx.x + i); asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.f)); }
  long long t0 = clock64();
  const float2* smp = reinterpret_cast<const float2*>(sm) + threadIdx.x % 32;
  float2 acc2 = make_float2(0, 0);
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(B), "l"(C));
      if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(B));
      if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(B));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, 0f3F800011, %1;" : "+f"(a[i]) : "f"(c));   // imm multiplier
      if (MODE == 6) { // FFMA2 + one LDS.64 per 4 FFMA2
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(B), "l"(C));
        if ((i & 3) == 0) { float2 v = smp[((it * 4 + (i >> 2)) & 63) * 32]; acc2.x += v.x; acc2.y += v.y; }
      }
      if (MODE == 7) { // scalar FFMA + one LDS.32 per 4 FFMA
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));
        if ((i & 3) == 0) { float v = sm[((it * 4 + (i >> 2)) & 127) * 32 + threadIdx.x % 32]; acc2.x += v; }
      }
      if (MODE == 8) { // alternate FFMA2 and scalar FADD (mixed)
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(B), "l"(C));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      }
    }
  }
  long long t1 = clock64();
  float s = acc2.x + acc2.y;
#pragma unroll
  for (int i = 0; i < NACC; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE> void run(const char* name, double lane_ops_per_instr, int threads, int blocks_per_sm) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int blocks = nsm * blocks_per_sm;
  float* out; long long* cyc; CK(cudaMalloc(&out, sizeof(float) * blocks * threads)); CK(cudaMalloc(&cyc, 8 * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) kern<MODE><<<blocks, threads>>>(out, cyc, 1e-9f);
  CK(cudaEventRecord(e0));
  const int reps = 20;
  for (int w = 0; w < reps; w++) kern<MODE><<<blocks, threads>>>(out, cyc, 1e-9f);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  long long* h = (long long*)malloc(8 * blocks); CK(cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  double instr = (double)ITER * NACC * threads * blocks_per_sm;  // thread-instr per SM (of the main op)
  printf("%-28s thr=%4d bps=%d  ms=%.4f  cyc/SM=%.0f  eff_clk=%.0f MHz  main-op thread-instr/clk/SM=%.1f  lane-flops(ops)/clk/SM=%.1f\n",
         name, threads, blocks_per_sm, ms, avg, avg / (ms * 1e3), instr / avg, instr * lane_ops_per_instr / avg);
  cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
  for (int thr : {256, 512, 1024}) {
    int bps = 1024 / thr; if (bps < 1) bps = 1;
    run<0>("FFMA scalar", 1, thr, bps);
    run<1>("FFMA2 packed", 2, thr, bps);
    run<2>("FADD scalar", 1, thr, bps);
    run<3>("FADD2 packed", 2, thr, bps);
    run<4>("FMUL2 packed", 2, thr, bps);
    run<5>("FFMA imm", 1, thr, bps);
    run<6>("FFMA2 + LDS.64/4", 2, thr, bps);
    run<7>("FFMA + LDS.32/4", 1, thr, bps);
    run<8>("FFMA2 + FADD alt", 3, thr, bps);
  }
  // long sustained run for clocks under FP32 load
  for (int r = 0; r < 3; r++) run<1>("FFMA2 sustained", 2, 512, 2);
  return 0;
}
