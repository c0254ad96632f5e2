// TMA probe 2: 2-D (sample, clip) tensor, one 164-sample box per copy at an arbitrary sample offset.
// Checks the shared-memory destination alignment the hardware accepts and times N row copies per "tile".
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#ifndef KSHIFT
#define KSHIFT 0
#endif
#ifndef XOFF
#define XOFF 10036
#endif
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
__device__ __forceinline__ unsigned s32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nLW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD_%=;\nbra LW_%=;\nLD_%=:\n}" :: "r"(s32(bar)), "r"(parity) : "memory");
}
__global__ void k1(const __grid_constant__ CUtensorMap tm, int x, int clip, int dst_off_words, float* out) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar))); asm volatile("fence.proxy.async.shared::cta;"); }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(656));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(s32(sm + dst_off_words)), "l"(&tm), "r"(x), "r"(clip), "r"(s32(&bar)) : "memory");
  }
  wait(&bar, 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = sm[i];
}
// persistent timing kernel: each CTA repeats `iters` tiles of N row copies (pitch-165 layout), lanes issue in parallel
__global__ void __launch_bounds__(512, 1) k2(const __grid_constant__ CUtensorMap tm, int iters, int ncopies, long long* cyc, float* sink) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ unsigned long long bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar))); asm volatile("fence.proxy.async.shared::cta;"); }
  __syncthreads();
  long long t0 = clock64();
  unsigned parity = 0; float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int f0 = 64 + 64 * ((it + blockIdx.x) % 40);
    const int clip = blockIdx.x % 4;
    if (warp == 13) {
      if (lane == 0) { asm volatile("fence.proxy.async.shared::cta;"); asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(656 * ncopies)); }
      __syncwarp();
      for (int r = lane; r < ncopies; r += 32) {
        const int s = r & 3;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(s32(sm + 165 * r - s)), "l"(&tm), "r"(160 * f0 - 200 + 160 * r - KSHIFT * s), "r"(clip), "r"(s32(&bar)) : "memory");
      }
    }
    wait(&bar, parity); parity ^= 1;
    acc += sm[(tid * 21) % (165 * 66)];
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 512 + tid] = acc;
}
int main() {
  const long long stride = 480000; const int batch = 4;
  float* d; CK(cudaMalloc(&d, sizeof(float) * stride * batch));
  float* h = (float*)malloc(sizeof(float) * stride * batch);
  for (long long i = 0; i < stride * batch; i++) h[i] = (float)(i % 1000003);
  CK(cudaMemcpy(d, h, sizeof(float) * stride * batch, cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out, 1 << 20)); float* ho = (float*)malloc(4096);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)batch}; cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
  cuuint32_t box[2] = {164, 1}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* cyc; CK(cudaMalloc(&cyc, 8 * 148)); long long hc[148];
  const int offs[] = {0, 32, 16, 8, 4};
  for (int off : offs) {
    k1<<<1, 128, 8192>>>(tm, XOFF, 2, off, out);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("dst offset %3d words: RUN FAILED: %s\n", off, cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost));
    int bad = 0; for (int i = 0; i < 164; i++) if (ho[off + i] != h[2 * stride + XOFF + i]) bad++;
    printf("dst offset %3d words (%4d B): ok, mismatches %d/164, before=%g after=%g\n", off, off * 4, bad, off ? ho[off - 1] : -1.f, ho[off + 164]);
  }
  for (int nc : {66, 80, 33}) {
    const int iters = 200;
    k2<<<148, 512, 200 * 1024>>>(tm, iters, nc, cyc, out);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k2 RUN FAILED: %s\n", cudaGetErrorString(e)); return 3; }
    CK(cudaMemcpy(hc, cyc, 8 * 148, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; i++) avg += hc[i]; avg /= 148;
    printf("k2: %d copies/tile, all 148 SMs: %.0f cycles per tile (issue + wait, nothing else running)\n", nc, avg / iters);
  }
  return 0;
}
