// Minimal TMA probe: which tensor-map shapes does cp.async.bulk.tensor accept on sm_100a?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
__device__ __forceinline__ unsigned s32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
template<int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int bytes, float* out, int nout) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)));
    asm volatile("fence.proxy.async.shared::cta;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(bytes));
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   :: "r"(s32(sm)), "l"(&tm), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(s32(sm)), "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nL1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra L2;\nbra L1;\nL2:\n}" :: "r"(s32(&bar)));
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = sm[i];
}
int main() {
  const long long stride = 480000; const int batch = 4;
  float* d; CK(cudaMalloc(&d, sizeof(float) * stride * batch));
  float* h = (float*)malloc(sizeof(float) * stride * batch);
  for (long long i = 0; i < stride * batch; i++) h[i] = (float)(i % 1000003);
  CK(cudaMemcpy(d, h, sizeof(float) * stride * batch, cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out, 65536)); float* ho = (float*)malloc(65536);
  struct Cfg { const char* name; int rank; cuuint64_t dims[3]; cuuint64_t strides[2]; cuuint32_t box[3]; int x, y, z; };
  Cfg cfgs[] = {
    {"2d plain 160x3000 box 160x8", 2, {160, 3000, 1}, {640, 0}, {160, 8, 1}, 0, 5, 0},
    {"2d plain box 164 > dim 160", 2, {160, 3000, 1}, {640, 0}, {164, 8, 1}, 0, 5, 0},
    {"2d overlap X=284 ys=160 box 164x10", 2, {284, 2999, 1}, {640, 0}, {164, 10, 1}, 116, 62, 0},
    {"3d plain 160x3000xB box 160x8x1", 3, {160, 3000, 4}, {640, 1920000}, {160, 8, 1}, 0, 5, 1},
    {"3d overlap X=284 box 164x10x1", 3, {284, 2999, 4}, {640, 1920000}, {164, 10, 1}, 116, 62, 1},
  };
  for (auto& c : cfgs) {
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.rank, d, c.dims, c.strides, c.box, es,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", c.name, (int)r); continue; }
    int n = c.box[0] * c.box[1] * c.box[2];
    CK(cudaMemset(out, 0, 65536));
    if (c.rank == 2) k<2><<<1, 128, 32768>>>(tm, c.x, c.y, c.z, n * 4, out, n);
    else k<3><<<1, 128, 32768>>>(tm, c.x, c.y, c.z, n * 4, out, n);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s RUN FAILED: %s\n", c.name, cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(ho, out, n * 4, cudaMemcpyDeviceToHost));
    // verify
    int bad = 0;
    for (unsigned j = 0; j < c.box[1]; j++) for (unsigned i = 0; i < c.box[0]; i++) {
      long long xi = c.x + i, yi = c.y + j;
      float exp = (xi < (long long)c.dims[0] && yi < (long long)c.dims[1]) ? h[c.z * (c.rank == 3 ? c.strides[1] / 4 : 0) + yi * (c.strides[0] / 4) + xi] : 0.f;
      if (ho[j * c.box[0] + i] != exp) bad++;
    }
    printf("%-40s ok, mismatches %d / %d   got %g %g %g ... exp %g %g\n", c.name, bad, n, ho[0], ho[1], ho[c.box[0]], h[c.z * (c.rank == 3 ? c.strides[1] / 4 : 0) + c.y * (c.strides[0] / 4) + c.x], h[c.z * (c.rank == 3 ? c.strides[1] / 4 : 0) + (c.y+1) * (c.strides[0] / 4) + c.x]);
  }
  return 0;
}
