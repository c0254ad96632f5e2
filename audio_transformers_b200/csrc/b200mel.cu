// b200mel.cu -- hand-written sm_100a kernels + the C ABI of libb200mel.so (include/b200mel.h).
//
// Whisper preset (replaces HF:models/whisper/feature_extraction_whisper.py:135-164 plus the
// pad/trim of HF:feature_extraction_sequence_utils.py:263-278,327-332):
//
//   One fused kernel turns float32 audio into normalised log-mel.  A CTA owns a tile of 32
//   consecutive frames of one clip and keeps ONE FRAME PER LANE, so every index-dependent
//   constant (window tap, twiddle, filter weight, shared-memory offset) is warp-uniform and is
//   encoded in the instruction stream (immediates / constant-bank operands):
//
//     stage   audio [160 f0 - 200, 160 f0 + 5160) -> shared memory, reflect-padded at the clip
//             edges, zero beyond the clip length; row pitch 161 words so that the 32 lanes (frames,
//             160 samples apart) hit 32 different banks.
//     pass 1  16 tasks (a = n mod-16 class): windowed real 25-point DFT          -> E[400][32]
//     pass 2  13 tasks (k2 = k mod 25):      complex 16-point DFT, |X|^2          -> P[201][32]
//     mel     80 filters, sparse (391 taps), log, per-clip max via warp shuffle + one atomicMax
//             per CTA; unclamped features are stored with coalesced 128-byte rows.
//   A second, tiny in-place pass applies max(y, ymax - 2) once the clip maximum is known.
//
// Nothing but the audio (read once per tile, +7 % halo) and the features touches HBM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200mel.h"

#define B200MEL_CONST __constant__
#include "generated/tables.inc"          // device copies (constant bank)
#undef B200MEL_CONST
namespace host_tab {                      // host copies, so table queries need no device
#define B200MEL_CONST static const
#include "generated/tables.inc"
#undef B200MEL_CONST
}  // namespace host_tab
#include "fft_codelets.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Whisper geometry
// ------------------------------------------------------------------------------------------------
constexpr int W_NFFT = 400, W_HOP = 160, W_NMEL = 80, W_NSAMP = 480000, W_NFRAME = 3000;
constexpr int W_TILE = 32;                                       // frames per CTA tile (one per lane)
constexpr int W_TILES_PER_CLIP = (W_NFRAME + W_TILE - 1) / W_TILE;   // 94
constexpr int W_THREADS = 256;
constexpr int W_WARPS = W_THREADS / 32;
constexpr int W_SPAN = (W_TILE - 1) * W_HOP + W_NFFT;            // 5360 samples staged per tile
constexpr int W_PITCH = W_HOP + 1;                               // 161: odd pitch -> conflict-free lanes
constexpr int W_ROWS = (W_SPAN + W_HOP - 1) / W_HOP;             // 34
constexpr int W_SM_AUDIO = W_ROWS * W_PITCH;                     // floats
constexpr int W_SM_E = 400 * 32;
constexpr int W_SM_P = 201 * 32;
constexpr int W_SMEM_BYTES = (W_SM_AUDIO + W_SM_E + W_SM_P + 32) * 4;

// y = (log10(e) + 4) / 4 = log2(e) * (log10(2)/4) + 1
__device__ __forceinline__ float w_norm_log(float e) {
  return __fmaf_rn(__log2f(e), 0.07525749891599529f, 1.0f);
}

// ---- pass 1: windowed real 25-point DFT for residue class A ---------------------------------
template <int A>
__device__ __forceinline__ void w_pass1(const float* __restrict__ audio_lane, float* __restrict__ e_lane) {
  float x[25], w[25], o[25];
#pragma unroll
  for (int b = 0; b < 25; ++b) {
    const int n = b2::pfa400_n(A, b);
    x[b] = audio_lane[(n / W_HOP) * W_PITCH + (n % W_HOP)];
    w[b] = c_win400[n];
  }
  b2::real_dft25(x, w, o);
#pragma unroll
  for (int c = 0; c < 25; ++c) e_lane[(A * 25 + c) * 32] = o[c];
}

// ---- pass 2: complex 16-point DFT for k2 = K2, power into P ----------------------------------
template <int K2>
__device__ __forceinline__ void w_pass2(const float* __restrict__ e_lane, float* __restrict__ p_lane) {
  float yr[16], yi[16], Xr[16], Xi[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    if (K2 == 0) { yr[a] = e_lane[(a * 25) * 32]; yi[a] = 0.0f; }
    else { yr[a] = e_lane[(a * 25 + 2 * K2 - 1) * 32]; yi[a] = e_lane[(a * 25 + 2 * K2) * 32]; }
  }
  b2::cplx_dft16(yr, yi, Xr, Xi);
#pragma unroll
  for (int k1 = 0; k1 < (K2 == 0 ? 9 : 16); ++k1)
    p_lane[b2::pfa400_bin(k1, K2) * 32] = __fmaf_rn(Xr[k1], Xr[k1], Xi[k1] * Xi[k1]);
}

// constexpr views of the generated sparse-filter tables, callable in device constant expressions
B2_CX int w_mel_start(int m) { const int t[80] = kWMelStart_INIT; return t[m]; }
B2_CX int w_mel_len(int m) { const int t[80] = kWMelLen_INIT; return t[m]; }
B2_CX int w_mel_off(int m) { const int t[80] = kWMelOff_INIT; return t[m]; }

// ---- mel: one filter -----------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void w_mel_one(const float* __restrict__ p_lane, float* __restrict__ out_row,
                                          bool valid, float& emax) {
  constexpr int START = w_mel_start(M), LEN = w_mel_len(M), OFF = w_mel_off(M);
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < LEN; ++j)
    acc = __fmaf_rn(p_lane[(START + j) * 32], c_wmelw[OFF + j], acc);
  const float e = fmaxf(acc, 1e-10f);
  if (valid) {
    emax = fmaxf(emax, e);
    out_row[(size_t)M * W_NFRAME] = w_norm_log(e);
  }
}

template <int W>
__device__ __forceinline__ void w_mel_warp(const float* p_lane, float* out_row, bool valid, float& emax) {
  w_mel_one<W + 0>(p_lane, out_row, valid, emax);  w_mel_one<W + 8>(p_lane, out_row, valid, emax);
  w_mel_one<W + 16>(p_lane, out_row, valid, emax); w_mel_one<W + 24>(p_lane, out_row, valid, emax);
  w_mel_one<W + 32>(p_lane, out_row, valid, emax); w_mel_one<W + 40>(p_lane, out_row, valid, emax);
  w_mel_one<W + 48>(p_lane, out_row, valid, emax); w_mel_one<W + 56>(p_lane, out_row, valid, emax);
  w_mel_one<W + 64>(p_lane, out_row, valid, emax); w_mel_one<W + 72>(p_lane, out_row, valid, emax);
}

__global__ void __launch_bounds__(W_THREADS, 2)
whisper_logmel_kernel(const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                      int batch, float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  extern __shared__ __align__(16) float smem[];
  float* s_audio = smem;
  float* s_e = s_audio + W_SM_AUDIO;
  float* s_p = s_e + W_SM_E;
  float* s_red = s_p + W_SM_P;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int clip = blockIdx.x / W_TILES_PER_CLIP;
  const int tile = blockIdx.x - clip * W_TILES_PER_CLIP;
  const int f0 = tile * W_TILE;
  if (clip >= batch) return;

  // ---- stage audio (reflect pad at both ends of the 480000-sample padded clip, zeros past L) ----
  {
    long long len_ll = lengths ? (long long)lengths[clip] : stride;
    const int L = (int)(len_ll < 0 ? 0 : (len_ll > W_NSAMP ? W_NSAMP : len_ll));
    const float* __restrict__ src = wave + (size_t)clip * (size_t)stride;
    const int g0 = f0 * W_HOP - W_NFFT / 2;
#pragma unroll 4
    for (int s = tid; s < W_SPAN; s += W_THREADS) {
      int g = g0 + s;
      int j = g < 0 ? -g : (g >= W_NSAMP ? 2 * (W_NSAMP - 1) - g : g);
      float v = (j < L) ? __ldg(src + j) : 0.0f;
      s_audio[(s / W_HOP) * W_PITCH + (s % W_HOP)] = v;
    }
  }
  __syncthreads();

  // ---- pass 1 ---------------------------------------------------------------------------------
  {
    const float* al = s_audio + lane * W_PITCH;
    float* el = s_e + lane;
    switch (warp) {
      case 0: w_pass1<0>(al, el); w_pass1<8>(al, el); break;
      case 1: w_pass1<1>(al, el); w_pass1<9>(al, el); break;
      case 2: w_pass1<2>(al, el); w_pass1<10>(al, el); break;
      case 3: w_pass1<3>(al, el); w_pass1<11>(al, el); break;
      case 4: w_pass1<4>(al, el); w_pass1<12>(al, el); break;
      case 5: w_pass1<5>(al, el); w_pass1<13>(al, el); break;
      case 6: w_pass1<6>(al, el); w_pass1<14>(al, el); break;
      default: w_pass1<7>(al, el); w_pass1<15>(al, el); break;
    }
  }
  __syncthreads();

  // ---- pass 2 ---------------------------------------------------------------------------------
  {
    const float* el = s_e + lane;
    float* pl = s_p + lane;
    switch (warp) {
      case 0: w_pass2<1>(el, pl); w_pass2<9>(el, pl); break;
      case 1: w_pass2<2>(el, pl); w_pass2<10>(el, pl); break;
      case 2: w_pass2<3>(el, pl); w_pass2<11>(el, pl); break;
      case 3: w_pass2<4>(el, pl); w_pass2<12>(el, pl); break;
      case 4: w_pass2<5>(el, pl); w_pass2<0>(el, pl); break;
      case 5: w_pass2<6>(el, pl); break;
      case 6: w_pass2<7>(el, pl); break;
      default: w_pass2<8>(el, pl); break;
    }
  }
  __syncthreads();

  // ---- mel + log + per-clip max ---------------------------------------------------------------
  {
    const int frame = f0 + lane;
    const bool valid = frame < W_NFRAME;
    const float* pl = s_p + lane;
    float* out_row = out + (size_t)clip * (W_NMEL * W_NFRAME) + frame;
    float emax = 0.0f;
    switch (warp) {
      case 0: w_mel_warp<0>(pl, out_row, valid, emax); break;
      case 1: w_mel_warp<1>(pl, out_row, valid, emax); break;
      case 2: w_mel_warp<2>(pl, out_row, valid, emax); break;
      case 3: w_mel_warp<3>(pl, out_row, valid, emax); break;
      case 4: w_mel_warp<4>(pl, out_row, valid, emax); break;
      case 5: w_mel_warp<5>(pl, out_row, valid, emax); break;
      case 6: w_mel_warp<6>(pl, out_row, valid, emax); break;
      default: w_mel_warp<7>(pl, out_row, valid, emax); break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    if (lane == 0) s_red[warp] = emax;
  }
  __syncthreads();
  if (tid == 0) {
    float m = s_red[0];
#pragma unroll
    for (int w = 1; w < W_WARPS; ++w) m = fmaxf(m, s_red[w]);
    // positive floats order like their bit patterns; the slot is zeroed before the launch
    atomicMax(clip_max_bits + clip, __float_as_uint(m));
  }
}

// In-place clamp: y = max(y, ymax - 2)  (== (max(log10 e, log10 emax - 8) + 4) / 4).
__global__ void __launch_bounds__(256)
whisper_clamp_kernel(float* __restrict__ out, const unsigned int* __restrict__ clip_max_bits, int batch) {
  constexpr int VEC_PER_CLIP = W_NMEL * W_NFRAME / 4;
  const int clip = blockIdx.y;
  const float thr = w_norm_log(__uint_as_float(clip_max_bits[clip])) - 2.0f;
  float4* p = reinterpret_cast<float4*>(out + (size_t)clip * (W_NMEL * W_NFRAME));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < VEC_PER_CLIP; i += gridDim.x * blockDim.x) {
    float4 v = p[i];
    if (v.x < thr || v.y < thr || v.z < thr || v.w < thr) {
      v.x = fmaxf(v.x, thr); v.y = fmaxf(v.y, thr); v.z = fmaxf(v.z, thr); v.w = fmaxf(v.w, thr);
      p[i] = v;
    }
  }
}

__global__ void whisper_frame_mask_kernel(const int* __restrict__ lengths, int batch, int* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * W_NFRAME) return;
  const int b = i / W_NFRAME, t = i - b * W_NFRAME;
  int L = lengths[b];
  L = L > W_NSAMP ? W_NSAMP : L;
  mask[i] = (t * W_HOP < L) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
int fail_cuda(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", where, cudaGetErrorString(e));
  return B200MEL_ERR_CUDA;
}

}  // namespace

struct b200mel_handle {
  int device;
  int preset;
  int sm_count;
};

extern "C" {

int b200mel_version(void) { return B200MEL_VERSION; }
const char* b200mel_last_error(void) { return g_err; }

int b200mel_create(int device, int preset, b200mel_handle** out) {
  if (!out) return fail(B200MEL_ERR_BAD_ARG, "b200mel_create: out is NULL");
  *out = nullptr;
  if (preset != B200MEL_PRESET_WHISPER && preset != B200MEL_PRESET_URBAN)
    return fail(B200MEL_ERR_BAD_ARG, "b200mel_create: unknown preset");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount (no CUDA device: this library has no CPU path)");
  if (device < 0 || device >= count) return fail(B200MEL_ERR_BAD_ARG, "b200mel_create: device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(B200MEL_ERR_UNSUPPORTED_ARCH, "b200mel_create: device is not compute capability 10.x (kernels are built for sm_100a only)");
  int prev = 0;
  cudaGetDevice(&prev);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
  e = cudaFuncSetAttribute(whisper_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
  b200mel_handle* h = new b200mel_handle{device, preset, prop.multiProcessorCount};
  *out = h;
  return B200MEL_OK;
}

int b200mel_destroy(b200mel_handle* h) {
  delete h;
  return B200MEL_OK;
}

size_t b200mel_workspace_bytes(const b200mel_handle* h, int32_t batch) {
  if (!h || batch <= 0) return 0;
  if (h->preset != B200MEL_PRESET_WHISPER) return 0;
  return ((size_t)batch * sizeof(unsigned int) + 255) & ~(size_t)255;
}

int b200mel_whisper_logmel_f32(b200mel_handle* h, const float* wave, int64_t stride_samples,
                               const int32_t* lengths, int32_t batch, float* out,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_WHISPER) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: handle is not a Whisper-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!wave || !out) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: NULL wave/out");
  if (stride_samples <= 0 || (stride_samples & 3)) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: stride_samples must be a positive multiple of 4");
  if (((uintptr_t)wave & 15) || ((uintptr_t)out & 15) || ((uintptr_t)workspace & 15))
    return fail(B200MEL_ERR_BAD_ALIGN, "whisper_logmel: wave/out/workspace must be 16-byte aligned");
  if (!workspace || workspace_bytes < b200mel_workspace_bytes(h, batch))
    return fail(B200MEL_ERR_WORKSPACE, "whisper_logmel: workspace too small (see b200mel_workspace_bytes)");
  if ((long long)batch * W_TILES_PER_CLIP > 0x7fffffffLL) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned int* clip_max = (unsigned int*)workspace;
  cudaError_t e = cudaMemsetAsync(clip_max, 0, (size_t)batch * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
  whisper_logmel_kernel<<<batch * W_TILES_PER_CLIP, W_THREADS, W_SMEM_BYTES, stream>>>(
      wave, (long long)stride_samples, lengths, batch, out, clip_max);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_logmel_kernel launch");
  dim3 grid(30, batch);
  whisper_clamp_kernel<<<grid, 256, 0, stream>>>(out, clip_max, batch);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_clamp_kernel launch");
  return B200MEL_OK;
}

int b200mel_whisper_frame_mask(b200mel_handle* h, const int32_t* lengths, int32_t batch,
                               int32_t* mask_out, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_WHISPER) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: handle is not a Whisper-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!lengths || !mask_out) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: NULL lengths/mask_out");
  const int n = batch * W_NFRAME;
  whisper_frame_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(lengths, batch, mask_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_frame_mask_kernel launch");
  return B200MEL_OK;
}

int b200mel_mel_f32(b200mel_handle* h, const float* wave, int64_t stride_samples, int32_t n_samples,
                    int32_t batch, float log_eps, float* out, void* stream_) {
  (void)wave; (void)stride_samples; (void)n_samples; (void)batch; (void)log_eps; (void)out; (void)stream_;
  if (!h || h->preset != B200MEL_PRESET_URBAN) return fail(B200MEL_ERR_BAD_ARG, "mel: handle is not an urban-preset handle");
  return fail(B200MEL_ERR_BAD_ARG, "mel: urban preset kernel not built yet");
}

int64_t b200mel_get_table(int preset, int table, float* dst, int64_t capacity) {
  if (!dst) return fail(B200MEL_ERR_BAD_ARG, "get_table: dst is NULL");
  const bool whisper = preset == B200MEL_PRESET_WHISPER;
  if (!whisper && preset != B200MEL_PRESET_URBAN) return fail(B200MEL_ERR_BAD_ARG, "get_table: unknown preset");
  const int nfft = whisper ? 400 : 1024, nmel = whisper ? 80 : 64, nbin = nfft / 2 + 1;
  if (table == B200MEL_TABLE_WINDOW) {
    if (capacity < nfft) return fail(B200MEL_ERR_BAD_ARG, "get_table: capacity too small");
    memcpy(dst, whisper ? host_tab::c_win400 : host_tab::c_win1024, sizeof(float) * nfft);
    return nfft;
  }
  if (table == B200MEL_TABLE_FILTERBANK) {
    if (capacity < (int64_t)nbin * nmel) return fail(B200MEL_ERR_BAD_ARG, "get_table: capacity too small");
    const float* w = whisper ? host_tab::c_wmelw : host_tab::c_umelw;
    memset(dst, 0, sizeof(float) * (size_t)nbin * nmel);
    for (int m = 0; m < nmel; ++m) {
      const int s = whisper ? host_tab::kWMelStart[m] : host_tab::kUMelStart[m];
      const int l = whisper ? host_tab::kWMelLen[m] : host_tab::kUMelLen[m];
      const int o = whisper ? host_tab::kWMelOff[m] : host_tab::kUMelOff[m];
      for (int j = 0; j < l; ++j) dst[(size_t)(s + j) * nmel + m] = w[o + j];
    }
    return (int64_t)nbin * nmel;
  }
  return fail(B200MEL_ERR_BAD_ARG, "get_table: unknown table");
}

}  // extern "C"
