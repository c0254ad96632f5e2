// b200mel.cu -- the C ABI of libb200mel.so (include/b200mel.h) over the hand-written sm_100a kernels.
//
// Whisper preset (replaces HF:models/whisper/feature_extraction_whisper.py:135-164 plus the pad/trim of
// HF:feature_extraction_sequence_utils.py:263-278,327-332): one fused persistent kernel turns float32 audio into
// normalised log-mel -- TMA-staged audio tiles, two-pass prime-factor real FFT (25 x 16) on packed f32x2 frame
// pairs, |X|^2, sparse mel projection, log, per-clip max -- followed by a small in-place pass that applies the
// clip-wide floor (it re-reads the features once, from L2 at the reference's batch sizes).  whisper_common.cuh holds
// geometry and primitives, whisper_tile32.cuh the kernel (32-frame tiles, two independent halves per SM),
// whisper_post.cuh the floor / mask kernels.  Urban preset: urban_packed.cuh (fused 1024-point mel kernel) and
// urban.cuh (the pre-step kernels: mono mix, resample, peak normalisation).  DESIGN.md has the layouts and the measurements.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <unistd.h>

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

#include "whisper_common.cuh"
#include "whisper_tile32.cuh"
#include "whisper_post.cuh"
#include "urban.cuh"
#include "urban_packed.cuh"
#include "encoder_stem.cuh"

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

// float64 -> float32 into the pinned staging buffer with non-temporal stores: the destination is only read by the copy
// engine, so pulling its lines into the cache first (write-allocate) would be a third of the cast's memory traffic.
#if defined(__x86_64__)
#include <immintrin.h>
namespace {
__attribute__((target("avx2"))) void cast_f64_f32_avx2(const double* __restrict__ s, float* __restrict__ d, int64_t n) {
  int64_t k = 0;
  for (; k < n && ((uintptr_t)(d + k) & 31); ++k) d[k] = (float)s[k];
  for (; k + 8 <= n; k += 8) {
    const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(s + k)), hi = _mm256_cvtpd_ps(_mm256_loadu_pd(s + k + 4));
    _mm256_stream_ps(d + k, _mm256_set_m128(hi, lo));
  }
  for (; k < n; ++k) d[k] = (float)s[k];
  _mm_sfence();
}
}  // namespace
#endif
namespace {
void cast_f64_f32(const double* __restrict__ s, float* __restrict__ d, int64_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) { cast_f64_f32_avx2(s, d, n); return; }
#endif
  for (int64_t k = 0; k < n; ++k) d[k] = (float)s[k];
}
}  // namespace

// Persistent host worker pool for b200mel_host_pack: creating threads per call costs more than converting a 30 s
// clip (measured: 0.33 ms on one thread, 0.68 ms when 7 fresh threads share the clip).  Workers sleep on a condition
// variable between calls.  One job at a time (callers are serialised by a mutex); a process that forked after the
// pool was built (DataLoader workers) gets a fresh pool, the inherited one has no threads behind it.
namespace {
class PackPool {
 public:
  static PackPool& get() {
    static std::mutex guard;
    static PackPool* inst = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(guard);
    if (!inst || owner != getpid()) { inst = new PackPool(); owner = getpid(); }   // the stale pool of the parent is leaked on purpose
    return *inst;
  }
  // run fn(part), part = 0..parts-1, on the caller plus up to parts-1 workers; returns when all parts are done
  void run(int parts, const std::function<void(int)>& fn) {
    if (parts <= 1) { fn(0); return; }
    std::lock_guard<std::mutex> job_lk(job_mutex_);
    grow(parts - 1);
    std::unique_lock<std::mutex> lk(m_);
    fn_ = &fn; parts_ = parts; next_ = 0; pending_ = parts; ++generation_;
    gen_hint_.store(generation_, std::memory_order_release);
    const unsigned long long mine = generation_;
    cv_.notify_all();
    drain(lk, mine);                                   // the caller works too, it does not just wait
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  void grow(int want) {
    while ((int)workers_.size() < want && workers_.size() < 64) {
      workers_.emplace_back([this] { loop(); });
      workers_.back().detach();
    }
  }
  // claim and run parts of job `gen` until none are left; called and returns with m_ held.  Parts are claimed under
  // the lock, so a worker that is late can never run a part of a newer job with the older job's function.
  void drain(std::unique_lock<std::mutex>& lk, unsigned long long gen) {
    while (generation_ == gen && next_ < parts_) {
      const int part = next_++;
      const std::function<void(int)>* fn = fn_;
      lk.unlock();
      (*fn)(part);
      lk.lock();
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  void loop() {
    unsigned long long seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      if (generation_ == seen) {
        // calls tend to come back to back (one per clip in the reference's loop): look for the next job for ~100 us
        // before going to sleep -- waking a sleeping thread costs several times the work it is then given
        lk.unlock();
        const auto t0 = std::chrono::steady_clock::now();
        while (gen_hint_.load(std::memory_order_acquire) == seen &&
               std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(100)) {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
        }
        lk.lock();
      }
      cv_.wait(lk, [&] { return generation_ != seen; });
      seen = generation_;
      drain(lk, seen);
    }
  }
  std::mutex job_mutex_, m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, parts_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  std::atomic<unsigned long long> gen_hint_{0};      // copy of generation_ the workers may poll without the lock
};
}  // namespace

// cuTensorMapEncodeTiled, resolved through the runtime so that the library does not link libcuda directly
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct b200mel_handle {
  int device;
  int preset;
  int sm_count;
  tmap_encode_fn encode = nullptr;   // Whisper preset: TMA descriptor encoder
  float* uimg = nullptr;             // urban preset: table image of the packed kernel (device)
  bool no_tma = false;               // test hook (B200MEL_DEBUG_NO_TMA=1 at create time): every audio tile takes the ordinary-store path
  // optional benchmark instrumentation (b200mel_profile_begin/end)
  bool prof_on = false;
  int prof_cap = 0, prof_n = 0;
  cudaEvent_t* prof_ev = nullptr;   // 2 * prof_cap events
  b200mel_handle(int d, int p, int s) : device(d), preset(p), sm_count(s) {}
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
  e = cudaFuncSetAttribute(whisper_logmel_kernel32, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_kernel32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(urban_mel_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(es_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ES_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(es_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ES_SMEM_BYTES);
  if (e != cudaSuccess) { cudaSetDevice(prev); return fail_cuda(e, "cudaFuncSetAttribute"); }
  b200mel_handle* h = new b200mel_handle(device, preset, prop.multiProcessorCount);
  if (preset == B200MEL_PRESET_URBAN) {
    std::vector<float> img;
    u2_build_image(img);
    if (img.empty()) { delete h; cudaSetDevice(prev); return fail(B200MEL_ERR_BAD_ARG, "b200mel_create: urban filter table does not fit its image"); }
    e = cudaMalloc((void**)&h->uimg, sizeof(float) * U2_IMG);
    if (e == cudaSuccess) e = cudaMemcpy(h->uimg, img.data(), sizeof(float) * U2_IMG, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (h->uimg) cudaFree(h->uimg);
      delete h;
      cudaSetDevice(prev);
      return fail_cuda(e, "b200mel_create: table upload");
    }
  }
  cudaSetDevice(prev);
  if (preset == B200MEL_PRESET_WHISPER) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      delete h;
      return fail(B200MEL_ERR_CUDA, "b200mel_create: cuTensorMapEncodeTiled is not available from this driver");
    }
    h->encode = (tmap_encode_fn)fn;
    h->no_tma = getenv("B200MEL_DEBUG_NO_TMA") != nullptr;
  }
  *out = h;
  return B200MEL_OK;
}

static void prof_free(b200mel_handle* h) {
  if (h->prof_ev) {
    for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
  }
  h->prof_ev = nullptr; h->prof_cap = 0; h->prof_n = 0; h->prof_on = false;
}

int b200mel_destroy(b200mel_handle* h) {
  if (h) prof_free(h);
  if (h && h->uimg) cudaFree(h->uimg);
  delete h;
  return B200MEL_OK;
}

int b200mel_profile_begin(b200mel_handle* h, int32_t max_launches) {
  if (!h || max_launches <= 0) return fail(B200MEL_ERR_BAD_ARG, "profile_begin: bad handle or max_launches");
  prof_free(h);
  h->prof_ev = new cudaEvent_t[2 * (size_t)max_launches];
  for (int i = 0; i < 2 * max_launches; ++i) {
    cudaError_t e = cudaEventCreate(&h->prof_ev[i]);
    if (e != cudaSuccess) { h->prof_cap = i / 2; prof_free(h); return fail_cuda(e, "cudaEventCreate"); }
  }
  h->prof_cap = max_launches; h->prof_n = 0; h->prof_on = true;
  return B200MEL_OK;
}

int b200mel_profile_end(b200mel_handle* h, double* total_ms, int32_t* launches) {
  if (!h || !total_ms || !launches) return fail(B200MEL_ERR_BAD_ARG, "profile_end: NULL argument");
  double sum = 0.0;
  for (int i = 0; i < h->prof_n; ++i) {
    cudaError_t e = cudaEventSynchronize(h->prof_ev[2 * i + 1]);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]);
    if (e != cudaSuccess) { prof_free(h); return fail_cuda(e, "profile_end"); }
    sum += ms;
  }
  *total_ms = sum; *launches = h->prof_n;
  prof_free(h);
  return B200MEL_OK;
}

size_t b200mel_workspace_bytes(const b200mel_handle* h, int32_t batch) {
  if (!h || batch <= 0) return 0;
  if (h->preset != B200MEL_PRESET_WHISPER) return 0;
  // one {max, min} pair per (clip, tile, warp)
  return ((size_t)batch * V_SLOTS_PER_CLIP * sizeof(float2) + 255) & ~(size_t)255;
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
  if ((long long)batch * V_TILES_PER_CLIP > 0x7fffffffLL) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  // TMA view of the audio: [clip][y][x], element (x, y, clip) = wave[clip * stride + 160 y + x], x < 284.  Rows
  // overlap (y-stride 160 samples < 284), which is what lets a box start at any sample with 16-byte aligned
  // strides.  NY is chosen so that every in-bounds element lies inside its clip's row of the buffer.
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int use_tma = 0;
  if (stride_samples >= W_TMAP_X) {
    const cuuint64_t dims[3] = {(cuuint64_t)W_TMAP_X, (cuuint64_t)((stride_samples - W_TMAP_X) / W_HOP + 1), (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)W_HOP * sizeof(float), (cuuint64_t)stride_samples * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)W_PITCH, (cuuint32_t)V_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)wave, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200MEL_ERR_CUDA, "whisper_logmel: cuTensorMapEncodeTiled failed");
    use_tma = h->no_tma ? 0 : 1;
  }
  const bool prof = h->prof_on && h->prof_n < h->prof_cap;
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], stream);
  const long long nt = (long long)batch * V_TILES_PER_CLIP;
  const int grid32 = nt < (long long)h->sm_count ? (int)nt : h->sm_count;     // one 512-thread CTA (two halves) per SM
  // programmatic stream serialisation: the kernel may become resident while the previous kernel of the stream is
  // finishing (it waits with griddepcontrol.wait before its first global access); with profiling on the event records
  // sit between the kernels and switch the overlap off
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid32);
  cfg.blockDim = dim3(V_THREADS);
  cfg.dynamicSmemBytes = V_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, whisper_logmel_kernel32, tmap, use_tma, wave, (long long)stride_samples, lengths,
                                     (int)batch, out, (float2*)workspace);
  if (e != cudaSuccess) return fail_cuda(e, "whisper_logmel_kernel32 launch");
  if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
  const int items = batch * CL_PARTS;
  const int grid = items < CL_CTAS_PER_SM * h->sm_count ? items : CL_CTAS_PER_SM * h->sm_count;
  whisper_clamp_kernel32<<<grid, CL_THREADS, 0, stream>>>(out, (const float2*)workspace, batch, lengths, (long long)stride_samples);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_clamp_kernel32 launch");
  return B200MEL_OK;
}

// ---- encoder stem ----------------------------------------------------------------------------
static size_t es_a1_bytes(int32_t batch) { return (((size_t)batch * ES_T * ES_K1 * 2) + 1023) & ~(size_t)1023; }
static size_t es_h_bytes(int32_t batch) { return (((size_t)batch * ES_HROWS * ES_D * 2) + 1023) & ~(size_t)1023; }

size_t b200mel_encoder_stem_workspace_bytes(const b200mel_handle* h, int32_t batch) {
  if (!h || batch <= 0 || h->preset != B200MEL_PRESET_WHISPER) return 0;
  return es_a1_bytes(batch) + es_h_bytes(batch);
}

static bool es_encode(b200mel_handle* h, CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box) {
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  memset(tm, 0, sizeof(*tm));
  return h->encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int b200mel_encoder_stem_bf16(b200mel_handle* h, const float* features, int32_t batch, const void* w1, const float* bias1,
                              const void* w2, const float* bias2, const float* positions, float* out,
                              void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_WHISPER) return fail(B200MEL_ERR_BAD_ARG, "encoder_stem: handle is not a Whisper-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "encoder_stem: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!features || !w1 || !bias1 || !w2 || !bias2 || !positions || !out) return fail(B200MEL_ERR_BAD_ARG, "encoder_stem: NULL argument");
  if (((uintptr_t)features | (uintptr_t)w1 | (uintptr_t)bias1 | (uintptr_t)w2 | (uintptr_t)bias2 | (uintptr_t)positions |
       (uintptr_t)out | (uintptr_t)workspace) & 15)
    return fail(B200MEL_ERR_BAD_ALIGN, "encoder_stem: every pointer must be 16-byte aligned");
  if (!workspace || workspace_bytes < b200mel_encoder_stem_workspace_bytes(h, batch))
    return fail(B200MEL_ERR_WORKSPACE, "encoder_stem: workspace too small (see b200mel_encoder_stem_workspace_bytes)");
  if ((long long)batch * 48 > 0x7fffffffLL / 2) return fail(B200MEL_ERR_BAD_ARG, "encoder_stem: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  __nv_bfloat16* a1 = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* hid = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(workspace) + es_a1_bytes(batch));

  CUtensorMap tm_a1, tm_w1, tm_h, tm_w2;
  {
    // im2col rows of conv1 as (k, -, t, clip); the second dimension exists only so that both GEMMs use 4-D coordinates
    const cuuint64_t dims[4] = {ES_K1, 1, ES_T, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {ES_K1 * 2, ES_K1 * 2, (cuuint64_t)ES_T * ES_K1 * 2};
    const cuuint32_t box[4] = {ES_BK, 1, ES_BM, 1};
    if (!es_encode(h, &tm_a1, a1, 4, dims, strides, box)) return fail(B200MEL_ERR_CUDA, "encoder_stem: tensor map (A1)");
  }
  {
    // rows of h as (channel, row parity, row pair, clip): tap k of output row t' is row 2 t' + k = (pair t' + k / 2, parity k % 2)
    const cuuint64_t dims[4] = {ES_D, 2, ES_HROWS / 2, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {ES_D * 2, 2 * ES_D * 2, (cuuint64_t)ES_HROWS * ES_D * 2};
    const cuuint32_t box[4] = {ES_BK, 1, ES_BM, 1};
    if (!es_encode(h, &tm_h, hid, 4, dims, strides, box)) return fail(B200MEL_ERR_CUDA, "encoder_stem: tensor map (h)");
  }
  {
    const cuuint64_t dims[2] = {ES_K1, ES_D};
    const cuuint64_t strides[1] = {ES_K1 * 2};
    const cuuint32_t box[2] = {ES_BK, ES_BN};
    if (!es_encode(h, &tm_w1, w1, 2, dims, strides, box)) return fail(B200MEL_ERR_CUDA, "encoder_stem: tensor map (W1)");
  }
  {
    const cuuint64_t dims[2] = {ES_K2, ES_D};
    const cuuint64_t strides[1] = {ES_K2 * 2};
    const cuuint32_t box[2] = {ES_BK, ES_BN};
    if (!es_encode(h, &tm_w2, w2, 2, dims, strides, box)) return fail(B200MEL_ERR_CUDA, "encoder_stem: tensor map (W2)");
  }
  es_im2col_kernel<<<dim3((ES_T + ES_IC_FRAMES - 1) / ES_IC_FRAMES, batch), ES_IC_THREADS, 0, stream>>>(features, a1, hid);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "es_im2col_kernel launch");
  EsGemm g1{batch, (ES_T + ES_BM - 1) / ES_BM, ES_T, ES_K1 / ES_BK, ES_K1 / ES_BK, bias1, nullptr, hid};
  EsGemm g2{batch, (ES_T2 + ES_BM - 1) / ES_BM, ES_T2, ES_K2 / ES_BK, ES_D / ES_BK, bias2, positions, out};
  const int t1 = batch * g1.mtiles * (ES_D / ES_BN), t2 = batch * g2.mtiles * (ES_D / ES_BN);
  // conv1: a CTA keeps one channel half of W1 resident, so the grid is an even number of CTAs (half of them per channel half)
  es_gemm_kernel<0><<<t1 < (h->sm_count & ~1) ? t1 : (h->sm_count & ~1), ES_THREADS, ES_SMEM_BYTES, stream>>>(tm_a1, tm_w1, g1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "es_gemm_kernel<conv1> launch");
  es_gemm_kernel<1><<<t2 < h->sm_count ? t2 : h->sm_count, ES_THREADS, ES_SMEM_BYTES, stream>>>(tm_h, tm_w2, g2);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "es_gemm_kernel<conv2> launch");
  return B200MEL_OK;
}

int b200mel_whisper_frame_mask(b200mel_handle* h, const int32_t* lengths, int32_t batch,
                               int32_t* mask_out, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_WHISPER) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: handle is not a Whisper-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!lengths || !mask_out) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: NULL lengths/mask_out");
  if (batch > 0x7fffffff / W_NFRAME) return fail(B200MEL_ERR_BAD_ARG, "frame_mask: batch too large for one launch");
  const int n = batch * W_NFRAME;
  whisper_frame_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(lengths, batch, mask_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_frame_mask_kernel launch");
  return B200MEL_OK;
}

int b200mel_mel_f32(b200mel_handle* h, const float* wave, int64_t stride_samples, int32_t n_samples,
                    int32_t batch, float log_eps, float* out, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_URBAN) return fail(B200MEL_ERR_BAD_ARG, "mel: handle is not an urban-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "mel: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!wave || !out) return fail(B200MEL_ERR_BAD_ARG, "mel: NULL wave/out");
  if (n_samples <= U_NFFT / 2) return fail(B200MEL_ERR_BAD_ARG, "mel: n_samples must exceed n_fft/2 = 512 (reflect padding)");
  if (stride_samples < n_samples || (stride_samples & 3)) return fail(B200MEL_ERR_BAD_ARG, "mel: stride_samples must be >= n_samples and a multiple of 4");
  if (((uintptr_t)wave & 15) || ((uintptr_t)out & 3)) return fail(B200MEL_ERR_BAD_ALIGN, "mel: wave must be 16-byte aligned");
  const int n_frames = 1 + n_samples / U_HOP;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool prof = h->prof_on && h->prof_n < h->prof_cap;
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], stream);
  const long long total_frames = (long long)batch * n_frames;
  if (total_frames > 0x7fffff00LL) return fail(B200MEL_ERR_BAD_ARG, "mel: batch * frames must fit 31 bits");
  const int n_tiles = (int)((total_frames + 31) / 32);
  urban_mel_packed_kernel<<<n_tiles < h->sm_count ? n_tiles : h->sm_count, U2_THREADS, U2_SMEM_BYTES, stream>>>(
      wave, (long long)stride_samples, n_samples, n_frames, (unsigned)total_frames, n_tiles, log_eps, h->uimg, out);
  if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "urban_mel_packed_kernel launch");
  return B200MEL_OK;
}

size_t b200mel_urban_prep_workspace_bytes(const b200mel_handle* h, int32_t batch) {
  if (!h || batch <= 0 || h->preset != B200MEL_PRESET_URBAN) return 0;
  return ((size_t)batch * sizeof(unsigned int) + 255) & ~(size_t)255;
}

int b200mel_urban_prep_f32(b200mel_handle* h, const float* audio, int64_t in_stride, const int32_t* in_lengths,
                           int32_t channels, int32_t batch, int32_t orig_freq, int32_t new_freq,
                           const float* taps, int32_t width, float* out, int64_t out_stride, int32_t out_samples,
                           void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_URBAN) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: handle is not an urban-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!audio || !out) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: NULL audio/out");
  if (channels < 1 || in_stride < 1 || out_samples < 1 || out_stride < out_samples)
    return fail(B200MEL_ERR_BAD_ARG, "urban_prep: channels, in_stride, out_samples must be positive and out_stride >= out_samples");
  if (orig_freq < 1 || new_freq < 1) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: frequencies must be positive (pass them divided by their gcd)");
  if (orig_freq != new_freq && (!taps || width < 1)) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: resampling needs the tap table and its width");
  if (!workspace || workspace_bytes < b200mel_urban_prep_workspace_bytes(h, batch))
    return fail(B200MEL_ERR_WORKSPACE, "urban_prep: workspace too small (see b200mel_urban_prep_workspace_bytes)");
  if (batch > 65535) return fail(B200MEL_ERR_BAD_ARG, "urban_prep: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned int* clip_max = (unsigned int*)workspace;
  cudaError_t e = cudaMemsetAsync(clip_max, 0, (size_t)batch * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
  const int per_block = UP_THREADS * UP_ITEMS;
  dim3 grid((out_samples + per_block - 1) / per_block, batch);
  urban_prep_kernel<<<grid, UP_THREADS, 0, stream>>>(audio, (long long)in_stride, in_lengths, channels, orig_freq, new_freq,
                                                     taps, width, out, (long long)out_stride, out_samples, clip_max);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "urban_prep_kernel launch");
  dim3 grid2(24, batch);
  urban_peak_norm_kernel<<<grid2, UP_THREADS, 0, stream>>>(out, (long long)out_stride, out_samples, clip_max);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "urban_peak_norm_kernel launch");
  return B200MEL_OK;
}

// Host-side staging helper: ragged clips (float32 or float64, one pointer per clip) -> one row-major float32 buffer
// (pinned, ideally) with `dst_stride` floats per row, converting and copying with `threads` host threads.  Only
// min(len, max_samples) samples of a clip are copied; the tail of a row is left untouched (the kernels never read it).
int b200mel_host_pack(const void* const* clips, const int64_t* lengths, int32_t n, int32_t src_is_f64,
                      int64_t max_samples, float* dst, int64_t dst_stride, int32_t* out_lengths, int32_t threads) {
  if (n < 0 || (n > 0 && (!clips || !lengths || !dst))) return fail(B200MEL_ERR_BAD_ARG, "host_pack: NULL argument");
  if (dst_stride <= 0 || max_samples < 0) return fail(B200MEL_ERR_BAD_ARG, "host_pack: bad stride / max_samples");
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    int64_t L = lengths[i] < 0 ? 0 : (lengths[i] > max_samples ? max_samples : lengths[i]);
    if (L > dst_stride) return fail(B200MEL_ERR_BAD_ARG, "host_pack: a clip is longer than dst_stride");
    if (L > 0 && !clips[i]) return fail(B200MEL_ERR_BAD_ARG, "host_pack: NULL clip pointer");
    if (out_lengths) out_lengths[i] = (int32_t)L;
    total += L;
  }
  if (total == 0) return B200MEL_OK;
  // work is cut into equal sample ranges over the concatenation of all clips, so one long clip is shared by threads
  int nt = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  const int64_t min_chunk = 1 << 16;
  if ((int64_t)nt > (total + min_chunk - 1) / min_chunk) nt = (int)((total + min_chunk - 1) / min_chunk);
  auto work = [&](int64_t begin, int64_t end) {
    int64_t pos = 0;
    for (int i = 0; i < n && pos < end; ++i) {
      const int64_t L = lengths[i] < 0 ? 0 : (lengths[i] > max_samples ? max_samples : lengths[i]);
      const int64_t lo = begin > pos ? begin - pos : 0, hi = (end - pos) < L ? (end - pos) : L;
      if (lo < hi) {
        float* d = dst + (size_t)i * (size_t)dst_stride;
        if (src_is_f64) {
          cast_f64_f32((const double*)clips[i] + lo, d + lo, hi - lo);
        } else {
          memcpy(d + lo, (const float*)clips[i] + lo, (size_t)(hi - lo) * sizeof(float));
        }
      }
      pos += L;
    }
  };
  if (nt <= 1) { work(0, total); return B200MEL_OK; }
  const int64_t per = (total + nt - 1) / nt;
  PackPool::get().run(nt, [&](int part) {
    const int64_t b = per * part, e = per * (part + 1) < total ? per * (part + 1) : total;
    if (b < e) work(b, e);
  });
  return B200MEL_OK;
}

// The reference's call shape in one native call: ragged HOST clips (float32 or float64, per clip) are converted into the
// pinned staging buffer by the worker pool, and every worker copies its piece to the device as soon as it is
// converted (plain per-piece cudaMemcpyAsync on the call's stream), so the cast, the PCIe transfer and -- once the last
// piece is queued -- the kernels overlap instead of running one after the other.
int b200mel_whisper_logmel_host(b200mel_handle* h, const void* const* clips, const int64_t* lengths,
                                const uint8_t* is_f64, int32_t n, float* pinned, int64_t width,
                                int32_t* pinned_lengths, float* dev_wave, int32_t* dev_lengths, float* out,
                                void* workspace, size_t workspace_bytes, int32_t threads, void* stream_) {
  if (!h || h->preset != B200MEL_PRESET_WHISPER) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: handle is not a Whisper-preset handle");
  if (n < 0) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: negative batch");
  if (n == 0) return B200MEL_OK;
  if (!clips || !lengths || !is_f64 || !pinned || !pinned_lengths || !dev_wave || !dev_lengths || !out)
    return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: NULL argument");
  if (width <= 0 || (width & 3)) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: width must be a positive multiple of 4");
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t L = lengths[i] < 0 ? 0 : (lengths[i] > W_NSAMP ? W_NSAMP : lengths[i]);
    if (L > width) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: a clip is longer than the staging row");
    if (L > 0 && !clips[i]) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel_host: NULL clip pointer");
    pinned_lengths[i] = (int32_t)L;
    total += L;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t first_err = cudaSuccess;
  cudaError_t e = cudaMemcpyAsync(dev_lengths, pinned_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync (lengths)");
  if (total > 0) {
    int nt = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
    // pieces of >= 128 K samples (0.5 MB of float32): small enough to start the first copy early and to give every
    // thread several pieces, large enough that a copy is not all launch overhead
    const int64_t piece = total <= (1 << 19) ? (1 << 16) : (1 << 17);
    const int64_t npieces64 = (total + piece - 1) / piece;
    const int npieces = (int)(npieces64 > 4096 ? 4096 : npieces64);
    const int64_t per = (total + npieces - 1) / npieces;
    if (nt > npieces) nt = npieces;
    const int device = h->device;
    std::atomic<int> next_piece{0};
    std::vector<std::atomic<unsigned char>> done((size_t)npieces);
    for (auto& f : done) f.store(0, std::memory_order_relaxed);
    // the valid samples of piece range [begin, end) of the flattened batch -> per-clip segments
    auto for_segments = [&](int64_t begin, int64_t end, auto&& fn) {
      int64_t pos = 0;
      for (int i = 0; i < n && pos < end; ++i) {
        const int64_t L = pinned_lengths[i];
        const int64_t lo = begin > pos ? begin - pos : 0, hi = (end - pos) < L ? (end - pos) : L;
        if (lo < hi) fn(i, lo, hi);
        pos += L;
      }
    };
    // Every part casts pieces; part 0 also issues the copies, between its own pieces and at the end: ONE thread makes
    // every CUDA call, in order and for runs of finished pieces (up to 4 MB per copy).  Sixteen threads calling
    // cudaMemcpyAsync for 0.5 MB each spent more time in the driver's lock than the copies took (the call cost 3.8 ms for
    // 64 clips against 2.7 ms for the cast alone); a thread that only issued would burn a core the cast can use (with
    // eight ranks on a 32-core host every rank has four).
    const int max_run = (int)(((int64_t)1 << 20) / per > 0 ? ((int64_t)1 << 20) / per : 1);
    auto work = [&](int part) {
      const bool issuer = part == 0;
      int p_issue = 0;
      if (issuer) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
      }
      auto issue = [&](bool drain) {
        while (p_issue < npieces) {
          if (!done[(size_t)p_issue].load(std::memory_order_acquire)) {
            if (!drain) return;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            continue;
          }
          int q = p_issue + 1;
          while (q < npieces && q - p_issue < max_run && done[(size_t)q].load(std::memory_order_acquire)) ++q;
          const int64_t begin = per * p_issue, end = per * q < total ? per * q : total;
          for_segments(begin, end, [&](int i, int64_t lo, int64_t hi) {
            const size_t off = (size_t)i * (size_t)width + (size_t)lo;
            const cudaError_t ce = cudaMemcpyAsync(dev_wave + off, pinned + off, (size_t)(hi - lo) * sizeof(float),
                                                   cudaMemcpyHostToDevice, stream);
            if (ce != cudaSuccess && first_err == cudaSuccess) first_err = ce;
          });
          p_issue = q;
        }
      };
      for (;;) {
        const int p = next_piece.fetch_add(1);
        if (p >= npieces) break;
        const int64_t begin = per * p, end = per * (p + 1) < total ? per * (p + 1) : total;
        for_segments(begin, end, [&](int i, int64_t lo, int64_t hi) {
          float* d = pinned + (size_t)i * (size_t)width;
          if (is_f64[i]) cast_f64_f32((const double*)clips[i] + lo, d + lo, hi - lo);
          else memcpy(d + lo, (const float*)clips[i] + lo, (size_t)(hi - lo) * sizeof(float));
        });
        done[(size_t)p].store(1, std::memory_order_release);
        if (issuer) issue(false);
      }
      if (issuer) issue(true);
    };
    if (nt <= 1) work(0);
    else PackPool::get().run(nt, work);
    if (first_err != cudaSuccess) return fail_cuda(first_err, "cudaMemcpyAsync (audio piece)");
  }
  return b200mel_whisper_logmel_f32(h, dev_wave, width, dev_lengths, n, out, workspace, workspace_bytes, stream_);
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
