// b200mel.cu -- the C ABI of libb200mel.so (include/b200mel.h) over the hand-written sm_100a kernels.
//
// Whisper preset (replaces HF:models/whisper/feature_extraction_whisper.py:135-164 plus the pad/trim of
// HF:feature_extraction_sequence_utils.py:263-278,327-332): one fused persistent kernel turns float32 audio into
// normalised log-mel -- TMA-staged audio tiles, two-pass prime-factor real FFT (25 x 16) on packed f32x2 frame
// pairs, |X|^2, sparse mel projection, log, per-clip max -- followed by a small in-place pass that applies the
// clip-wide floor.  whisper_common.cuh holds the building blocks, whisper_tile32.cuh the default kernel (32-frame
// tiles, two CTAs per SM), whisper_tile64.cuh the 64-frame variant, whisper_post.cuh the floor / mask kernels.
// Urban preset (urban.cuh): fused 1024-point mel kernel and the pre-step kernels (resample, peak normalisation).
//
// Nothing but the audio (read once) and the features touches HBM.  DESIGN.md has the layouts and the measurements.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <utility>
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
namespace dev_tab {                       // global-memory copies, for tables a kernel indexes per lane
#define B200MEL_CONST static __device__ const
#include "generated/tables.inc"
#undef B200MEL_CONST
}  // namespace dev_tab
#include "fft_codelets.cuh"

namespace {

#include "whisper_common.cuh"
#include "whisper_tile64.cuh"
#include "whisper_tile32.cuh"
#include "whisper_post.cuh"
#include "whisper_pipe.cuh"
#include "urban.cuh"
#include "urban_packed.cuh"

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
  const float* win400 = nullptr;     // Whisper preset: device address of the window table (global memory)
  float* uimg = nullptr;             // urban preset: table image of the packed kernel (device)
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
  e = cudaFuncSetAttribute(whisper_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_kernel32, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_kernel32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xp::X_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(whisper_logmel_pipe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(urban_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(urban_mel_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM_BYTES);
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
  if (preset == B200MEL_PRESET_WHISPER) {
    void* sym = nullptr;
    e = cudaGetSymbolAddress(&sym, dev_tab::c_win400);
    if (e != cudaSuccess) { delete h; cudaSetDevice(prev); return fail_cuda(e, "cudaGetSymbolAddress"); }
    h->win400 = (const float*)sym;
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
  // one 64-bit word per clip plus the CTA counter (whisper_pipe.cuh); the legacy kernels want a float per (clip, tile, warp)
  size_t need = ((size_t)batch + 2 + 8 * 2 * 160) * sizeof(unsigned long long);   // + development counters
#ifdef B200MEL_LEGACY_KERNELS
  const size_t old = (size_t)batch * V_SLOTS_PER_CLIP * sizeof(float);
  need = need > old ? need : old;
#endif
  return (need + 255) & ~(size_t)255;
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
  static const bool legacy = getenv("B200MEL_KERNEL32") != nullptr || getenv("B200MEL_KERNEL64") != nullptr;
  if (!legacy) {
    // one persistent warp-specialised CTA per SM; the clip floor is applied inside the same kernel
    const long long nt = (long long)batch * xp::X_TILES_PER_CLIP;
    const int grid = nt < (long long)h->sm_count ? (int)nt : h->sm_count;
    xp::XArgs args;
    args.wave = wave; args.stride = (long long)stride_samples; args.lengths = lengths; args.batch = batch;
    args.out = out; args.ws = (unsigned long long*)workspace; args.win400 = h->win400;
    { static const char* dbg = getenv("B200MEL_PIPE_DEBUG"); args.debug = dbg ? atoi(dbg) : 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(xp::X_THREADS);
    cfg.dynamicSmemBytes = xp::X_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    { static const bool no_pdl = getenv("B200MEL_NO_PDL") != nullptr; cfg.numAttrs = no_pdl ? 0 : 1; }
    const bool prof = h->prof_on && h->prof_n < h->prof_cap;
    if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], stream);
    cudaError_t le = cudaLaunchKernelEx(&cfg, whisper_logmel_pipe_kernel, args);
    if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
    if (le != cudaSuccess) return fail_cuda(le, "whisper_logmel_pipe_kernel launch");
    return B200MEL_OK;
  }
  const bool k32 = getenv("B200MEL_KERNEL64") == nullptr;        // default: 32-frame tiles, two CTAs per SM
  unsigned int* clip_max = (unsigned int*)workspace;
  cudaError_t e = cudaSuccess;
  if (!k32) {
    e = cudaMemsetAsync(clip_max, 0, (size_t)batch * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
  }
  // TMA view of the audio: [clip][y][x], element (x, y, clip) = wave[clip * stride + 160 y + x], x < 284.  Rows
  // overlap (y-stride 160 samples < 284), which is what lets a box start at any sample with 16-byte aligned
  // strides.  NY is chosen so that every in-bounds element lies inside its clip's row of the buffer.
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int use_tma = 0;
  if (stride_samples >= W_TMAP_X) {
    const cuuint64_t dims[3] = {(cuuint64_t)W_TMAP_X, (cuuint64_t)((stride_samples - W_TMAP_X) / W_HOP + 1), (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)W_HOP * sizeof(float), (cuuint64_t)stride_samples * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)W_PITCH, (cuuint32_t)(k32 ? V_ROWS : W_ROWS), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)wave, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200MEL_ERR_CUDA, "whisper_logmel: cuTensorMapEncodeTiled failed");
    use_tma = getenv("B200MEL_DEBUG_NO_TMA") ? 0 : 1;   // debug knob: every tile through the generic staging path
  }
  const int ntiles = batch * W_TILES_PER_CLIP;
  const int grid_main = ntiles < h->sm_count ? ntiles : h->sm_count;   // persistent: one 512-thread CTA per SM
  const bool prof = h->prof_on && h->prof_n < h->prof_cap;
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], stream);
  if (k32) {
    const long long nt = (long long)batch * V_TILES_PER_CLIP;
    const int grid32 = nt < (long long)h->sm_count ? (int)nt : h->sm_count;     // one 512-thread CTA (two halves) per SM
    // programmatic stream serialisation: the kernel may begin while the previous kernel of the stream is finishing
    // (it waits with griddepcontrol.wait before its first global write); with profiling on the event records sit
    // between the kernels and switch the overlap off
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
    e = cudaLaunchKernelEx(&cfg, whisper_logmel_kernel32, tmap, use_tma, wave, (long long)stride_samples, lengths, (int)batch, out,
                           (float*)workspace);
    if (e != cudaSuccess) return fail_cuda(e, "whisper_logmel_kernel32 launch");
  } else {
    whisper_logmel_kernel<<<grid_main, W_THREADS, W_SMEM_BYTES, stream>>>(
        tmap, use_tma, wave, (long long)stride_samples, lengths, batch, out, clip_max);
  }
  if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "whisper_logmel_kernel launch");
  if (k32) {
    const int items = batch * CL_PARTS;
    const int grid = items < CL_CTAS_PER_SM * h->sm_count ? items : CL_CTAS_PER_SM * h->sm_count;
    whisper_clamp_kernel32<<<grid, CL_THREADS, 0, stream>>>(out, (const float*)workspace, batch, lengths, (long long)stride_samples);
  } else {
    dim3 grid(30, batch);
    whisper_clamp_kernel<<<grid, 256, 0, stream>>>(out, clip_max, batch);
  }
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
  if (!h || h->preset != B200MEL_PRESET_URBAN) return fail(B200MEL_ERR_BAD_ARG, "mel: handle is not an urban-preset handle");
  if (batch < 0) return fail(B200MEL_ERR_BAD_ARG, "mel: negative batch");
  if (batch == 0) return B200MEL_OK;
  if (!wave || !out) return fail(B200MEL_ERR_BAD_ARG, "mel: NULL wave/out");
  if (n_samples <= U_NFFT / 2) return fail(B200MEL_ERR_BAD_ARG, "mel: n_samples must exceed n_fft/2 = 512 (reflect padding)");
  if (stride_samples < n_samples || (stride_samples & 3)) return fail(B200MEL_ERR_BAD_ARG, "mel: stride_samples must be >= n_samples and a multiple of 4");
  if (((uintptr_t)wave & 15) || ((uintptr_t)out & 3)) return fail(B200MEL_ERR_BAD_ALIGN, "mel: wave must be 16-byte aligned");
  const int n_frames = 1 + n_samples / U_HOP;
  const int tiles_per_clip = (n_frames + U_TILE - 1) / U_TILE;
  if ((long long)batch * tiles_per_clip > 0x7fffffffLL) return fail(B200MEL_ERR_BAD_ARG, "mel: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool prof = h->prof_on && h->prof_n < h->prof_cap;
  if (prof) cudaEventRecord(h->prof_ev[2 * h->prof_n], stream);
  static const bool v1 = getenv("B200MEL_URBAN_V1") != nullptr;     // first kernel, kept for A/B runs
  if (v1) {
    urban_mel_kernel<<<batch * tiles_per_clip, U_THREADS, U_SMEM_BYTES, stream>>>(
        wave, (long long)stride_samples, n_samples, n_frames, tiles_per_clip, batch, log_eps, out);
  } else {
    const long long total_frames = (long long)batch * n_frames;
    if (total_frames > 0x7fffff00LL) return fail(B200MEL_ERR_BAD_ARG, "mel: batch * frames must fit 31 bits");
    const int n_tiles = (int)((total_frames + 31) / 32);
    urban_mel_packed_kernel<<<n_tiles < h->sm_count ? n_tiles : h->sm_count, U2_THREADS, U2_SMEM_BYTES, stream>>>(
        wave, (long long)stride_samples, n_samples, n_frames, (unsigned)total_frames, n_tiles, log_eps, h->uimg, out);
  }
  if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "urban_mel_kernel launch");
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

#ifdef W_TRACE
int b200mel_debug_set_trace(void* dev_ptr) {
  long long* p = (long long*)dev_ptr;
  return cudaMemcpyToSymbol(g_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
#endif

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
          const double* s = (const double*)clips[i];
          for (int64_t k = lo; k < hi; ++k) d[k] = (float)s[k];
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
