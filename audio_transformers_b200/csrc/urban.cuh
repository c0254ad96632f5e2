// urban.cuh -- urban preset: fused mel kernel and the pre-step kernels (resample, peak normalisation).
#pragma once
// ------------------------------------------------------------------------------------------------
// Urban preset (replaces TA:transforms/_transforms.py:621-631 MelSpectrogram.forward and the
// torch.log(mel + 1e-9) of REF:urban_sounds/dataset.py:56)
//
// Same frame-per-lane scheme: a CTA owns 32 consecutive frames of one clip.  1024 = 32 x 32 is not
// coprime, so this is a Cooley-Tukey split (n = r + 32 j, k = k2 + 32 k1) with twiddles
// W1024^(r k2) between the passes:
//   pass 1  32 tasks (r):  windowed real 32-point DFT, k2 = 0..16                  -> E[32*34][32]
//   pass 2  17 tasks (k2): twiddle, complex 32-point DFT over r, |X|^2 to bin rows  -> P[513][32]
//           (P overlays the audio tile, which is dead after pass 1)
//   mel     64 HTK filters (998 taps), optional log(. + eps), coalesced stores
// ------------------------------------------------------------------------------------------------
constexpr int U_NFFT = 1024, U_HOP = 512, U_NMEL = 64, U_NBIN = 513;
constexpr int U_TILE = 32, U_THREADS = 512, U_WARPS = U_THREADS / 32;
constexpr int U_SPAN = (U_TILE - 1) * U_HOP + U_NFFT;           // 16896 samples
constexpr int U_PITCH = U_HOP + 1;                              // 513
constexpr int U_ROWS = U_SPAN / U_HOP;                          // 33
constexpr int U_SM_AUDIO = ((U_ROWS * U_PITCH + 31) / 32) * 32; // >= 513*32 (P overlay)
constexpr int U_EROWS = 32 * 34;
constexpr int U_SM_E = U_EROWS * 32;
constexpr int U_SMEM_BYTES = (U_SM_E + U_SM_AUDIO) * 4;
static_assert(U_SM_AUDIO >= U_NBIN * 32, "P overlay must fit in the audio tile");

__constant__ int c_umel_start[64] = kUMelStart_INIT;
__constant__ int c_umel_len[64] = kUMelLen_INIT;
__constant__ int c_umel_off[64] = kUMelOff_INIT;

__device__ __forceinline__ void u_pass1(int r, const float* __restrict__ audio_lane, float* __restrict__ e_lane) {
  float x[32], w[32], Xr[17], Xi[17];
  const float* src = audio_lane + r;
  const float* win = c_win1024 + r;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    x[j] = src[(j / 16) * U_PITCH + (j % 16) * 32];
    w[j] = win[32 * j];
  }
  b2::real_dft32(x, w, Xr, Xi);
  float* dst = e_lane + r * (34 * 32);
#pragma unroll
  for (int k = 0; k < 17; ++k) { dst[(2 * k) * 32] = Xr[k]; dst[(2 * k + 1) * 32] = Xi[k]; }
}

__device__ __forceinline__ void u_pass2(int k2, const float* __restrict__ e_lane, float* __restrict__ p_lane) {
  float zr[32], zi[32], Xr[32], Xi[32];
  const float* base = e_lane + k2 * 64;
  const float* tc = c_utw_cos + k2 * 32;
  const float* ts = c_utw_sin + k2 * 32;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float yr = base[r * (34 * 32)], yi = base[r * (34 * 32) + 32];
    const float c = tc[r], s = ts[r];                 // W = c - i s
    zr[r] = __fmaf_rn(yi, s, yr * c);
    zi[r] = __fmaf_rn(-yr, s, yi * c);
  }
  b2::cplx_dft32(zr, zi, Xr, Xi);
  float* direct = p_lane + k2 * 32;                   // bin = k2 + 32 k1,          k1 = 0..15
  float* mirror = p_lane - k2 * 32;                   // bin = 32 (32 - k1) - k2,   k1 = 16..31
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) direct[(32 * k1) * 32] = __fmaf_rn(Xr[k1], Xr[k1], Xi[k1] * Xi[k1]);
#pragma unroll
  for (int k1 = 16; k1 < 32; ++k1) mirror[(32 * (32 - k1)) * 32] = __fmaf_rn(Xr[k1], Xr[k1], Xi[k1] * Xi[k1]);
}

__global__ void __launch_bounds__(U_THREADS, 1)
urban_mel_kernel(const float* __restrict__ wave, long long stride, int n_samples, int n_frames, int tiles_per_clip,
                 int batch, float log_eps, float* __restrict__ out) {
  extern __shared__ __align__(1024) float smem[];
  float* s_e = smem;
  float* s_audio = smem + U_SM_E;                     // later reused as P[513][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int clip = blockIdx.x / tiles_per_clip;
  const int f0 = (blockIdx.x - clip * tiles_per_clip) * U_TILE;
  if (clip >= batch) return;
  const float* __restrict__ src = wave + (size_t)clip * (size_t)stride;

  {   // stage: reflect padding (n_fft/2 each side, no edge repeat) on the n_samples-long clip
    const int g0 = f0 * U_HOP - U_NFFT / 2;
    const bool interior = (g0 >= 0) && (g0 + U_ROWS * U_HOP <= n_samples);
    for (int r = warp; r < U_ROWS; r += U_WARPS) {
      if (interior) {
#pragma unroll
        for (int k = 0; k < U_HOP / 32; ++k) cp_async4(s_audio + r * U_PITCH + lane + 32 * k, src + g0 + r * U_HOP + lane + 32 * k);
      } else {
#pragma unroll 4
        for (int k = 0; k < U_HOP / 32; ++k) {
          const int g = g0 + r * U_HOP + lane + 32 * k;
          const int j = g < 0 ? -g : (g >= n_samples ? 2 * (n_samples - 1) - g : g);
          s_audio[r * U_PITCH + lane + 32 * k] = (j >= 0 && j < n_samples) ? __ldg(src + j) : 0.0f;
        }
      }
    }
    cp_async_commit_wait_all();
  }
  __syncthreads();
  {
    const float* al = s_audio + lane * U_PITCH;
    float* el = s_e + lane;
#pragma unroll 1
    for (int r = warp; r < 32; r += U_WARPS) u_pass1(r, al, el);
  }
  __syncthreads();
  {
    const float* el = s_e + lane;
    float* pl = s_audio + lane;
#pragma unroll 1
    for (int k2 = warp; k2 < 17; k2 += U_WARPS) u_pass2(k2, el, pl);
  }
  __syncthreads();
  {
    const int frame = f0 + lane;
    const float* pl = s_audio + lane;
    float* out_col = out + (size_t)clip * ((size_t)U_NMEL * n_frames) + frame;
#pragma unroll 1
    for (int m = warp; m < U_NMEL; m += U_WARPS) {
      const int start = c_umel_start[m], len = c_umel_len[m], off = c_umel_off[m];
      const float* p = pl + start * 32;
      float acc = 0.0f;
#pragma unroll 4
      for (int j = 0; j < len; ++j) acc = __fmaf_rn(p[j * 32], c_umelw[off + j], acc);
      if (frame < n_frames) out_col[(size_t)m * n_frames] = (log_eps >= 0.0f) ? __logf(acc + log_eps) : acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Urban pre-steps (REF:urban_sounds/dataset.py:26-52, process_audio before the mel transform): mono mean,
// torchaudio's sinc/Hann polyphase resampler (TA:functional/functional.py _get_sinc_resample_kernel /
// _apply_sinc_resample_kernel: y[m*new + p] = sum_k kernel[p][k] * xpad[m*orig + k], xpad = x zero-padded by
// `width` on the left), pad/trim to the target length and peak normalisation.
// ------------------------------------------------------------------------------------------------
constexpr int UP_THREADS = 256, UP_ITEMS = 4;

__global__ void __launch_bounds__(UP_THREADS)
urban_prep_kernel(const float* __restrict__ audio, long long in_stride, const int* __restrict__ in_lengths, int channels,
                  int orig, int nw, const float* __restrict__ taps, int width,
                  float* __restrict__ out, long long out_stride, int out_samples, unsigned int* __restrict__ clip_max_bits) {
  const int clip = blockIdx.y;
  const float* __restrict__ src = audio + (size_t)clip * (size_t)channels * (size_t)in_stride;
  long long L = in_lengths ? (long long)in_lengths[clip] : in_stride;
  L = L < 0 ? 0 : (L > in_stride ? in_stride : L);
  const long long resampled = (L * nw + orig - 1) / orig;            // ceil(new * L / orig)
  const int ktaps = 2 * width + orig;
  const float inv_ch = 1.0f / (float)channels;
  float amax = 0.0f;
#pragma unroll
  for (int it = 0; it < UP_ITEMS; ++it) {
    const int j = (blockIdx.x * UP_ITEMS + it) * UP_THREADS + threadIdx.x;
    if (j >= out_samples) continue;
    float y = 0.0f;
    if (j < resampled) {
      if (orig == nw) {                                              // Resample is skipped when the rates agree
        float sacc = 0.0f;
        for (int c = 0; c < channels; ++c) sacc += __ldg(src + (size_t)c * in_stride + j);
        y = channels > 1 ? sacc * inv_ch : sacc;
      } else {
        const int m = j / nw, p = j - m * nw;
        const float* __restrict__ kp = taps + (size_t)p * ktaps;
        const long long i0 = (long long)m * orig - width;            // first input sample under the filter
        int k0 = i0 < 0 ? (int)(-i0) : 0;
        int k1 = (i0 + ktaps > L) ? (int)(L - i0) : ktaps;
        float acc = 0.0f;
        if (channels == 1) {
          for (int k = k0; k < k1; ++k) acc = __fmaf_rn(__ldg(kp + k), __ldg(src + i0 + k), acc);
        } else {
          for (int k = k0; k < k1; ++k) {
            float sacc = 0.0f;
            for (int c = 0; c < channels; ++c) sacc += __ldg(src + (size_t)c * in_stride + i0 + k);
            acc = __fmaf_rn(__ldg(kp + k), sacc * inv_ch, acc);
          }
        }
        y = acc;
      }
    }
    out[(size_t)clip * out_stride + j] = y;
    amax = fmaxf(amax, fabsf(y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.0f) atomicMax(clip_max_bits + clip, __float_as_uint(amax));
}

__global__ void __launch_bounds__(UP_THREADS)
urban_peak_norm_kernel(float* __restrict__ out, long long out_stride, int out_samples, const unsigned int* __restrict__ clip_max_bits) {
  const int clip = blockIdx.y;
  const float m = __uint_as_float(clip_max_bits[clip]);
  if (!(m > 0.0f)) return;                                           // REF:urban_sounds/dataset.py:51: only if there is sound
  float* p = out + (size_t)clip * out_stride;
  for (int j = blockIdx.x * UP_THREADS + threadIdx.x; j < out_samples; j += gridDim.x * UP_THREADS) p[j] = __fdiv_rn(p[j], m);
}

