// urban.cuh -- urban preset: geometry and the pre-step kernels (mono mix, resample, pad / trim, peak normalisation).
#pragma once
// ------------------------------------------------------------------------------------------------
// Urban preset geometry (TA:transforms/_transforms.py:566-631 with the reference's arguments,
// REF:urban_sounds/dataset.py:19-24).  The fused mel kernel lives in urban_packed.cuh.
// ------------------------------------------------------------------------------------------------
constexpr int U_NFFT = 1024, U_HOP = 512, U_NMEL = 64, U_NBIN = 513;

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

