// whisper_post.cuh -- the clip-floor pass and the attention-mask kernel of the Whisper preset.
#pragma once
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

