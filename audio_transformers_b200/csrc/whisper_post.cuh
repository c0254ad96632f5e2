// whisper_post.cuh -- the clip-floor pass and the attention-mask kernel of the Whisper preset.
#pragma once
// In-place clamp: y = max(y, ymax - 2)  (== (max(log10 e, log10 emax - 8) + 4) / 4); the clip maximum is the maximum over
// the clip's (tile, warp) slots.  The slots also carry the smallest energy: a clip whose minimum is not below the floor
// and that has no tile of pure padding has nothing to clamp, and its items return after the 752 slots without touching
// the features (white noise, tone + noise: a dynamic range under 80 dB; chirps and gated noise do reach the floor).
// A few resident CTAs per SM loop over (clip, sixteenth of a clip) work items, so that the whole grid is running from the
// first instant and `launch_dependents` lets the NEXT call's log-mel kernel start underneath this pass.
#ifndef CL_CTAS_PER_SM
#define CL_CTAS_PER_SM 8     // 4 -> 8: the 640 work items of a 64-clip call fit in one wave (step 87.6 -> 83.3 us)
#endif
#ifndef CL_NPARTS
#define CL_NPARTS 16     // 10 -> 16: 1024 items of 60 KB for a 64-clip call, still one wave at 8 CTAs per SM (83.1 -> 82.5 us)
#endif
constexpr int CL_THREADS = 256, CL_PARTS = CL_NPARTS;
static_assert((W_NMEL * W_NFRAME / 4) % CL_PARTS == 0, "a clip must split evenly into parts");
__global__ void __launch_bounds__(CL_THREADS, 8)
whisper_clamp_kernel32(float* __restrict__ out, const float2* __restrict__ tile_max, int batch,
                       const int* __restrict__ lengths, long long stride) {
  asm volatile("griddepcontrol.launch_dependents;");
  __shared__ float s_red[2 * (CL_THREADS / 32)];
  constexpr int VEC_PER_PART = W_NMEL * W_NFRAME / 4 / CL_PARTS;       // 3750
  constexpr int VEC_PER_ROW = W_NFRAME / 4;                            // 750
  const int tid = threadIdx.x;
  const float y_silent = v_norm_log(V_EFLOOR);
  for (int item = blockIdx.x; item < batch * CL_PARTS; item += gridDim.x) {
    const int clip = item / CL_PARTS, part = item - clip * CL_PARTS;
    float m = 0.0f, lo = 3.0e38f;
    for (int i = tid; i < V_SLOTS_PER_CLIP; i += CL_THREADS) {
      const float2 mm = __ldcg(tile_max + (size_t)clip * V_SLOTS_PER_CLIP + i);
      m = fmaxf(m, mm.x);
      lo = fminf(lo, mm.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
    __syncthreads();                                   // s_red of the previous item has been read
    if ((tid & 31) == 0) { s_red[tid >> 5] = m; s_red[CL_THREADS / 32 + (tid >> 5)] = lo; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < CL_THREADS / 32; ++w) { m = fmaxf(m, s_red[w]); lo = fminf(lo, s_red[CL_THREADS / 32 + w]); }
    // slots hold energies scaled by V_ESCALE; the log-mel kernel leaves the reference's clamp at 1e-10 to this pass
    const float thr = fmaxf(v_norm_log(m) - 2.0f, y_silent);
    // frames from `fs` on belong to tiles of pure zero padding, which the log-mel kernel did not write: their value is
    // known without reading (the silent tiles are a suffix of the clip: v_tile_silent is monotone in f0)
    const long long len_ll = lengths ? (long long)lengths[clip] : stride;
    const int L = (int)(len_ll < 0 ? 0 : (len_ll > W_NSAMP ? W_NSAMP : len_ll));
    int fs = W_NFRAME;
    for (int f0 = (V_TILES_PER_CLIP - 1) * V_TILE; f0 >= 0 && v_tile_silent(f0, L); f0 -= V_TILE) fs = f0;
    // nothing below the floor (the comparison is on the values the log-mel kernel stored: same function of the energy)
    // and no tile left for this pass to write: the clip is final as it is
    if (fs == W_NFRAME && !(v_norm_log(lo) < thr)) continue;
    const float ys = fmaxf(y_silent, thr);
    const float4 fill = make_float4(ys, ys, ys, ys);
    float4* p = reinterpret_cast<float4*>(out + (size_t)clip * (W_NMEL * W_NFRAME)) + part * VEC_PER_PART;
    for (int i0 = tid; i0 < VEC_PER_PART; i0 += 4 * CL_THREADS) {
      float4 v[4];
      bool sil[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * CL_THREADS;
        sil[j] = true;
        if (i < VEC_PER_PART) {
          const int frame = ((part * VEC_PER_PART + i) % VEC_PER_ROW) * 4;       // fs is a multiple of 32
          sil[j] = frame >= fs;
          if (!sil[j]) v[j] = p[i];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * CL_THREADS;
        if (i >= VEC_PER_PART) continue;
        if (sil[j]) {
          p[i] = fill;
        } else if (v[j].x < thr || v[j].y < thr || v[j].z < thr || v[j].w < thr) {
          v[j].x = fmaxf(v[j].x, thr); v[j].y = fmaxf(v[j].y, thr); v[j].z = fmaxf(v[j].z, thr); v[j].w = fmaxf(v[j].w, thr);
          p[i] = v[j];
        }
      }
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

