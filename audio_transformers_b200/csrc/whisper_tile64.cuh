// whisper_tile64.cuh -- the 64-frame, one-CTA-per-SM cut of the Whisper kernel (kept behind B200MEL_KERNEL64=1).
#pragma once
#ifdef W_TRACE
__device__ long long* g_trace = nullptr;     // [iter][4 marks][16 warps] clock64 of CTA 0 (debug builds only)
#define W_MARK(k) do { if (blockIdx.x == 0 && lane == 0 && it < 32 && g_trace) g_trace[(it * 4 + (k)) * 16 + warp] = clock64(); } while (0)
#else
#define W_MARK(k) do { } while (0)
#endif

#ifndef W_MEL_FIRST_MASK
#define W_MEL_FIRST_MASK 0x0f0f            // warps (bit set) that run their mel share before their pass-1 task
#endif

// Persistent CTA, one per SM, looping over (clip, 64-frame tile).  Two block barriers per tile:
//   phase A   mel(previous tile, from P)  +  pass 1(this tile, audio -> E)     [LSU-heavy + FMA-heavy work
//             run side by side: half of the warps do their mel share first, the other half their DFT task]
//   phase B   TMA prefetch of the next tile's audio (one box, issued by one thread, lands on an mbarrier)
//             +  pass 2(this tile, E -> P) on warps 0..12
__global__ void __launch_bounds__(W_THREADS, 1)
whisper_logmel_kernel(const __grid_constant__ CUtensorMap tmap, int use_tma,
                      const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                      int batch, float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  extern __shared__ __align__(1024) float smem[];
  float* s_audio = smem;
  float2* s_e = reinterpret_cast<float2*>(smem + W_SM_AUDIO);
  float2* s_p = reinterpret_cast<float2*>(smem + W_SM_AUDIO + W_SM_E);
  int* s_off = reinterpret_cast<int*>(smem + W_SM_AUDIO + W_SM_E + W_SM_P);
  float* s_win = smem + W_SM_AUDIO + W_SM_E + W_SM_P + 16 * 28;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + W_SM_AUDIO + W_SM_E + W_SM_P + W_SM_TAB);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = batch * W_TILES_PER_CLIP;
  if (tid < 16 * 28) { s_off[tid] = c_wp1_off[tid]; s_win[tid] = c_wp1_win[tid]; }
  if (tid == 0) { mbar_init(s_bar, 1); fence_proxy_async(); }
  __syncthreads();

  // pass-1 role of this lane: class a, frame pair (16 fg + i, 16 fg + 8 + i) = column 8 fg + i of E
  const int p1_a = 4 * (warp & 3) + (lane >> 3);
  const int p1_col = 8 * (warp >> 2) + (lane & 7);
  const float* audio_lane = s_audio + (16 * (warp >> 2) + (lane & 7)) * W_PITCH;
  float2* p1_dst = s_e + p1_a * W_EBLK + p1_col;
  const bool mel_first = (W_MEL_FIRST_MASK >> warp) & 1;
  constexpr int STAGE_WARP = W_P2_TASKS;     // first warp without a pass-2 task

  int tile = blockIdx.x;
  unsigned tma_parity = 0;
  bool cur_tma = false;
  if (tile < ntiles) {
    const WTile t = w_tile(wave, stride, lengths, tile, use_tma);
    cur_tma = t.tma;
    if (t.tma) { if (tid == STAGE_WARP * 32) w_stage_tma(t, &tmap, s_audio, s_bar); }
    else w_stage_generic(t, s_audio, warp, W_WARPS, lane);
  }
  int prev_clip = -1, prev_f0 = 0;

  for (int it = 0;; tile += gridDim.x, ++it) {
    const bool have = tile < ntiles;
    if (have && cur_tma) { mbar_wait(s_bar, tma_parity); tma_parity ^= 1u; }
    __syncthreads();                       // audio(tile) visible; P(previous tile) complete; E is free
    W_MARK(0);

    // ---- phase A ------------------------------------------------------------------------------------
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
      if ((step == 0) == mel_first) {
        if (prev_clip >= 0) w_mel_phase(warp, lane, prev_clip, prev_f0, s_p, out, clip_max_bits);
      } else if (have) {
        w_pass1(p1_a, audio_lane, p1_dst, s_off, s_win);
      }
    }
    W_MARK(1);
    if (!have) break;
    __syncthreads();                       // E complete; the audio tile and P are dead from here on

    // ---- phase B ------------------------------------------------------------------------------------
    {
      const int next = tile + gridDim.x;
      cur_tma = false;
      if (next < ntiles) {
        const WTile t = w_tile(wave, stride, lengths, next, use_tma);
        cur_tma = t.tma;
        if (t.tma) { if (tid == STAGE_WARP * 32) w_stage_tma(t, &tmap, s_audio, s_bar); }
        else w_stage_generic(t, s_audio, warp, W_WARPS, lane);
      }
    }
    if (warp < W_P2_TASKS) w_pass2(warp, s_e + lane, s_p + lane);
    W_MARK(2);
    prev_clip = tile / W_TILES_PER_CLIP;
    prev_f0 = (tile - prev_clip * W_TILES_PER_CLIP) * W_TILE;
  }
}

