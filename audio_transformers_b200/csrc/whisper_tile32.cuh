// whisper_tile32.cuh -- the Whisper log-mel kernel: 32-frame tiles, one 512-thread CTA per SM running two halves.
#pragma once
// ================================================================================================
// Replaces HF:models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features) plus the pad / trim
// of HF:feature_extraction_sequence_utils.py:263-278,327-332, up to the clip-wide floor (whisper_post.cuh).
//
// ONE 512-thread CTA per SM runs TWO independent halves (8 warps each, own audio / E / P buffers of 103 KB, own named
// barrier and mbarriers): while one half waits at its barrier or on shared-memory loads the other computes.  Of a half's
// two phase boundaries one is a block barrier (E complete before pass 2) and the other an mbarrier with one arrival per
// warp ("phase B is over"), which a warp that starts with pass 1 only waits for when it has its DFT in registers and
// wants to store it: the warps with little pass-2 work run ahead and carry more of the mel stage instead.  A half
// works on 32-frame tiles; every FP32 value is a packed float2 holding the same quantity of two frames:
//   pass 1   a warp is 8 frame pairs x 4 residue classes: windowed real 25-point DFTs (Good-Thomas 16 x 25);
//   pass 2   the 16-point DFT is the same code for every k2 (no twiddles), so a warp takes 16 columns x 2 tasks;
//   mel      a warp takes the 16 columns unpacked: lanes 0..15 the first frame of each pair, lanes 16..31 the
//            second, scalar FFMA with immediate weights (the FMA pipe time per frame is unchanged).
// ================================================================================================
constexpr int V_TILE = 32, V_WARPS = 8;                                          // per half: 8 warps, one 32-frame tile in flight
constexpr int V_HALVES = 2, V_HALF_THREADS = V_WARPS * 32, V_THREADS = V_HALVES * V_HALF_THREADS;
constexpr int V_TILES_PER_CLIP = (W_NFRAME + V_TILE - 1) / V_TILE;               // 94
#ifndef V_NS
#define V_NS 16
#endif
#ifndef V_S6
#define V_S6 5
#endif
constexpr int V_MEL_NS = V_NS, V_MEL_S6 = V_S6;                                  // mel shares in all / of the role with the real pass-2 task (v_mel_phase)
constexpr int V_ROWS = ((V_TILE - 1) * W_HOP + W_NFFT + W_HOP - 1) / W_HOP;      // 34
constexpr int V_COLS = V_TILE / 2;                                               // 16 float2 columns
constexpr int V_SM_AUDIO = ((V_ROWS * W_PITCH + 31) / 32) * 32;                  // floats
constexpr int V_TX_BYTES = V_ROWS * W_PITCH * 4;
constexpr int V_EBLK = 26 * V_COLS + 8;                                          // float2 per class block (+8: two classes of a half-warp store to different banks)
constexpr int V_SM_E = 16 * V_EBLK * 2;                                          // floats
constexpr int V_SM_P = W_PROWS * V_COLS * 2;                                     // floats
constexpr int V_HALF_FLOATS = V_SM_AUDIO + V_SM_E + V_SM_P;                        // one half's audio | E | P
static_assert((V_HALF_FLOATS * 4) % 128 == 0, "the second half's audio tile must stay 128-byte aligned for TMA");
constexpr int V_SMEM_BYTES = (V_HALVES * V_HALF_FLOATS + W_SM_TAB) * 4 + 128;       // + two mbarriers, the tile counter, the tile descriptor rings
static_assert(V_SMEM_BYTES <= 227 * 1024, "both halves must fit in one SM");

// Does the 32-frame tile starting at frame f0 of a clip with L valid samples see only zero padding?  The smallest
// clip index any of its rows maps to (left reflection reaches index 0; right reflection maps g >= 480000 to
// 959998 - g) is already past the clip.
__device__ __forceinline__ bool v_tile_silent(int f0, int L) {
  const long long g0 = (long long)f0 * W_HOP - W_NFFT / 2, gend = g0 + V_ROWS * W_HOP;
  long long jmin = g0 < 0 ? 0 : g0;
  if (gend > W_NSAMP) { const long long r = 2LL * (W_NSAMP - 1) - (gend - 1); jmin = r < jmin ? r : jmin; }
  return jmin >= L;
}

__device__ __forceinline__ WTile v_tile(const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                                        int tile, int use_tma) {
  WTile t;
  t.clip = tile / V_TILES_PER_CLIP;
  t.f0 = (tile - t.clip * V_TILES_PER_CLIP) * V_TILE;
  const long long len_ll = lengths ? (long long)__ldg(lengths + t.clip) : stride;
  t.L = (int)(len_ll < 0 ? 0 : (len_ll > W_NSAMP ? W_NSAMP : len_ll));
  t.src = wave + (size_t)t.clip * (size_t)stride;
  const long long g0 = (long long)t.f0 * W_HOP - W_NFFT / 2, gend = g0 + V_ROWS * W_HOP;
  t.tma = use_tma && g0 >= 0 && gend <= t.L && gend + 4 <= stride;
  // A tile that sees only zero padding has exactly the floor value in every feature: neither the audio copy nor the
  // FFT is needed, and the floor pass writes its features (it knows the final clip maximum).  Whisper inputs are
  // mostly much shorter than the 30 s they are padded to, so for real batches this is the common tile.
  t.silent = v_tile_silent(t.f0, t.L);
  return t;
}

// The 32-frame kernel carries its mel energies scaled by 1e4 (the scale is folded into the filter weights, which are
// immediates): (log10(e) + 4) / 4 = log10(1e4 e) / 4 is then a single multiply after lg2, with no "+ 1" whose
// constant the compiler re-materialised with a MOV for every value.  The workspace slots and the floor pass use the
// same scaled unit (V_EFLOOR is the reference's 1e-10 clamp).
constexpr float V_ESCALE = 1e4f, V_EFLOOR = 1e-10f * V_ESCALE;
__device__ __forceinline__ float v_norm_log(float e_scaled) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e_scaled));
  return l * 0.07525749891599529f;
}

// Workspace of this kernel: one float2 per (clip, tile, warp) = {largest, smallest} mel energy that warp saw in that
// tile.  Every slot is written exactly once per launch, so the workspace needs no zeroing (no memset node, no atomics);
// the clip-floor pass reduces a clip's 94 x 8 slots itself: the maximum gives the floor, and a clip whose minimum is
// not below it has nothing to clamp and is not read at all.
constexpr int V_SLOTS_PER_CLIP = V_TILES_PER_CLIP * V_WARPS;
__device__ __forceinline__ float2* v_slot(float2* __restrict__ tile_max, int clip, int f0, int warp) {
  return tile_max + (size_t)clip * V_SLOTS_PER_CLIP + (f0 / V_TILE) * V_WARPS + warp;
}

// A tile of pure zero padding: its features are written by the floor pass; here only its maximum (mel = 0 ->
// max(., 1e-10)) is recorded so that an all-silent clip still has a defined clip maximum.  Called by ONE thread.
__device__ __forceinline__ void v_record_silent(const WTile& t, float2* __restrict__ tile_max) {
#pragma unroll
  for (int w = 0; w < V_WARPS; ++w) *v_slot(tile_max, t.clip, t.f0, w) = make_float2(V_EFLOOR, V_EFLOOR);
}

__device__ __forceinline__ void v_stage_generic(const WTile& t, float* __restrict__ s_audio, int part, int nparts, int lane) {
  const int g0 = t.f0 * W_HOP - W_NFFT / 2;
  for (int r = part; r < V_ROWS; r += nparts) {
    float* d = s_audio + r * W_PITCH + lane;
    const int gs = g0 + r * W_HOP + lane;
#pragma unroll
    for (int k = 0; k < W_HOP / 32; ++k) {
      const int g = gs + 32 * k;
      const int j = g < 0 ? -g : (g >= W_NSAMP ? 2 * (W_NSAMP - 1) - g : g);
      d[32 * k] = (j >= 0 && j < t.L) ? __ldg(t.src + j) : 0.0f;
    }
  }
}

// pass 1: windowed real 25-point DFT of residue class a for one frame pair per lane (E layout: 16 columns per row).
// Loads + butterflies and the stores are separate so that a warp can compute before E is free (v_run).
__device__ __forceinline__ void v_pass1_compute(int a, const float* __restrict__ audio_lane, const int* __restrict__ s_off,
                                                const float* __restrict__ s_win, float2 (&o)[25]) {
  float2 x[25];
  int off[28];
  float w[28];
  const int4* off4 = reinterpret_cast<const int4*>(s_off + a * 28);
  const float4* win4 = reinterpret_cast<const float4*>(s_win + a * 28);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const int4 v = off4[q];
    off[4 * q] = v.x; off[4 * q + 1] = v.y; off[4 * q + 2] = v.z; off[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const float4 v = win4[q];
    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int b = 0; b < 25; ++b) {
    const float* p = audio_lane + off[b];
    x[b] = make_float2(p[0], p[W_LANE2]);
  }
  b2::real_dft25(x, w, o);
}
__device__ __forceinline__ void v_pass1_store(float2* __restrict__ e_dst, const float2 (&o)[25]) {
  e_dst[0] = o[0];
#pragma unroll
  for (int c = 1; c < 25; ++c) e_dst[(c + 1) * V_COLS] = o[c];
}
__device__ __forceinline__ void v_pass1(int a, const float* __restrict__ audio_lane, float2* __restrict__ e_dst,
                                        const int* __restrict__ s_off, const float* __restrict__ s_win) {
  float2 o[25];
  v_pass1_compute(a, audio_lane, s_off, s_win, o);
  v_pass1_store(e_dst, o);
}

// pass 2 for one (k2, column) per lane; k2 >= 1
__device__ __forceinline__ void v_pass2(int k2, const float2* __restrict__ e_col, float2* __restrict__ p_col) {
  float2 yr[16], yi[16], Xr[16], Xi[16];
  const float2* base = e_col + k2 * (2 * V_COLS);
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    yr[a] = base[a * V_EBLK];
    yi[a] = base[a * V_EBLK + V_COLS];
  }
  b2::cplx_dft16(yr, yi, Xr, Xi);
  float2* dst = p_col + k2 * (16 * V_COLS);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) dst[k1 * V_COLS] = b2::vfma(Xr[k1], Xr[k1], b2::vmul(Xi[k1], Xi[k1]));
}

__device__ __forceinline__ void v_pass2_real(const float2* __restrict__ e_col, float2* __restrict__ p_col) {
  float2 y[16], P[9];
#pragma unroll
  for (int a = 0; a < 16; ++a) y[a] = e_col[a * V_EBLK];
  b2::real_dft16_power(y, P);
#pragma unroll
  for (int k1 = 0; k1 < 9; ++k1) p_col[k1 * V_COLS] = P[k1];
}

// mel, one frame per lane (scalar): p_lane points at this lane's float inside row 0 of P, rows are 32 floats apart
template <int J, int LEN, int OFF, int REL, int NB>
__device__ __forceinline__ void v_mel_taps(const float (&pb)[NB], float& acc) {
  if constexpr (J < LEN) {
    constexpr float wt = w_mel_wt(OFF + J) * V_ESCALE;
    acc = (J == 0) ? pb[REL + J] * wt : __fmaf_rn(pb[REL + J], wt, acc);
    v_mel_taps<J + 1, LEN, OFF, REL, NB>(pb, acc);
  }
}

template <int M, int FE, int BLO, int NB>
__device__ __forceinline__ void v_mel_filters(const float (&pb)[NB], float* __restrict__ out_col, bool valid, float& emax, float& emin) {
  if constexpr (M < FE) {
    float acc;
    v_mel_taps<0, w_mel_len(M), w_mel_off(M), w_mel_start(M) - BLO, NB>(pb, acc);
    // No clamp at the reference's 1e-10 floor here: the floor pass raises every value to max(clip max - 8 decades,
    // floor) anyway (an exact zero gives -inf for the moment).  Lanes past frame 3000 are masked once, in v_mel_phase.
    emax = fmaxf(emax, acc);
    emin = fminf(emin, acc);
    const float y = v_norm_log(acc);
    if (valid) out_col[(size_t)M * W_NFRAME] = y;
    v_mel_filters<M + 1, FE, BLO, NB>(pb, out_col, valid, emax, emin);
  }
}

template <int S, int NS>
__device__ __forceinline__ void v_mel_share(const float* __restrict__ p_lane, float* __restrict__ out_col, bool valid, float& emax, float& emin) {
  constexpr int FB = w_mel_first(S, NS), FE = w_mel_first(S + 1, NS);
  static_assert(FE > FB, "every share needs at least one filter");
  constexpr int BLO = w_mel_start(FB), BHI = w_mel_start(FE - 1) + w_mel_len(FE - 1);
  constexpr int NB = BHI - BLO;
  float pb[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) pb[k] = p_lane[w_bin_row(BLO + k) * (2 * V_COLS)];
  v_mel_filters<FB, FE, BLO, NB>(pb, out_col, valid, emax, emin);
}

__device__ __forceinline__ void v_mel_phase(int role, int warp, int lane, int clip, int f0, const float2* __restrict__ s_p,
                                            float* __restrict__ out, float2* __restrict__ tile_max) {
  // lane l < 16: first frame of column l; lane l >= 16: second frame (8 frames later) of column l - 16
  const int col = lane & 15, half = lane >> 4;
  const int frame = f0 + 16 * (col >> 3) + (col & 7) + 8 * half;
  const bool valid = frame < W_NFRAME;
  const float* pl = reinterpret_cast<const float*>(s_p) + 2 * col + half;
  float* out_col = out + (size_t)clip * (W_NMEL * W_NFRAME) + frame;
  float emax = 0.0f, emin = 3.0e38f;
  // The 80 filters are cut into V_MEL_NS shares of equal cost (w_mel_first).  Warps 0..5 carry a full pass-2 task each and
  // take ONE share; warp 6 (half a pass-2 task) takes V_MEL_S6 and warp 7 (none: it stages and draws) the rest -- those
  // two reach the end of phase B early, start their pass-1 task before the others have left pass 2 (v_run) and so have
  // the time: measured at batch 512, kernel 0.494 ms with two shares per warp, 0.480 (2,..,2,4,5), 0.470 (2,..,2,7,8),
  // 0.468 with this split (1,..,1,5,5), 0.475 (1,..,1,7,7).
  static_assert(V_MEL_S6 >= 1 && V_MEL_S6 <= 8 && V_MEL_NS > 6 + V_MEL_S6 && V_MEL_NS <= 6 + V_MEL_S6 + 8,
                "every share must belong to exactly one role: six single shares, V_MEL_S6 for role 6, at most eight for role 7");
#define V_MS(s) if constexpr ((s) < V_MEL_NS) v_mel_share<((s) < V_MEL_NS ? (s) : 0), V_MEL_NS>(pl, out_col, valid, emax, emin);
#define V_M6(k) if constexpr ((k) < V_MEL_S6) { V_MS(6 + (k)) }
#define V_M7(k) V_MS(6 + V_MEL_S6 + (k))
  switch (role) {
    case 0: V_MS(0) break;
    case 1: V_MS(1) break;
    case 2: V_MS(2) break;
    case 3: V_MS(3) break;
    case 4: V_MS(4) break;
    case 5: V_MS(5) break;
    case 6: V_M6(0) V_M6(1) V_M6(2) V_M6(3) V_M6(4) V_M6(5) V_M6(6) V_M6(7) break;
    default: V_M7(0) V_M7(1) V_M7(2) V_M7(3) V_M7(4) V_M7(5) V_M7(6) V_M7(7) break;
  }
#undef V_MS
#undef V_M6
#undef V_M7
  if (!valid) { emax = 0.0f; emin = 3.0e38f; }
  // energies are >= +0, so their bit patterns order like the values: one REDUX each instead of five shuffle rounds
  emax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(emax)));
  emin = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(emin)));
  if (lane == 0) *v_slot(tile_max, clip, f0, warp) = make_float2(emax, emin);
}

#ifndef V_L2_PREFETCH
#define V_L2_PREFETCH 1
#endif
#ifndef V_ROTATE
#define V_ROTATE 1      // the second half shifts its roles by two warps: its two light-pass-2 / heavy-mel warps then sit on the
                        // schedulers that carry four ordinary warps (64-clip step 78.1 -> 76.4 us)
#endif
// which warps run their mel share before their pass-1 task: three of warps 0..5, the other three in the other half (the
// halves then tend to be in complementary parts of phase A: 1.00 M vs 0.965 M clips/s with the same choice in both);
// warps 6 and 7 always start with pass 1, which they may begin before the phase barrier (see v_run)
#ifndef V_MEL_FIRST
#define V_MEL_FIRST(w) ((w) < 6 && ((((w) >= 3) ? 1 : 0) ^ half) == 0)
#endif

// One 512-thread CTA per SM runs TWO independent halves (8 warps each, own audio / E / P buffers, own named barrier
// and mbarrier): while one half waits at its barrier or on shared-memory loads the other computes.  The halves draw
// tiles from a counter in shared memory, so they finish together whatever share of the issue slots each one gets
// (as two separate CTAs with a static split, the CTA the hardware favoured finished 26 % early and left its SM
// half empty for the rest of the kernel).  CTA c owns tiles c, c + gridDim, c + 2 gridDim, ...
template <int DUMMY>
__device__ __forceinline__ void v_run(const CUtensorMap* tmap, int use_tma, const float* __restrict__ wave, long long stride,
                                      const int* __restrict__ lengths, int batch, float* __restrict__ out,
                                      float2* __restrict__ tile_max, float* smem) {
  const int half = threadIdx.x >> 8, tid = threadIdx.x & (V_HALF_THREADS - 1), lane = tid & 31, warp = tid >> 5;
  float* s_audio = smem + half * V_HALF_FLOATS;
  float2* s_e = reinterpret_cast<float2*>(s_audio + V_SM_AUDIO);
  float2* s_p = reinterpret_cast<float2*>(s_audio + V_SM_AUDIO + V_SM_E);
  const int* s_off = reinterpret_cast<const int*>(smem + V_HALVES * V_HALF_FLOATS);
  const float* s_win = smem + V_HALVES * V_HALF_FLOATS + 16 * 28;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + V_HALVES * V_HALF_FLOATS + W_SM_TAB) + half;
  int* s_ctl = reinterpret_cast<int*>(smem + V_HALVES * V_HALF_FLOATS + W_SM_TAB) + 4;   // tile counter
  // two-slot ring of drawn tiles per half: {tile, clip, first frame, valid samples | TMA flag << 30}
  int4* s_desc = reinterpret_cast<int4*>(smem + V_HALVES * V_HALF_FLOATS + W_SM_TAB + 8) + 2 * half;
  const int ntiles = batch * V_TILES_PER_CLIP;
  auto half_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(V_HALF_THREADS) : "memory"); };

  // pass-1 role: class a = 4 (warp & 3) + lane / 8, frame pair (16 fg + i, 16 fg + 8 + i) = column 8 fg + i, fg = warp >> 2
  const int p1_a = 4 * (warp & 3) + (lane >> 3);
  const float* audio_lane = s_audio + (16 * (warp >> 2) + (lane & 7)) * W_PITCH;
  float2* p1_dst = s_e + p1_a * V_EBLK + 8 * (warp >> 2) + (lane & 7);
  // The ROLE of a warp decides its pass-2 task, its mel shares and whether it stages: roles 0..5 take the tasks
  // k2 = 1 + 2 role + lane / 16 and one mel share each; role 6 the real task k2 = 0 (lanes 0..15) and V_MEL_S6 shares;
  // role 7 no pass-2 task (it issues the copies and draws tiles) and the remaining shares.  Roles 6 and 7 are the ones
  // that run ahead into the next tile's pass 1.  In the first half role = warp; the second half rotates by two warps, so
  // that its roles 6 and 7 (warps 4 and 5) sit on other schedulers than the first half's.
  const int p2_warp = (warp + (V_ROTATE ? 2 * half : 0)) & 7;                 // the role
  const int p2_col = lane & 15;
  const int p2_k2 = 1 + 2 * p2_warp + (lane >> 4);
  const bool mel_first = V_MEL_FIRST(p2_warp);
  const int STAGE_TID = ((7 - (V_ROTATE ? 2 * half : 0)) & 7) * 32;      // lane 0 of the warp whose role has no pass-2 task

  // Draw tiles from the shared counter until one has audio in it (tiles of pure zero padding only get their maximum
  // recorded); for that one, pull its TMA box into L2 already.  Called by ONE thread of the half, which publishes
  // the descriptor through shared memory: nobody else divides, reads the clip length or classifies the tile.
  auto draw = [&]() -> int4 {
    for (;;) {
      const int t = blockIdx.x + atomicAdd(s_ctl, 1) * gridDim.x;
      if (t >= ntiles) return make_int4(ntiles, 0, 0, 0);
      const WTile wt = v_tile(wave, stride, lengths, t, use_tma);
      if (!wt.silent) {
#if V_L2_PREFETCH
        if (wt.tma) tma_prefetch_l2_3d(tmap, 120, wt.f0 - 2, wt.clip);
#endif
        return make_int4(t, wt.clip, wt.f0, wt.L | (wt.tma ? 0x40000000 : 0));
      }
      v_record_silent(wt, tile_max);
    }
  };
  auto unpack = [&](const int4& d) -> WTile {
    WTile wt;
    wt.clip = d.y; wt.f0 = d.z; wt.L = d.w & 0x3fffffff; wt.tma = (d.w >> 30) & 1; wt.silent = false;
    wt.src = wave + (size_t)d.y * (size_t)stride;
    return wt;
  };
  auto stage = [&](const WTile& wt) -> bool {
    if (wt.tma) {
      if (tid == STAGE_TID) {
        fence_proxy_async();
        mbar_arrive_expect_tx(s_bar, V_TX_BYTES);
        tma_load_3d(s_audio, tmap, 120, wt.f0 - 2, wt.clip, s_bar);
      }
    } else {
      v_stage_generic(wt, s_audio, warp, V_WARPS, lane);
    }
    return wt.tma;
  };

  if (tid == STAGE_TID) { s_desc[0] = draw(); s_desc[1] = draw(); }
  half_sync();
  int4 cur = s_desc[0];
  unsigned tma_parity = 0;
  bool cur_tma = false;
  if (cur.x < ntiles) cur_tma = stage(unpack(cur));
  int prev_clip = -1, prev_f0 = 0;
  // "every warp of the half has left phase B": an mbarrier with one arrival per warp instead of a block barrier, so that a
  // warp may load and transform its pass-1 inputs (the audio tile is complete once the TMA barrier says so) while
  // slower warps are still in pass 2; only its E stores and the mel stage wait
  unsigned long long* s_done = reinterpret_cast<unsigned long long*>(smem + V_HALVES * V_HALF_FLOATS + W_SM_TAB) + 12 + half;
  __syncwarp();
  if (lane == 0) mbar_arrive(s_done);

#pragma unroll 1
  for (int it = 0;; ++it) {
    const bool have = cur.x < ntiles;
    if (have && cur_tma) { mbar_wait(s_bar, tma_parity); tma_parity ^= 1u; }
    // one call site per stage (the instruction cache does not hold the tile loop twice)
    const bool early = have && cur_tma && !mel_first;
    if (!early) mbar_wait(s_done, (unsigned)(it & 1));       // audio (ordinary stores) visible; P complete; E free
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
      if ((step == 0) == mel_first) {
        if (prev_clip >= 0) v_mel_phase(p2_warp, warp, lane, prev_clip, prev_f0, s_p, out, tile_max);
      } else if (have) {
        float2 o[25];
        v_pass1_compute(p1_a, audio_lane, s_off, s_win, o);
        if (early) mbar_wait(s_done, (unsigned)(it & 1));    // P(previous tile) complete; E is free
        v_pass1_store(p1_dst, o);
      }
    }
    if (!have) break;
    half_sync();                           // E complete; the audio tile and P are dead from here on

    // ---- phase B: copy of the next tile (TMA, or plain stores at a clip edge) + pass 2(this tile) -------
    // The tile after the next one is drawn here, by a thread of the warp that has no pass-2 task, into the ring slot
    // this tile's descriptor occupied (last read one trip ago, two barriers back).
    const int4 next = s_desc[(it + 1) & 1];
    cur_tma = (next.x < ntiles) ? stage(unpack(next)) : false;
    if (tid == STAGE_TID) s_desc[it & 1] = draw();
    if (p2_warp < 6) v_pass2(p2_k2, s_e + p2_col, s_p + p2_col);
    else if (p2_warp == 6 && lane < 16) v_pass2_real(s_e + p2_col, s_p + p2_col);
    prev_clip = cur.y;
    prev_f0 = cur.z;
    cur = next;
    __syncwarp();
    if (lane == 0) mbar_arrive(s_done);
  }
}

__global__ void __launch_bounds__(V_THREADS, 1)
whisper_logmel_kernel32(const __grid_constant__ CUtensorMap tmap, int use_tma,
                        const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                        int batch, float* __restrict__ out, float2* __restrict__ tile_max) {
  extern __shared__ __align__(1024) float smem[];
  {
    int* s_off = reinterpret_cast<int*>(smem + V_HALVES * V_HALF_FLOATS);
    float* s_win = smem + V_HALVES * V_HALF_FLOATS + 16 * 28;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + V_HALVES * V_HALF_FLOATS + W_SM_TAB);
    int* s_ctl = reinterpret_cast<int*>(s_bar) + 4;
    for (int i = threadIdx.x; i < 16 * 28; i += V_THREADS) { s_off[i] = c_wp1_off[i]; s_win[i] = c_wp1_win[i]; }
    if (threadIdx.x == 0) {
      mbar_init(s_bar, 1); mbar_init(s_bar + 1, 1);
      mbar_init(s_bar + 12, V_WARPS); mbar_init(s_bar + 13, V_WARPS);
      fence_proxy_async(); s_ctl[0] = 0;
    }
  }
  __syncthreads();
  // The kernel is launched with programmatic stream serialisation: it may become resident and run its prologue
  // (tables, mbarriers: nothing that touches global memory) while the previous kernel of the stream -- the clip-floor
  // pass of the previous call, or whatever produced the audio -- is still finishing.  Everything after this wait sees
  // that kernel's results; nothing before it reads or writes global memory.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  v_run<0>(&tmap, use_tma, wave, stride, lengths, batch, out, tile_max, smem);
}
