// whisper_common.cuh -- geometry, staging primitives, pass-1 / pass-2 / mel building blocks shared by the two
// Whisper kernels (included inside the anonymous namespace of b200mel.cu).
#pragma once

// ------------------------------------------------------------------------------------------------
// Whisper geometry
// ------------------------------------------------------------------------------------------------
constexpr int W_NFFT = 400, W_HOP = 160, W_NMEL = 80, W_NSAMP = 480000, W_NFRAME = 3000;
constexpr int W_TILE = 64;                                       // frames per CTA tile: TWO per lane (packed f32x2)
constexpr int W_TILES_PER_CLIP = (W_NFRAME + W_TILE - 1) / W_TILE;   // 47
constexpr int W_THREADS = 512;
constexpr int W_WARPS = W_THREADS / 32;                          // 16 = number of pass-1 tasks
constexpr int W_P2_TASKS = 13;                                   // warps 0..12 run pass 2, warps 13..15 prefetch audio
constexpr int W_ROWS = ((W_TILE - 1) * W_HOP + W_NFFT + W_HOP - 1) / W_HOP;   // 66 rows of 160 samples span one tile
// Audio tile layout: row r holds samples [160 r, 160 r + 164) of the tile at a pitch of 164 words, written by
// ONE TMA box per tile (TMA is 16-byte granular on both sides, so an odd pitch is not available; cp.async
// at 4-byte granularity costs ~8 LSU cycles per warp instruction and was 30 % of the kernel).  With 16-byte
// aligned rows, 32 lanes reading the same sample of 32 different rows would hit only 8 banks, so pass 1
// gives a warp 8 frame pairs x 4 CONSECUTIVE tasks instead: task a -> a+1 moves the sample index by 25
// (= 1 mod 4), which spreads the four 8-lane groups over the four bank residues: conflict free.
constexpr int W_PITCH = W_HOP + 4;                               // 164
constexpr int W_SM_AUDIO = ((W_ROWS * W_PITCH + 31) / 32) * 32;  // floats
constexpr int W_TX_BYTES = W_ROWS * W_PITCH * 4;                 // bytes one TMA box delivers
constexpr int W_TMAP_X = 284;                                    // tensor-map extent of the sample axis (see the host code)
constexpr int W_EBLK = 26 * 32 + 8;                              // float2 per task block: 26 rows + 8 pad (pass-1 stores of two tasks in one half-warp land in different banks)
constexpr int W_SM_E = 16 * W_EBLK * 2;                          // floats
constexpr int W_SM_TAB = 2 * 16 * 28;                            // pass-1 offsets (int) + window taps (float)
constexpr int W_PROWS = 13 * 16;                                 // power rows: k2 * 16 + k1
constexpr int W_SM_P = W_PROWS * 32 * 2;                         // floats (float2 per lane and row)
constexpr int W_SMEM_BYTES = (W_SM_AUDIO + W_SM_E + W_SM_P + W_SM_TAB) * 4 + 16;   // + the TMA mbarrier
static_assert(W_SMEM_BYTES <= 227 * 1024, "Whisper tile does not fit in shared memory");
constexpr int W_LANE2 = 8 * W_PITCH;                             // float offset of a lane's second frame (8 frames on)

// y = (log10(e) + 4) / 4 = log2(e) * (log10(2)/4) + 1; e >= 1e-10 so the ftz approx form is exact enough
__device__ __forceinline__ float w_norm_log(float e) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e));
  return __fmaf_rn(l, 0.07525749891599529f, 1.0f);
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
}

// ---- mbarrier / TMA primitives -----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra W_DONE_%=;\n"
      "bra W_WAIT_%=;\n"
      "W_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(float* smem_dst, const CUtensorMap* tmap, int x, int y, int z,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* tmap, int x, int y, int z) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(x), "r"(y), "r"(z) : "memory");
}

// ---- stage one tile of audio into shared memory ------------------------------------------------
struct WTile {
  const float* src;     // clip base
  int clip, f0;
  int L;                // valid samples (<= 480000)
  bool tma;             // interior tile: fetched by TMA; otherwise the generic path below
  bool silent;          // every sample the tile touches lies in the zero padding past the clip (32-frame kernel only)
};

__device__ __forceinline__ WTile w_tile(const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                                        int tile, int use_tma) {
  WTile t;
  t.clip = tile / W_TILES_PER_CLIP;
  t.f0 = (tile - t.clip * W_TILES_PER_CLIP) * W_TILE;
  const long long len_ll = lengths ? (long long)__ldg(lengths + t.clip) : stride;
  t.L = (int)(len_ll < 0 ? 0 : (len_ll > W_NSAMP ? W_NSAMP : len_ll));
  t.src = wave + (size_t)t.clip * (size_t)stride;
  const long long g0 = (long long)t.f0 * W_HOP - W_NFFT / 2;
  // every sample of the tile is real audio (no reflection, no zero fill), and the 4 words of row slack the
  // boxes also fetch stay inside the clip's row of the buffer
  t.tma = use_tma && g0 >= 0 && g0 + W_ROWS * W_HOP <= t.L && g0 + W_ROWS * W_HOP + 4 <= stride;
  return t;
}

// Interior tiles: one TMA box of 66 rows x 164 samples.  The tensor map views the audio as
// [clip][hop index y][x < 284] with a y-stride of 160 samples (overlapping rows), so row r of the tile is
// (x = 120, y = f0 - 2 + r).  Issued by one thread.
__device__ __forceinline__ void w_stage_tma(const WTile& t, const CUtensorMap* tmap, float* s_audio, unsigned long long* bar) {
  fence_proxy_async();               // earlier generic-proxy accesses to the tile vs. the async-proxy writes
  mbar_arrive_expect_tx(bar, W_TX_BYTES);
  tma_load_3d(s_audio, tmap, 120, t.f0 - 2, t.clip, bar);
}

// Tiles touching a clip edge: the same layout written with ordinary stores, applying the reflect padding of
// the 480000-sample padded clip and the zero fill past the clip length.  `part`/`nparts` split the rows.
__device__ __forceinline__ void w_stage_generic(const WTile& t, float* __restrict__ s_audio, int part, int nparts, int lane) {
  const int g0 = t.f0 * W_HOP - W_NFFT / 2;
  for (int r = part; r < W_ROWS; r += nparts) {
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

// ---- pass 1: windowed real 25-point DFT of residue class a --------------------------------------
// A warp works on 8 frame pairs x 4 consecutive classes: lane = (g, i), class a = 4 q + g, frames
// 16 fg + i and 16 fg + 8 + i packed as a float2 (q = warp & 3, fg = warp >> 2).  Good-Thomas input order
// and window taps come from shared-memory tables, one row per class (one code body for all 16 classes: a
// fully specialised variant was instruction-cache bound, profiles/r01_v1); the 8 lanes of a group read the
// same 16 bytes, so a table load is 4 wavefronts.
__device__ __forceinline__ void w_pass1(int a, const float* __restrict__ audio_lane, float2* __restrict__ e_dst,
                                        const int* __restrict__ s_off, const float* __restrict__ s_win) {
  float2 x[25], o[25];
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
  e_dst[0] = o[0];                        // X0 is real: row 1 (its imaginary part) is never read
#pragma unroll
  for (int c = 1; c < 25; ++c) e_dst[(c + 1) * 32] = o[c];
}

// ---- pass 2: complex 16-point DFT for k2 (warp-uniform, runtime); |X|^2 written back in place ----
__device__ __forceinline__ void w_pass2(int k2, const float2* __restrict__ e_lane, float2* __restrict__ p_lane) {
  if (k2 == 0) {                              // the pass-1 outputs for k2 = 0 are real: half the work
    float2 y[16], P[9];
#pragma unroll
    for (int a = 0; a < 16; ++a) y[a] = e_lane[a * W_EBLK];
    b2::real_dft16_power(y, P);
#pragma unroll
    for (int k1 = 0; k1 < 9; ++k1) p_lane[k1 * 32] = P[k1];   // |X[16-k1]| = |X[k1]|: rows 0..8 cover the task
    return;
  }
  float2 yr[16], yi[16], Xr[16], Xi[16];
  const float2* base = e_lane + k2 * 64;      // rows a*26 + 2*k2 (re) and a*26 + 2*k2 + 1 (im)
  float2* dst = p_lane + k2 * (16 * 32);      // power rows k2*16 + k1
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    yr[a] = base[a * W_EBLK];
    yi[a] = base[a * W_EBLK + 32];
  }
  b2::cplx_dft16(yr, yi, Xr, Xi);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) dst[k1 * 32] = b2::vfma(Xr[k1], Xr[k1], b2::vmul(Xi[k1], Xi[k1]));
}

// ---- mel: filters are specialised at compile time per warp (row offsets and weights are immediates) ----
// Each warp owns a CONTIGUOUS run of filters, balanced by cost (taps + a fixed per-filter epilogue).  Neighbouring
// triangles overlap by half, so the warp first loads the union of its bins once (about half the loads of a
// filter-by-filter walk) and then runs the independent accumulation chains side by side.
B2_CX int w_mel_len(int m) { const int t[80] = kWMelLen_INIT; return t[m]; }
B2_CX int w_mel_off(int m) { const int t[80] = kWMelOff_INIT; return t[m]; }
B2_CX int w_mel_start(int m) { const int t[80] = kWMelStart_INIT; return t[m]; }
B2_CX float w_mel_wt(int i) { const float t[B200MEL_W_NNZ] = kWMelW_INIT; return t[i]; }
// power-buffer row of FFT bin k: pass-2 task k2 = k mod 25 (mirrored to <= 12) leaves bin k in row k2*16 + k1
B2_CX int w_bin_row(int k) {
  int k1 = k % 16, k2 = k % 25;
  if (k2 > 12) { const int kk = 400 - k; k1 = kk % 16; k2 = kk % 25; }
  if (k2 == 0 && k1 > 8) k1 = 16 - k1;        // real task: only k1 = 0..8 are stored
  return k2 * 16 + k1;
}
constexpr int W_MEL_FIXED_COST = 8;
B2_CX int w_mel_total_cost() { int c = 0; for (int m = 0; m < 80; ++m) c += w_mel_len(m) + W_MEL_FIXED_COST; return c; }
// first filter of share w out of ns (w = ns -> 80): the cumulative cost is cut into ns equal shares
B2_CX int w_mel_first(int w, int ns = 16) {
  if (w >= ns) return 80;
  const int total = w_mel_total_cost();
  int c = 0;
  for (int m = 0; m < 80; ++m) {
    if (c * ns >= w * total) return m;
    c += w_mel_len(m) + W_MEL_FIXED_COST;
  }
  return 80;
}

template <int J, int LEN, int OFF, int REL, int NB>
__device__ __forceinline__ void w_mel_taps(const float2 (&pb)[NB], float2& acc) {
  if constexpr (J < LEN) {
    constexpr float wt = w_mel_wt(OFF + J);
    acc = (J == 0) ? b2::vmulc(pb[REL + J], wt) : b2::vfmac(pb[REL + J], wt, acc);
    w_mel_taps<J + 1, LEN, OFF, REL, NB>(pb, acc);
  }
}

template <int M, int FE, int BLO, int NB>
__device__ __forceinline__ void w_mel_filters(const float2 (&pb)[NB], float* __restrict__ out_col,
                                              bool valid0, bool valid1, float& emax) {
  if constexpr (M < FE) {
    float2 acc;
    w_mel_taps<0, w_mel_len(M), w_mel_off(M), w_mel_start(M) - BLO, NB>(pb, acc);
    const float e0 = fmaxf(acc.x, 1e-10f), e1 = fmaxf(acc.y, 1e-10f);
    emax = fmaxf(emax, fmaxf(valid0 ? e0 : 0.0f, valid1 ? e1 : 0.0f));
    const float y0 = w_norm_log(e0), y1 = w_norm_log(e1);
    if (valid0) out_col[(size_t)M * W_NFRAME] = y0;
    if (valid1) out_col[(size_t)M * W_NFRAME + 8] = y1;
    w_mel_filters<M + 1, FE, BLO, NB>(pb, out_col, valid0, valid1, emax);
  }
}

template <int W>
__device__ __forceinline__ void w_mel_warp(const float2* __restrict__ p_lane, float* __restrict__ out_col,
                                           bool valid0, bool valid1, float& emax) {
  constexpr int FB = w_mel_first(W), FE = w_mel_first(W + 1);
  static_assert(FE > FB, "every warp needs at least one filter");
  constexpr int BLO = w_mel_start(FB), BHI = w_mel_start(FE - 1) + w_mel_len(FE - 1);
  constexpr int NB = BHI - BLO;
  float2 pb[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) pb[k] = p_lane[w_bin_row(BLO + k) * 32];
  w_mel_filters<FB, FE, BLO, NB>(pb, out_col, valid0, valid1, emax);
}

// ---- mel + log + per-clip max for one tile whose power spectrum sits in P ------------------------
__device__ __forceinline__ void w_mel_phase(int warp, int lane, int clip, int f0, const float2* __restrict__ s_p,
                                            float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  // column `lane` of E / P carries frames 16 (lane / 8) + lane % 8 and that + 8 (see w_pass1)
  const int frame0 = f0 + 16 * (lane >> 3) + (lane & 7), frame1 = frame0 + 8;
  const bool valid0 = frame0 < W_NFRAME, valid1 = frame1 < W_NFRAME;
  const float2* pl = s_p + lane;
  float* out_col = out + (size_t)clip * (W_NMEL * W_NFRAME) + frame0;
  float emax = 0.0f;
  switch (warp) {
    case 0: w_mel_warp<0>(pl, out_col, valid0, valid1, emax); break;
    case 1: w_mel_warp<1>(pl, out_col, valid0, valid1, emax); break;
    case 2: w_mel_warp<2>(pl, out_col, valid0, valid1, emax); break;
    case 3: w_mel_warp<3>(pl, out_col, valid0, valid1, emax); break;
    case 4: w_mel_warp<4>(pl, out_col, valid0, valid1, emax); break;
    case 5: w_mel_warp<5>(pl, out_col, valid0, valid1, emax); break;
    case 6: w_mel_warp<6>(pl, out_col, valid0, valid1, emax); break;
    case 7: w_mel_warp<7>(pl, out_col, valid0, valid1, emax); break;
    case 8: w_mel_warp<8>(pl, out_col, valid0, valid1, emax); break;
    case 9: w_mel_warp<9>(pl, out_col, valid0, valid1, emax); break;
    case 10: w_mel_warp<10>(pl, out_col, valid0, valid1, emax); break;
    case 11: w_mel_warp<11>(pl, out_col, valid0, valid1, emax); break;
    case 12: w_mel_warp<12>(pl, out_col, valid0, valid1, emax); break;
    case 13: w_mel_warp<13>(pl, out_col, valid0, valid1, emax); break;
    case 14: w_mel_warp<14>(pl, out_col, valid0, valid1, emax); break;
    default: w_mel_warp<15>(pl, out_col, valid0, valid1, emax); break;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  // positive floats order like their bit patterns; the slot is zeroed before the launch
  if (lane == 0 && emax > 0.0f) atomicMax(clip_max_bits + clip, __float_as_uint(emax));
}

