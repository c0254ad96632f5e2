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
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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

// ---- stage one tile of audio into shared memory ------------------------------------------------
struct WTile {
  const float* src;     // clip base
  int clip, f0;
  int L;                // valid samples (<= 480000)
  bool tma;             // interior tile: fetched by TMA; otherwise the generic path below
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
#if W_P1_STREAM
  // The first stage is five independent 5-point transforms on inputs {r, r+5, .., r+20}.  Their audio loads are
  // issued two groups ahead of the arithmetic instead of all up front, so that the shared-memory traffic of
  // the 16 warps is spread over the phase instead of arriving as one burst at its start.
  float2 Y0[5], Y1r[5], U1[5], Y2r[5], U2[5];
  auto load_group = [&](int r) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float* p = audio_lane + off[r + 5 * j];
      x[r + 5 * j] = make_float2(p[0], p[W_LANE2]);
    }
  };
  load_group(0);
  load_group(1);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    b2::rdft5w(x[r], x[r + 5], x[r + 10], x[r + 15], x[r + 20], w[r], w[r + 5], w[r + 10], w[r + 15], w[r + 20],
               Y0[r], Y1r[r], U1[r], Y2r[r], U2[r]);
    if (r + 2 < 5) {
      asm volatile("" : "+f"(Y0[r].x) :: "memory");   // pins the next loads behind this group's arithmetic
      load_group(r + 2);
    }
  }
  b2::real_dft25_stage2(Y0, Y1r, U1, Y2r, U2, o);
#else
#pragma unroll
  for (int b = 0; b < 25; ++b) {
    const float* p = audio_lane + off[b];
    x[b] = make_float2(p[0], p[W_LANE2]);
  }
  b2::real_dft25(x, w, o);
#endif
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
// first filter of warp w (w = 16 -> 80): the cumulative cost is cut into 16 equal shares
B2_CX int w_mel_first(int w) {
  if (w >= 16) return 80;
  const int total = w_mel_total_cost();
  int c = 0;
  for (int m = 0; m < 80; ++m) {
    if (c * 16 >= w * total) return m;
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

#ifdef W_TRACE
__device__ long long* g_trace = nullptr;     // [iter][4 marks][16 warps] clock64 of CTA 0 (debug builds only)
#define W_MARK(k) do { if (blockIdx.x == 0 && lane == 0 && it < 32 && g_trace) g_trace[(it * 4 + (k)) * 16 + warp] = clock64(); } while (0)
#else
#define W_MARK(k) do { } while (0)
#endif

#ifndef W_P1_STREAM
#define W_P1_STREAM 1
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
#ifdef W_EXP_STAGGER_B
    { const long long t_start = clock64(); const int wait = (warp >> 2) * W_EXP_STAGGER_B;
      while (clock64() - t_start < wait) { } }
#endif
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

// ================================================================================================
// Whisper kernel, 32-frame tiles, TWO CTAs per SM
//
// Same arithmetic and the same three stages as whisper_logmel_kernel, re-cut so that a tile needs 107 KB of
// shared memory instead of 208 KB: two independent 256-thread CTAs share an SM, and while one waits at a block
// barrier or on its shared-memory loads the other one computes.  (The 64-frame kernel is latency bound: 16 warps,
// all in the same phase.)  What makes the half-size tile possible without giving up the packed f32x2 arithmetic:
//   pass 1   a warp is 8 frame pairs x 4 classes anyway (w_pass1), so a 32-frame tile is simply 2 x 4 warp tasks;
//   pass 2   the 16-point DFT is the same code for every k2 (no twiddles), so a warp takes 16 columns x 2 tasks;
//   mel      a warp takes the 16 columns unpacked: lanes 0..15 the first frame of each pair, lanes 16..31 the
//            second, scalar FFMA with the same immediates (the FMA pipe time per frame is unchanged).
// ================================================================================================
constexpr int V_TILE = 32, V_THREADS = 256, V_WARPS = 8;
constexpr int V_TILES_PER_CLIP = (W_NFRAME + V_TILE - 1) / V_TILE;               // 94
constexpr int V_ROWS = ((V_TILE - 1) * W_HOP + W_NFFT + W_HOP - 1) / W_HOP;      // 34
constexpr int V_COLS = V_TILE / 2;                                               // 16 float2 columns
constexpr int V_SM_AUDIO = ((V_ROWS * W_PITCH + 31) / 32) * 32;                  // floats
constexpr int V_TX_BYTES = V_ROWS * W_PITCH * 4;
constexpr int V_EBLK = 26 * V_COLS + 8;                                          // float2 per class block (+8: two classes of a half-warp store to different banks)
constexpr int V_SM_E = 16 * V_EBLK * 2;                                          // floats
constexpr int V_SM_P = W_PROWS * V_COLS * 2;                                     // floats
constexpr int V_SMEM_BYTES = (V_SM_AUDIO + V_SM_E + V_SM_P + W_SM_TAB) * 4 + 16;
static_assert(2 * (V_SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs must fit in one SM");

__device__ __forceinline__ WTile v_tile(const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                                        int tile, int use_tma) {
  WTile t;
  t.clip = tile / V_TILES_PER_CLIP;
  t.f0 = (tile - t.clip * V_TILES_PER_CLIP) * V_TILE;
  const long long len_ll = lengths ? (long long)__ldg(lengths + t.clip) : stride;
  t.L = (int)(len_ll < 0 ? 0 : (len_ll > W_NSAMP ? W_NSAMP : len_ll));
  t.src = wave + (size_t)t.clip * (size_t)stride;
  const long long g0 = (long long)t.f0 * W_HOP - W_NFFT / 2;
  t.tma = use_tma && g0 >= 0 && g0 + V_ROWS * W_HOP <= t.L && g0 + V_ROWS * W_HOP + 4 <= stride;
  return t;
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

// pass 1: identical to w_pass1 but for the E layout of this kernel (16 columns per row)
__device__ __forceinline__ void v_pass1(int a, const float* __restrict__ audio_lane, float2* __restrict__ e_dst,
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
  e_dst[0] = o[0];
#pragma unroll
  for (int c = 1; c < 25; ++c) e_dst[(c + 1) * V_COLS] = o[c];
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
    constexpr float wt = w_mel_wt(OFF + J);
    acc = (J == 0) ? pb[REL + J] * wt : __fmaf_rn(pb[REL + J], wt, acc);
    v_mel_taps<J + 1, LEN, OFF, REL, NB>(pb, acc);
  }
}

template <int M, int FE, int BLO, int NB>
__device__ __forceinline__ void v_mel_filters(const float (&pb)[NB], float* __restrict__ out_col, bool valid, float& emax) {
  if constexpr (M < FE) {
    float acc;
    v_mel_taps<0, w_mel_len(M), w_mel_off(M), w_mel_start(M) - BLO, NB>(pb, acc);
    const float e = fmaxf(acc, 1e-10f);
    emax = fmaxf(emax, valid ? e : 0.0f);
    const float y = w_norm_log(e);
    if (valid) out_col[(size_t)M * W_NFRAME] = y;
    v_mel_filters<M + 1, FE, BLO, NB>(pb, out_col, valid, emax);
  }
}

template <int S>
__device__ __forceinline__ void v_mel_share(const float* __restrict__ p_lane, float* __restrict__ out_col, bool valid, float& emax) {
  constexpr int FB = w_mel_first(S), FE = w_mel_first(S + 1);
  constexpr int BLO = w_mel_start(FB), BHI = w_mel_start(FE - 1) + w_mel_len(FE - 1);
  constexpr int NB = BHI - BLO;
  float pb[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) pb[k] = p_lane[w_bin_row(BLO + k) * (2 * V_COLS)];
  v_mel_filters<FB, FE, BLO, NB>(pb, out_col, valid, emax);
}

__device__ __forceinline__ void v_mel_phase(int warp, int lane, int clip, int f0, const float2* __restrict__ s_p,
                                            float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  // lane l < 16: first frame of column l; lane l >= 16: second frame (8 frames later) of column l - 16
  const int col = lane & 15, half = lane >> 4;
  const int frame = f0 + 16 * (col >> 3) + (col & 7) + 8 * half;
  const bool valid = frame < W_NFRAME;
  const float* pl = reinterpret_cast<const float*>(s_p) + 2 * col + half;
  float* out_col = out + (size_t)clip * (W_NMEL * W_NFRAME) + frame;
  float emax = 0.0f;
#define V_MEL_CASE(w) case w: v_mel_share<2 * w>(pl, out_col, valid, emax); v_mel_share<2 * w + 1>(pl, out_col, valid, emax); break;
  switch (warp) { V_MEL_CASE(0) V_MEL_CASE(1) V_MEL_CASE(2) V_MEL_CASE(3) V_MEL_CASE(4) V_MEL_CASE(5) V_MEL_CASE(6) default: V_MEL_CASE(7) }
#undef V_MEL_CASE
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if (lane == 0 && emax > 0.0f) atomicMax(clip_max_bits + clip, __float_as_uint(emax));
}

#ifndef V_ROTATE
#define V_ROTATE 1
#endif
#ifndef V_MEL_FIRST
#define V_MEL_FIRST(w) (((w) >> 2) & 1)
#endif

__global__ void __launch_bounds__(V_THREADS, 2)
whisper_logmel_kernel32(const __grid_constant__ CUtensorMap tmap, int use_tma,
                        const float* __restrict__ wave, long long stride, const int* __restrict__ lengths,
                        int batch, float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  extern __shared__ __align__(1024) float smem[];
  float* s_audio = smem;
  float2* s_e = reinterpret_cast<float2*>(smem + V_SM_AUDIO);
  float2* s_p = reinterpret_cast<float2*>(smem + V_SM_AUDIO + V_SM_E);
  int* s_off = reinterpret_cast<int*>(smem + V_SM_AUDIO + V_SM_E + V_SM_P);
  float* s_win = smem + V_SM_AUDIO + V_SM_E + V_SM_P + 16 * 28;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + V_SM_AUDIO + V_SM_E + V_SM_P + W_SM_TAB);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = batch * V_TILES_PER_CLIP;
  for (int i = tid; i < 16 * 28; i += V_THREADS) { s_off[i] = c_wp1_off[i]; s_win[i] = c_wp1_win[i]; }
  if (tid == 0) { mbar_init(s_bar, 1); fence_proxy_async(); }
  __syncthreads();

  // pass-1 role: class a = 4 (warp & 3) + lane / 8, frame pair (16 fg + i, 16 fg + 8 + i) = column 8 fg + i, fg = warp >> 2
  const int p1_a = 4 * (warp & 3) + (lane >> 3);
  const float* audio_lane = s_audio + (16 * (warp >> 2) + (lane & 7)) * W_PITCH;
  float2* p1_dst = s_e + p1_a * V_EBLK + 8 * (warp >> 2) + (lane & 7);
  // pass-2 role: warps 0..5 take tasks k2 = 1 + 2 warp + lane / 16, warp 6 the real task k2 = 0 (lanes 0..15), warp 7 none
  // The second CTA of an SM (CTAs are dealt round-robin, so blockIdx >= gridDim / 2) rotates the roles by two warps:
  // the light pass-2 warps (real task, idle) then sit on the schedulers that carry two full tasks in the first CTA.
  const int rot = (V_ROTATE && blockIdx.x >= (gridDim.x >> 1)) ? 2 : 0;
  const int p2_warp = (warp + rot) & 7;
  const int p2_col = lane & 15;
  const int p2_k2 = 1 + 2 * p2_warp + (lane >> 4);
  const bool mel_first = V_MEL_FIRST(warp);
  constexpr int STAGE_TID = 7 * 32;

  auto stage = [&](int t) -> bool {
    const WTile wt = v_tile(wave, stride, lengths, t, use_tma);
    if (wt.tma) {
      if (tid == STAGE_TID) {
        fence_proxy_async();
        mbar_arrive_expect_tx(s_bar, V_TX_BYTES);
        tma_load_3d(s_audio, &tmap, 120, wt.f0 - 2, wt.clip, s_bar);
      }
    } else {
      v_stage_generic(wt, s_audio, warp, V_WARPS, lane);
    }
    return wt.tma;
  };

  int tile = blockIdx.x;
  unsigned tma_parity = 0;
  bool cur_tma = false;
  if (tile < ntiles) cur_tma = stage(tile);
  int prev_clip = -1, prev_f0 = 0;

  for (;; tile += gridDim.x) {
    const bool have = tile < ntiles;
    if (have && cur_tma) { mbar_wait(s_bar, tma_parity); tma_parity ^= 1u; }
    __syncthreads();                       // audio(tile) visible; P(previous tile) complete; E is free

    // ---- phase A: mel(previous tile) + pass 1(this tile) ----------------------------------------------
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
      if ((step == 0) == mel_first) {
        if (prev_clip >= 0) v_mel_phase(warp, lane, prev_clip, prev_f0, s_p, out, clip_max_bits);
      } else if (have) {
        v_pass1(p1_a, audio_lane, p1_dst, s_off, s_win);
      }
    }
    if (!have) break;
    __syncthreads();                       // E complete; the audio tile and P are dead from here on

    // ---- phase B: TMA prefetch of the next tile + pass 2(this tile) -----------------------------------
    {
      const int next = tile + gridDim.x;
      cur_tma = (next < ntiles) ? stage(next) : false;
    }
    if (p2_warp < 6) v_pass2(p2_k2, s_e + p2_col, s_p + p2_col);
    else if (p2_warp == 6 && lane < 16) v_pass2_real(s_e + p2_col, s_p + p2_col);
    prev_clip = tile / V_TILES_PER_CLIP;
    prev_f0 = (tile - prev_clip * V_TILES_PER_CLIP) * V_TILE;
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
constexpr int U_TILE = 32, U_THREADS = 256, U_WARPS = U_THREADS / 32;
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

// cuTensorMapEncodeTiled, resolved through the runtime so that the library does not link libcuda directly
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct b200mel_handle {
  int device;
  int preset;
  int sm_count;
  tmap_encode_fn encode = nullptr;   // Whisper preset: TMA descriptor encoder
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
    e = cudaFuncSetAttribute(urban_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U_SMEM_BYTES);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
  b200mel_handle* h = new b200mel_handle(device, preset, prop.multiProcessorCount);
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
  if ((long long)batch * V_TILES_PER_CLIP > 0x7fffffffLL) return fail(B200MEL_ERR_BAD_ARG, "whisper_logmel: batch too large for one launch");
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned int* clip_max = (unsigned int*)workspace;
  cudaError_t e = cudaMemsetAsync(clip_max, 0, (size_t)batch * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
  // TMA view of the audio: [clip][y][x], element (x, y, clip) = wave[clip * stride + 160 y + x], x < 284.  Rows
  // overlap (y-stride 160 samples < 284), which is what lets a box start at any sample with 16-byte aligned
  // strides.  NY is chosen so that every in-bounds element lies inside its clip's row of the buffer.
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int use_tma = 0;
  if (stride_samples >= W_TMAP_X) {
    const cuuint64_t dims[3] = {(cuuint64_t)W_TMAP_X, (cuuint64_t)((stride_samples - W_TMAP_X) / W_HOP + 1), (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)W_HOP * sizeof(float), (cuuint64_t)stride_samples * sizeof(float)};
    const bool k32 = getenv("B200MEL_KERNEL64") == nullptr;
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
  if (getenv("B200MEL_KERNEL64") == nullptr) {     // default: 32-frame tiles, two CTAs per SM
    const long long nt = (long long)batch * V_TILES_PER_CLIP;
    const int grid32 = nt < 2LL * h->sm_count ? (int)nt : 2 * h->sm_count;
    whisper_logmel_kernel32<<<grid32, V_THREADS, V_SMEM_BYTES, stream>>>(
        tmap, use_tma, wave, (long long)stride_samples, lengths, batch, out, clip_max);
  } else {
    whisper_logmel_kernel<<<grid_main, W_THREADS, W_SMEM_BYTES, stream>>>(
        tmap, use_tma, wave, (long long)stride_samples, lengths, batch, out, clip_max);
  }
  if (prof) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream); ++h->prof_n; }
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
  urban_mel_kernel<<<batch * tiles_per_clip, U_THREADS, U_SMEM_BYTES, stream>>>(
      wave, (long long)stride_samples, n_samples, n_frames, tiles_per_clip, batch, log_eps, out);
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
