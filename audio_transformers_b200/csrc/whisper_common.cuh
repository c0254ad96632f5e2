// whisper_common.cuh -- geometry, mbarrier / TMA primitives and the compile-time mel tables of the Whisper kernel
// (included inside the anonymous namespace of b200mel.cu).
#pragma once

// ------------------------------------------------------------------------------------------------
// Whisper geometry
// ------------------------------------------------------------------------------------------------
constexpr int W_NFFT = 400, W_HOP = 160, W_NMEL = 80, W_NSAMP = 480000, W_NFRAME = 3000;
// Audio tile layout: row r holds samples [160 r, 160 r + 164) of the tile at a pitch of 164 words, written by
// ONE TMA box per tile (TMA is 16-byte granular on both sides, so an odd pitch is not available; cp.async
// at 4-byte granularity costs ~8 LSU cycles per warp instruction and was 30 % of the kernel).  With 16-byte
// aligned rows, 32 lanes reading the same sample of 32 different rows would hit only 8 banks, so pass 1
// gives a warp 8 frame pairs x 4 CONSECUTIVE tasks instead: task a -> a+1 moves the sample index by 25
// (= 1 mod 4), which spreads the four 8-lane groups over the four bank residues: conflict free.
constexpr int W_PITCH = W_HOP + 4;                               // 164
constexpr int W_TMAP_X = 284;                                    // tensor-map extent of the sample axis (see the host code)
constexpr int W_SM_TAB = 2 * 16 * 28;                            // pass-1 offsets (int) + window taps (float)
constexpr int W_PROWS = 13 * 16;                                 // power rows: k2 * 16 + k1
constexpr int W_LANE2 = 8 * W_PITCH;                             // float offset of a lane's second frame (8 frames on)

// y = (log10(e) + 4) / 4 = log2(e) * (log10(2)/4) + 1; e >= 1e-10 so the ftz approx form is exact enough
__device__ __forceinline__ float w_norm_log(float e) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e));
  return __fmaf_rn(l, 0.07525749891599529f, 1.0f);
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  bool silent;          // every sample the tile touches lies in the zero padding past the clip
};

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

