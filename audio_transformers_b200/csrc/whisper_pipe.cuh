// whisper_pipe.cuh -- the Whisper log-mel kernel: one persistent, warp-specialised CTA per SM.
//
// Replaces HF:models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features) plus the pad / trim
// of HF:feature_extraction_sequence_utils.py:263-278,327-332 in ONE kernel, the clip-wide floor (:157-158) included.
//
// A tile is 32 frames of one clip (16 float2 columns: column c packs frames 16 (c >> 3) + (c & 7) and that + 8).
// Tiles stream through three audio slots and three E slots; twenty warps in five roles, each a loop of its own,
// coupled only by mbarriers and one 128-thread named barrier per tile and group:
//
//   control   (warp 16)      enumerates this CTA's tiles (blockIdx + k * gridDim), skips tiles of pure zero padding,
//                            publishes a descriptor and starts the audio copy (two 1-D bulk copies per tile, TMA),
//                            after pulling the tile two places further on into L2.
//   pass 1    (warps 0..7)   windowed real 25-point DFTs of the 16 Good-Thomas residue classes.  lane = (class a,
//                            half h), warp = frame slot.  Every lane keeps ITS 25 window taps and 25 shared-memory
//                            addresses in registers for the whole kernel (its class never changes): a task is
//                            50 LDS.32 + 203 packed FP + 25 STS.64, no table traffic.
//   pass 2 + mel (warps 8..11 | 12..15)  two groups that take alternate tiles, so that one is in its FMA-bound part
//                            while the other loads, stores or waits.  Complex 16-point DFTs over the classes with
//                            |X|^2 written IN PLACE over the E rows just read (a lane reads and writes only its own
//                            column), one barrier, then sparse mel + log straight from there.  The mel runs are cut
//                            so that the four warps of a group carry the same load although their pass-2 shares
//                            differ (2, 2, 1.5, 1 tasks).
//   commit    (warp 17)      after a tile's features are stored: ONE gpu-scope fence for everything stored since
//                            the last look, then the tile maxima and the tile counts (global atomics).
//   floor     (warps 18, 19) wait until a clip is complete, then apply max(y, ymax - 2) to this CTA's share of the
//                            clip while it still sits in L2, and write the features of the silent tiles.
//
// Audio tile in shared memory: two LINEAR sub-copies of 2800 samples (frames 0..15 and 16..31 of the tile); 2800 = 16
// (mod 32), so for one DFT input index the 16 classes x 2 halves of a warp fall on 32 different banks: 25 a + 16 h
// (mod 32) is a bijection.  E (pass-1 output, then power) [class][row][column] with an odd class pitch.
#pragma once

namespace xp {

constexpr int X_TILE = 32, X_COLS = 16;
constexpr int X_TILES_PER_CLIP = (W_NFRAME + X_TILE - 1) / X_TILE;          // 94
constexpr int X_P1_WARPS = 8;
constexpr int X_GROUPS = 2, X_GROUP_WARPS = 4, X_GROUP_THREADS = X_GROUP_WARPS * 32;
constexpr int X_WARP_B = X_P1_WARPS;                                        // first pass-2 warp
constexpr int X_WARP_AUX = X_WARP_B + X_GROUPS * X_GROUP_WARPS;             // first auxiliary warp
constexpr int X_THREADS = (X_WARP_AUX + 4) * 32;                            // 640
constexpr int X_SUB = 15 * W_HOP + W_NFFT;                                  // 2800 samples: 16 consecutive frames
constexpr int X_SPAN = 31 * W_HOP + W_NFFT;                                 // 5360 samples: the whole tile
constexpr int X_AUD_BYTES = 2 * X_SUB * 4;                                  // 22400
constexpr int X_EB = 25 * X_COLS + 1;                                       // float2 per class block (odd: conflict-free stores)
constexpr int X_E_BYTES = 16 * X_EB * 8;                                    // 51328
constexpr int X_NSLOT = 3;                                                  // audio slots = E slots: tile n uses slot n % 3
constexpr int X_OFF_AUD = 0;
constexpr int X_OFF_E = X_OFF_AUD + X_NSLOT * X_AUD_BYTES;
constexpr int X_OFF_CTL = X_OFF_E + X_NSLOT * X_E_BYTES;
constexpr int X_RING = 8;                                                   // tile descriptors in flight
constexpr int X_CRING = 4;                                                  // stored-but-uncommitted tiles per group
// control block: descriptors | per group: commit ring (clip, energy of 4 warps), committed count | mbarriers
constexpr int X_CTL_DESC = 0;                                               // int4[X_RING]
constexpr int X_CTL_CCLIP = X_CTL_DESC + X_RING * 16;                       // int[group][X_CRING]
constexpr int X_CTL_CMAX = X_CTL_CCLIP + X_GROUPS * X_CRING * 4;            // float[group][X_CRING][4]
constexpr int X_CTL_CNT = X_CTL_CMAX + X_GROUPS * X_CRING * 4 * 4;          // int[group] tiles committed, int[group] warps x tiles stored,
                                                                            // int: 1 + last clip committed, int: commit warp done
constexpr int X_CTL_BAR = X_CTL_CNT + 32;                                   // mbarriers (8 bytes each)
// mbarriers are indexed by tile number mod 6 = (slot, parity of the tile): the even and the odd tiles are two chains
// that advance independently, and a parity wait may only ever be one phase behind its barrier -- with one barrier per
// slot a chain that runs ahead would wait for the phase after next and be let through by the previous one.
constexpr int X_NBAR = 2 * X_NSLOT;
enum { XB_AUD_FULL = 0, XB_AUD_EMPTY = X_NBAR, XB_E_FULL = 2 * X_NBAR, XB_E_EMPTY = 3 * X_NBAR, XB_COUNT = 4 * X_NBAR };
constexpr int X_P1_GROUP_WARPS = X_P1_WARPS / X_GROUPS;                     // pass-1 warps per tile parity
// floor warps: 2 warps x 2 staging buffers for their bulk copies, and 4 mbarriers
constexpr int X_FSHARES = 371;                                              // a clip is cut into 371 shares of <= 162 float4
constexpr int X_FBUF_BYTES = ((W_NMEL * W_NFRAME / 4 + X_FSHARES - 1) / X_FSHARES) * 16;   // 2592
constexpr int X_OFF_FBAR = X_OFF_CTL + X_CTL_BAR + XB_COUNT * 8;
constexpr int X_OFF_FBUF = (X_OFF_FBAR + 4 * 8 + 15) / 16 * 16;
constexpr int X_SMEM_BYTES = X_OFF_FBUF + 4 * X_FBUF_BYTES;
static_assert(X_OFF_E % 16 == 0 && X_OFF_CTL % 16 == 0 && X_CTL_BAR % 8 == 0, "alignment");
static_assert(X_SMEM_BYTES <= 227 * 1024, "the buffers must fit in one SM");
static_assert(X_SUB % 32 == 16, "the two sub-copies must sit 16 banks apart");

// register budgets of the roles (setmaxnreg).  The kernel is launched with 96 registers per thread (64 K / 640,
// rounded down to the allocation unit); the roles can only redistribute what the CTA was given at launch.
constexpr int X_REGS_LAUNCH = 96, X_REGS_P1 = 112, X_REGS_B = 96, X_REGS_AUX = 64;
static_assert(256 * X_REGS_P1 + 256 * X_REGS_B + 128 * X_REGS_AUX <= X_THREADS * X_REGS_LAUNCH, "register pool of the CTA");

// ---- small PTX helpers -------------------------------------------------------------------------
template <int OFF>
__device__ __forceinline__ float x_lds(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void x_sts2(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void x_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void x_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void x_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "X_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra X_DONE_%=;\n"
      "bra X_WAIT_%=;\n"
      "X_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool x_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void x_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void x_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
template <int ID, int N>
__device__ __forceinline__ void x_bar() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
__device__ __forceinline__ void x_bar_dyn(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ int x_lds_acquire(unsigned a) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void x_sts_release(unsigned a, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void x_sts_add_release(unsigned a, int v) {
  asm volatile("red.release.cta.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long x_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ---- tile geometry --------------------------------------------------------------------------------
// Does the tile starting at frame f0 of a clip with L valid samples see only zero padding?  (The smallest clip index
// any of its samples maps to -- the left reflection reaches 0, the right one maps g >= 480000 to 959998 - g -- is
// already past the clip.)  Monotone in f0: the silent tiles of a clip are a suffix.
B2_CX bool x_tile_silent(int f0, int L) {
  const int g0 = f0 * W_HOP - W_NFFT / 2, gend = g0 + X_SPAN;
  int jmin = g0 < 0 ? 0 : g0;
  if (gend > W_NSAMP) { const int r = 2 * (W_NSAMP - 1) - (gend - 1); jmin = r < jmin ? r : jmin; }
  return jmin >= L;
}
// number of tiles of a clip that are NOT silent (they are a prefix).  Closed form of the predicate above: the right
// reflection never reaches below the tile's first sample (it would need g0 > 477319, the last tile starts at 475960),
// so a tile t >= 1 is silent iff 5120 t - 200 >= L, and tile 0 iff L == 0.
B2_CX int x_live_tiles(int L) {
  if (L <= 0) return 0;
  const int t = (L + W_NFFT / 2 + X_TILE * W_HOP - 1) / (X_TILE * W_HOP);
  return t < X_TILES_PER_CLIP ? t : X_TILES_PER_CLIP;
}
B2_CX bool x_live_tiles_agree(int L) {
  for (int t = 0; t < X_TILES_PER_CLIP; ++t)
    if (x_tile_silent(t * X_TILE, L) != (t >= x_live_tiles(L))) return false;
  return true;
}
static_assert(x_live_tiles_agree(0) && x_live_tiles_agree(1) && x_live_tiles_agree(159) && x_live_tiles_agree(4919) &&
              x_live_tiles_agree(4920) && x_live_tiles_agree(4921) && x_live_tiles_agree(10040) && x_live_tiles_agree(10041) &&
              x_live_tiles_agree(240000) && x_live_tiles_agree(475959) && x_live_tiles_agree(475960) &&
              x_live_tiles_agree(475961) && x_live_tiles_agree(479999) && x_live_tiles_agree(480000),
              "closed form of the silent-tile count");
__device__ __forceinline__ int x_clip_len(const int* __restrict__ lengths, long long stride, int clip) {
  long long len = lengths ? (long long)__ldg(lengths + clip) : stride;
  len = len < 0 ? 0 : len;
  len = len > stride ? stride : len;                         // never read past the clip's row
  return (int)(len > W_NSAMP ? W_NSAMP : len);
}

// energies are carried scaled by 1e4 (folded into the immediate filter weights): (log10 e + 4) / 4 = lg2(1e4 e) * c
constexpr float X_ESCALE = 1e4f, X_EFLOOR = 1e-10f * X_ESCALE;
__device__ __forceinline__ float x_norm_log(float e_scaled) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e_scaled));
  return l * 0.07525749891599529f;
}

// ---- pass 1 -----------------------------------------------------------------------------------------
template <int BASE, int... C>
__device__ __forceinline__ void x_store_rows(unsigned a, const float2 (&o)[25], std::integer_sequence<int, C...>) {
  (x_sts2<BASE + C * X_COLS * 8>(a, o[C]), ...);
}
__device__ __forceinline__ void x_pass1_load(const unsigned (&ax)[25], unsigned soff, float2 (&x)[25]) {
#pragma unroll
  for (int b = 0; b < 25; ++b) {
    const unsigned aa = ax[b] + soff;                  // the slot rotates through three: one add per sample pair
    x[b] = make_float2(x_lds<0>(aa), x_lds<8 * W_HOP * 4>(aa));
  }
}

// tiles at a clip edge (reflection at the ends of the padded 30 s buffer, zero fill past the clip): the 128 pass-1
// threads that own the tile write the two sub-copies with ordinary stores
__device__ __forceinline__ void x_stage_generic(const float* __restrict__ src, int L, int f0, float* __restrict__ aud, int tid) {
  const int s0 = f0 * W_HOP - W_NFFT / 2;
#pragma unroll 4
  for (int e = tid; e < 2 * X_SUB; e += X_P1_GROUP_WARPS * 32) {
    const int hh = e >= X_SUB ? 1 : 0;
    const int g = s0 + e - hh * (X_SUB - 16 * W_HOP);
    const int j = g < 0 ? -g : (g >= W_NSAMP ? 2 * (W_NSAMP - 1) - g : g);
    aud[e] = (j >= 0 && j < L) ? __ldg(src + j) : 0.0f;
  }
}

// ---- pass 2 -----------------------------------------------------------------------------------------
// E rows of class a: row 0 = X0 (real), rows 2 k2 - 1 / 2 k2 = Re / Im of X_k2, k2 = 1..12.  The power of bin
// (k1, k2) replaces the real-part entry of class a = k1: a lane only ever touches its own column, and every entry
// it overwrites is one it has read, so no other warp or lane can observe the change too early.
__device__ __forceinline__ void x_pass2(int k2, float2* __restrict__ e_col) {
  float2 yr[16], yi[16], Xr[16], Xi[16];
  float2* base = e_col + (2 * k2 - 1) * X_COLS;
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    yr[a] = base[a * X_EB];
    yi[a] = base[a * X_EB + X_COLS];
  }
  b2::cplx_dft16(yr, yi, Xr, Xi);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) base[k1 * X_EB] = b2::vfma(Xr[k1], Xr[k1], b2::vmul(Xi[k1], Xi[k1]));
}
__device__ __forceinline__ void x_pass2_real(float2* __restrict__ e_col) {
  float2 y[16], P[9];
#pragma unroll
  for (int a = 0; a < 16; ++a) y[a] = e_col[a * X_EB];
  b2::real_dft16_power(y, P);
#pragma unroll
  for (int k1 = 0; k1 < 9; ++k1) e_col[k1 * X_EB] = P[k1];     // |X[16 - k1]| = |X[k1]|: k1 = 0..8 cover the task
}
// float2 offset (inside an E slot, before the column) of the power of FFT bin k
B2_CX int x_bin_off(int k) {
  int k1 = k % 16, k2 = k % 25;
  if (k2 > 12) { const int kk = 400 - k; k1 = kk % 16; k2 = kk % 25; }
  if (k2 == 0 && k1 > 8) k1 = 16 - k1;
  return k1 * X_EB + (k2 == 0 ? 0 : 2 * k2 - 1) * X_COLS;
}

// ---- mel ------------------------------------------------------------------------------------------------
// One frame per lane (lanes 0..15 the first frame of a column, 16..31 the second), scalar FFMA with immediate
// weights.  A role owns a contiguous run of filters, cut so that pass 2 + mel cost about the same on the four warps of
// a group; a run is walked in chunks whose union of bins fits a register array.
constexpr int X_MEL_MAXBINS = 24;
B2_CX int x_mel_cost(int m) { return (w_mel_len(m) * 13 + 4) / 8 + 4; }       // taps + their share of the bin loads + epilogue
// pass-2 share of a role: k2 = {1..4}, {5..8}, {9, 10} + the real task k2 = 0, {11, 12}
B2_CX int x_p2_cost(int role) { return role < 2 ? 448 : (role == 2 ? 329 : 224); }
B2_CX int x_role_first(int role) {
  if (role <= 0) return 0;
  if (role >= X_GROUP_WARPS) return 80;
  int total = 0;
  for (int m = 0; m < 80; ++m) total += x_mel_cost(m);
  for (int r = 0; r < X_GROUP_WARPS; ++r) total += x_p2_cost(r);
  // cumulative mel budget of roles 0..role-1 (a role's share is what its pass-2 task leaves of a quarter of the total)
  int budget = 0;
  for (int r = 0; r < role; ++r) { const int share = total - X_GROUP_WARPS * x_p2_cost(r); budget += share > 0 ? share : 0; }
  int c = 0;
  for (int m = 0; m < 80; ++m) {
    if (X_GROUP_WARPS * c >= budget) return m;
    c += x_mel_cost(m);
  }
  return 80;
}
B2_CX int x_chunk_end(int fb, int fe) {
  int e = fb + 1;
  while (e < fe && w_mel_start(e) + w_mel_len(e) - w_mel_start(fb) <= X_MEL_MAXBINS) ++e;
  return e;
}

template <int J, int LEN, int OFF, int REL, int NB>
__device__ __forceinline__ void x_mel_taps(const float (&pb)[NB], float& acc) {
  if constexpr (J < LEN) {
    constexpr float wt = w_mel_wt(OFF + J) * X_ESCALE;
    acc = (J == 0) ? pb[REL + J] * wt : __fmaf_rn(pb[REL + J], wt, acc);
    x_mel_taps<J + 1, LEN, OFF, REL, NB>(pb, acc);
  }
}
template <int M, int FE, int BLO, int NB>
__device__ __forceinline__ void x_mel_filters(const float (&pb)[NB], float* __restrict__ out_col, bool valid, float& emax) {
  if constexpr (M < FE) {
    float acc;
    x_mel_taps<0, w_mel_len(M), w_mel_off(M), w_mel_start(M) - BLO, NB>(pb, acc);
    // no clamp at the reference's 1e-10 here: the floor warps raise every value to max(clip max - 8 decades, floor)
    emax = fmaxf(emax, acc);
    const float y = x_norm_log(acc);
    if (valid) out_col[(size_t)M * W_NFRAME] = y;
    x_mel_filters<M + 1, FE, BLO, NB>(pb, out_col, valid, emax);
  }
}
template <int FB, int FE>
__device__ __forceinline__ void x_mel_run(const float* __restrict__ p_lane, float* __restrict__ out_col, bool valid, float& emax) {
  if constexpr (FB < FE) {
    constexpr int CE = x_chunk_end(FB, FE);
    constexpr int BLO = w_mel_start(FB), BHI = w_mel_start(CE - 1) + w_mel_len(CE - 1);
    constexpr int NB = BHI - BLO;
    float pb[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) pb[k] = p_lane[2 * x_bin_off(BLO + k)];
    x_mel_filters<FB, CE, BLO, NB>(pb, out_col, valid, emax);
    x_mel_run<CE, FE>(p_lane, out_col, valid, emax);
  }
}

// returns the largest (scaled) mel energy this warp saw in the tile (lanes past frame 3000 excluded)
template <int R>
__device__ __forceinline__ float x_mel_role(int lane, int clip, int f0, const float2* __restrict__ s_p, float* __restrict__ out) {
  const int col = lane & 15, half = lane >> 4;
  const int frame = f0 + 16 * (col >> 3) + (col & 7) + 8 * half;
  const bool valid = frame < W_NFRAME;
  const float* pl = reinterpret_cast<const float*>(s_p) + 2 * col + half;
  float* out_col = out + (size_t)clip * (W_NMEL * W_NFRAME) + frame;
  float emax = 0.0f;
  x_mel_run<x_role_first(R), x_role_first(R + 1)>(pl, out_col, valid, emax);
  if (!valid) emax = 0.0f;
  // energies are >= +0, so their bit patterns order like the values: one REDUX instead of five shuffle rounds
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(emax)));
}

// Workspace: one 64-bit word per clip, {low: tiles stored, high: bits of the largest mel energy}, and one more word
// counting the CTAs that are done.  Zero before the first use; every launch leaves it zero again (the last CTA to
// finish clears the clip words, the CTA counter wraps by itself).
struct XArgs {
  const float* wave;
  long long stride;
  const int* lengths;
  int batch;
  float* out;
  unsigned long long* ws;
  const float* win400;      // global-memory copy of the window (lane-dependent index: not a constant-bank read)
  int debug;                // development knock-outs (0 in production): 1 no commit / floor, 2 no mel, 4 no pass 2, 8 no pass-1 math
};

struct XCtx {               // addresses every role needs
  unsigned char* smem;
  unsigned sbase, bar0;
  __device__ __forceinline__ unsigned bar(unsigned i) const { return bar0 + 8u * i; }
  __device__ __forceinline__ int4* ring() const { return reinterpret_cast<int4*>(smem + X_OFF_CTL + X_CTL_DESC); }
  __device__ __forceinline__ int* c_clip(int g) const { return reinterpret_cast<int*>(smem + X_OFF_CTL + X_CTL_CCLIP) + g * X_CRING; }
  __device__ __forceinline__ float* c_max(int g) const { return reinterpret_cast<float*>(smem + X_OFF_CTL + X_CTL_CMAX) + g * X_CRING * 4; }
  __device__ __forceinline__ unsigned a_committed(int g) const { return sbase + X_OFF_CTL + X_CTL_CNT + 4u * (unsigned)g; }
  __device__ __forceinline__ unsigned a_stored(int g) const { return sbase + X_OFF_CTL + X_CTL_CNT + 8u + 4u * (unsigned)g; }
  __device__ __forceinline__ unsigned a_lastclip() const { return sbase + X_OFF_CTL + X_CTL_CNT + 16u; }
  __device__ __forceinline__ unsigned a_ended() const { return sbase + X_OFF_CTL + X_CTL_CNT + 20u; }
};

// =============================== pass 1 ==================================================================
// Warps 0..3 take the even tiles of the CTA's list, warps 4..7 the odd ones: the two warps a scheduler hosts then
// sit in different phases (one loads while the other is in its FMA-bound part).  A warp runs frame slots u and u + 4
// of its tile through ONE copy of the code.
__device__ __forceinline__ void x_role_pass1(const XArgs& A, const XCtx& C, int pg, int u, int lane) {
  const int a = lane & 15, h = lane >> 4;
  unsigned ax[25];
  float wn[25];
#pragma unroll
  for (int b = 0; b < 25; ++b) {
    const int n = (25 * a + 16 * b) % W_NFFT;
    ax[b] = C.sbase + X_OFF_AUD + 4u * (unsigned)(X_SUB * h + W_HOP * u + n);
    wn[b] = __ldg(A.win400 + n);          // a constant table: safe to read ahead of griddepcontrol.wait
  }
  // keep the 50 lane constants in registers: opaque to the optimiser, so they are neither re-derived nor re-loaded
#pragma unroll
  for (int b = 0; b < 25; ++b) asm volatile("" : "+r"(ax[b]), "+f"(wn[b]));
  unsigned e_addr = C.sbase + X_OFF_E + 8u * (unsigned)(a * X_EB + 8 * h + u);
  asm volatile("" : "+r"(e_addr));
  const int tid = (int)threadIdx.x - pg * (X_P1_GROUP_WARPS * 32);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (unsigned n = pg;; n += 2) {                     // tile n: buffers n % 3, barriers n % 6, phase parity (n / 6) & 1
    const unsigned i6 = n % 6u, par = (n / 6u) & 1u, slot = i6 >= X_NSLOT ? i6 - X_NSLOT : i6;
    const unsigned j6 = (n + 3u) % 6u, jpar = ((n - 3u) / 6u) & 1u;      // tile n - 3: the slot's previous user
    x_wait(C.bar(XB_AUD_FULL + i6), par);
    const int4 d = C.ring()[n & (X_RING - 1)];
    if (d.x < 0) {                                     // end marker of this parity: pass it on
      __syncwarp();
      if (lane == 0) x_arrive(C.bar(XB_E_FULL + i6));
      break;
    }
    if (!(d.w & 1)) {                                  // clip edge: ordinary loads and stores by the four warps
      x_stage_generic(A.wave + (size_t)d.x * (size_t)A.stride, d.z, d.y,
                      reinterpret_cast<float*>(C.smem + X_OFF_AUD + slot * X_AUD_BYTES), tid);
      x_bar_dyn(2 + pg, X_P1_GROUP_WARPS * 32);
    }
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {                      // frame slots u and u + 4
      float2 x[25], o[25];
      x_pass1_load(ax, slot * X_AUD_BYTES + t * (4 * W_HOP * 4), x);
      if (t == 1) {
        __syncwarp();
        if (lane == 0) x_arrive(C.bar(XB_AUD_EMPTY + i6));     // the samples are in registers: the slot is free
      }
      if (A.debug & 8) {
#pragma unroll
        for (int b = 0; b < 25; ++b) o[b] = b2::vmulc(x[b], wn[b]);
      } else {
        b2::real_dft25(x, wn, o);
      }
      if (t == 0 && n >= 3) x_wait(C.bar(XB_E_EMPTY + j6), jpar);   // the mel stage of tile n - 3 has read the slot
      x_store_rows<0>(e_addr + slot * X_E_BYTES + t * (4 * 8), o, std::make_integer_sequence<int, 25>{});
    }
    __syncwarp();
    if (lane == 0) x_arrive(C.bar(XB_E_FULL + i6));
  }
}

// =============================== pass 2 + mel, group g, role R ==============================================
__device__ __forceinline__ void x_role_pass2(const XArgs& A, const XCtx& C, int g, const int R, int lane) {
  const int col = lane & 15, j = lane >> 4;
  int* c_clip = C.c_clip(g);
  float* c_max = C.c_max(g);
  // pass-2 share of the role: k2 = {1..4}, {5..8}, {9, 10} + the real task k2 = 0, {11, 12}; ONE copy of the codelet
  const int k2_first = 1 + 4 * R - (R == 3 ? 2 : 0) + j, n_full = R < 2 ? 2 : 1;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (unsigned k = 0;; ++k) {                         // tile n = 2 k + g: buffers n % 3, barriers n % 6
    const unsigned n = 2 * k + g, i6 = n % 6u, slot = i6 >= X_NSLOT ? i6 - X_NSLOT : i6;
    x_wait(C.bar(XB_E_FULL + i6), (n / 6u) & 1u);
    const int4 d = C.ring()[n & (X_RING - 1)];
    if (d.x < 0) {
      // (the barrier keeps a fast warp's count for the end marker behind the other warps' counts for the last tile)
      x_bar_dyn(4 + g, X_GROUP_THREADS);
      if (R == 0 && lane == 0) {
        if (k >= X_CRING) { while (x_lds_acquire(C.a_committed(g)) < (int)k - (X_CRING - 1)) __nanosleep(64); }
        c_clip[k & (X_CRING - 1)] = -1;
      }
      __syncwarp();
      if (lane == 0) x_sts_add_release(C.a_stored(g), 1);
      break;
    }
    float2* s_e = reinterpret_cast<float2*>(C.smem + X_OFF_E + slot * X_E_BYTES);
    if (!(A.debug & 4)) {
#pragma unroll 1
      for (int t = 0; t < n_full; ++t) x_pass2(k2_first + 2 * t, s_e + col);
      if (R == 2 && j == 0) x_pass2_real(s_e + col);
    }
    x_bar_dyn(4 + g, X_GROUP_THREADS);                 // every power of the tile is in place
    float m = 0.0f;
    if (!(A.debug & 2)) {
      switch (R) {
        case 0: m = x_mel_role<0>(lane, d.x, d.y, s_e, A.out); break;
        case 1: m = x_mel_role<1>(lane, d.x, d.y, s_e, A.out); break;
        case 2: m = x_mel_role<2>(lane, d.x, d.y, s_e, A.out); break;
        default: m = x_mel_role<3>(lane, d.x, d.y, s_e, A.out); break;
      }
    }
    if (lane == 0) {
      // the commit ring entry of this tile, last used by tile k - 4 of the group, must have been committed
      if (k >= X_CRING) { while (x_lds_acquire(C.a_committed(g)) < (int)k - (X_CRING - 1)) __nanosleep(64); }
      c_max[(k & (X_CRING - 1)) * 4 + R] = m;
      if (R == 0) c_clip[k & (X_CRING - 1)] = d.x;
    }
    __syncwarp();
    if (lane == 0) {
      x_arrive(C.bar(XB_E_EMPTY + i6));                // the slot goes back to pass 1
      x_sts_add_release(C.a_stored(g), 1);             // four of these: the features of the tile are stored
    }
  }
}

}  // namespace xp

__global__ void __launch_bounds__(xp::X_THREADS, 1)
whisper_logmel_pipe_kernel(const xp::XArgs A) {
  using namespace xp;
  extern __shared__ __align__(128) unsigned char x_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  XCtx C;
  C.smem = x_smem;
  C.sbase = smem_u32(x_smem);
  C.bar0 = C.sbase + X_OFF_CTL + X_CTL_BAR;

  if (threadIdx.x == 0) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(x_smem + X_OFF_CTL + X_CTL_BAR);
    for (int i = 0; i < X_NBAR; ++i) {
      mbar_init(b + XB_AUD_FULL + i, 1);
      mbar_init(b + XB_AUD_EMPTY + i, X_P1_GROUP_WARPS);
      mbar_init(b + XB_E_FULL + i, X_P1_GROUP_WARPS);
      mbar_init(b + XB_E_EMPTY + i, X_GROUP_WARPS);
    }
    int* cnt = reinterpret_cast<int*>(x_smem + X_OFF_CTL + X_CTL_CNT);
    for (int i = 0; i < 8; ++i) cnt[i] = 0;
    fence_proxy_async();
  }
  __syncthreads();

  if (warp < X_WARP_B) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(X_REGS_P1));
    x_role_pass1(A, C, warp / X_P1_GROUP_WARPS, warp % X_P1_GROUP_WARPS, lane);
  } else if (warp < X_WARP_AUX) {
    static_assert(X_REGS_B == X_REGS_LAUNCH, "pass 2 + mel keeps the launch allocation");
    // warp -> (group, role): the second group takes the roles in reverse order, so every scheduler hosts one heavy
    // and one light pass-2 share (warp w sits on scheduler w % 4)
    const int g = (warp - X_WARP_B) / X_GROUP_WARPS;
    const int r = (warp - X_WARP_B) % X_GROUP_WARPS;
    x_role_pass2(A, C, g, g == 0 ? r : X_GROUP_WARPS - 1 - r, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(X_REGS_AUX));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long ntiles = (long long)A.batch * X_TILES_PER_CLIP;
    if (warp == X_WARP_AUX) {
      // =============================== control ============================================================
      unsigned n = 0;                                      // tile n: buffers n % 3, barriers n % 6
      for (long long k0 = 0;; k0 += 32) {
        const long long t = (long long)blockIdx.x + (k0 + lane) * (long long)gridDim.x;
        const bool valid = t < ntiles;
        int clip = 0, f0 = 0, L = 0;
        bool live = false, tma = false;
        if (valid) {
          clip = (int)(t / X_TILES_PER_CLIP);
          f0 = (int)(t - (long long)clip * X_TILES_PER_CLIP) * X_TILE;
          L = x_clip_len(A.lengths, A.stride, clip);
          live = !x_tile_silent(f0, L);
          const int s0 = f0 * W_HOP - W_NFFT / 2;
          tma = s0 >= 0 && s0 + X_SPAN <= L;                 // every sample is real audio of this clip
        }
        unsigned todo = __ballot_sync(0xffffffffu, live);
        const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          {   // the tile two places further down the list: pull its audio into L2 now, so that its copy is an L2 hit
            const unsigned ahead = todo & (todo - 1);
            if (ahead && lane == __ffs(ahead) - 1 && tma)
              x_prefetch_l2(A.wave + (size_t)clip * (size_t)A.stride + (f0 * W_HOP - W_NFFT / 2), X_SPAN * 4);
          }
          const int c = __shfl_sync(0xffffffffu, clip, src), fr = __shfl_sync(0xffffffffu, f0, src);
          const int len = __shfl_sync(0xffffffffu, L, src), tm = __shfl_sync(0xffffffffu, (int)tma, src);
          const unsigned i6 = n % 6u, slot = i6 >= X_NSLOT ? i6 - X_NSLOT : i6;
          if (n >= 3) x_wait(C.bar(XB_AUD_EMPTY + (n + 3u) % 6u), ((n - 3u) / 6u) & 1u);   // tile n - 3 has been read
          if (lane == 0) {
            C.ring()[n & (X_RING - 1)] = make_int4(c, fr, len, tm);
            if (tm) {
              const float* gsrc = A.wave + (size_t)c * (size_t)A.stride + (fr * W_HOP - W_NFFT / 2);
              const unsigned dst = C.sbase + X_OFF_AUD + slot * X_AUD_BYTES;
              fence_proxy_async();         // earlier generic-proxy accesses to the slot vs. the async-proxy writes
              x_expect_tx(C.bar(XB_AUD_FULL + i6), X_AUD_BYTES);
              x_bulk_g2s(dst, gsrc, X_SUB * 4, C.bar(XB_AUD_FULL + i6));
              x_bulk_g2s(dst + X_SUB * 4, gsrc + 16 * W_HOP, X_SUB * 4, C.bar(XB_AUD_FULL + i6));
            } else {
              x_arrive(C.bar(XB_AUD_FULL + i6));
            }
          }
          __syncwarp();
          ++n;
        }
        if (valid_mask != 0xffffffffu) break;
      }
      for (int rep = 0; rep < X_GROUPS; ++rep) {           // one end marker per pass-2 group
        // same wait as for a tile: it is what guarantees that the ring entry (last used by tile n - 8) has been read
        if (n >= 3) x_wait(C.bar(XB_AUD_EMPTY + (n + 3u) % 6u), ((n - 3u) / 6u) & 1u);
        if (lane == 0) {
          C.ring()[n & (X_RING - 1)] = make_int4(-1, 0, 0, 0);
          x_arrive(C.bar(XB_AUD_FULL + n % 6u));
        }
        __syncwarp();
        ++n;
      }
    } else if (warp == X_WARP_AUX + 1) {
      // =============================== commit =============================================================
      // Everything stored since the last look goes out behind ONE gpu-scope fence (a fence per tile would pace the
      // whole pipeline: it waits for the SM's outstanding feature stores).  Lane g looks after group g.
      static_assert(X_GROUPS == 2, "the commit warp is written for two groups");
      int done = 0;                                         // tiles of my group committed so far
      bool live = lane < X_GROUPS;
      for (;;) {
        bool got = false;
        unsigned* w = nullptr;
        // (a counter, not an mbarrier phase: this warp may fall several tiles behind, and a parity wait cannot)
        if (live && x_lds_acquire(C.a_stored(lane)) >= X_GROUP_WARPS * (done + 1)) {
          const int ent = done & (X_CRING - 1);
          const int clip = C.c_clip(lane)[ent];
          if (clip < 0) {
            live = false;
          } else {
            got = true;
            const float* cm = C.c_max(lane) + ent * 4;
            const float m = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
            w = reinterpret_cast<unsigned*>(A.ws + clip);
            if (!(A.debug & 1)) atomicMax(w + 1, __float_as_uint(m));
            asm volatile("red.relaxed.cta.shared.max.s32 [%0], %1;" ::"r"(C.a_lastclip()), "r"(clip + 1) : "memory");
          }
        }
        if (__any_sync(0xffffffffu, got)) {
          // release: the features of these tiles (stored by the pass-2 warps, ordered before this warp by the
          // mbarrier) and the maxima above are visible to whoever sees the counts
          if (!(A.debug & 1)) {
            if (!(A.debug & 32)) __threadfence();
            if (got) atomicAdd(w, 1u);
          }
          if (got) { ++done; x_sts_release(C.a_committed(lane), done); }
        } else if (!__any_sync(0xffffffffu, live)) {
          break;
        } else {
          __nanosleep(200);
        }
      }
      if (lane == 0) x_sts_release(C.a_ended(), 1);
    } else {
      // =============================== floor ==============================================================
      // Warp wk takes clips wk, wk + 2, ...  A clip is cut into 371 fixed shares of <= 162 float4; this CTA owns shares
      // blockIdx, blockIdx + gridDim, ...  Each (clip, share) item is pulled into shared memory by ONE bulk copy (two
      // buffers per warp, the next item in flight while this one is clamped): the pass needs memory-level parallelism,
      // not instructions, and a warp's registers hold too few loads (measured: ~1 us per L2 round trip here).
      const int wk = warp - (X_WARP_AUX + 2);
      constexpr int VEC_PER_CLIP = W_NMEL * W_NFRAME / 4, VEC_PER_ROW = W_NFRAME / 4;
      const float y_silent = x_norm_log(X_EFLOOR);
      const int my_shares = ((int)blockIdx.x < X_FSHARES) ? (X_FSHARES - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int nclips = (A.debug & 1) ? 0 : (A.batch - wk + 1) / 2;
      const int nitems = nclips * my_shares;
      const unsigned fbar = C.sbase + X_OFF_FBAR + 16u * (unsigned)wk;      // two mbarriers per warp
      const unsigned fbuf = C.sbase + X_OFF_FBUF + 2u * X_FBUF_BYTES * (unsigned)wk;
      const float4* fbuf_p = reinterpret_cast<const float4*>(C.smem + X_OFF_FBUF + 2 * X_FBUF_BYTES * wk);
      if (lane == 0) {
        mbar_init(reinterpret_cast<unsigned long long*>(C.smem + X_OFF_FBAR) + 2 * wk, 1);
        mbar_init(reinterpret_cast<unsigned long long*>(C.smem + X_OFF_FBAR) + 2 * wk + 1, 1);
        fence_proxy_async();
      }
      __syncwarp();
      long long t_gate = 0, t_poll = 0, n_spin = 0, t_wait = 0, t_clamp = 0;
      const bool tm = (A.debug & 16) != 0;
      const long long t_start = clock64();
      int ready_clip = -1; float rthr = 0.0f; int rfs = 0;          // last clip found complete, its floor and silent frame
      // an item in flight: where it goes back to, how many float4, first column, floor value, first silent frame
      struct Item { float4* base; int cnt, col; float thr; int fs; };
      auto open_item = [&](int cl, int k, int b) -> Item {         // clip wk + 2 cl, my k-th share: waits for the clip, starts the copy into buffer b
        const int c = wk + 2 * cl, sh = (int)blockIdx.x + k * (int)gridDim.x;
        if (c != ready_clip) {
          // wait first on this CTA's own progress (shared memory; the CTAs advance through the clips together), then on
          // the clip's word -- 296 warps polling one L2 line from the start would starve the commit warps' atomics on it
          const long long t0 = tm ? clock64() : 0;
          while (x_lds_acquire(C.a_lastclip()) <= c + 1 && !x_lds_acquire(C.a_ended())) __nanosleep(400);
          const long long t1 = tm ? clock64() : 0;
          t_gate += t1 - t0;
          const unsigned target = (unsigned)x_live_tiles(x_clip_len(A.lengths, A.stride, c));
          unsigned long long w = x_ld_acquire(A.ws + c);
          for (unsigned spins = 0; (unsigned)(w & 0xffffu) < target; ++spins) {
            __nanosleep(300);
            w = x_ld_acquire(A.ws + c);
            ++n_spin;
            if (spins > (1u << 22)) {                                // seconds: the workspace was not zeroed; give up
              if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(A.ws + A.batch) + 1, ((unsigned)(c + 1) << 8) | (unsigned)(w & 0xffu));
              break;
            }
          }
          if (tm) t_poll += clock64() - t1;
          rthr = fmaxf(x_norm_log(__uint_as_float((unsigned)(w >> 32))) - 2.0f, y_silent);   // -inf for 0: the floor wins
          rfs = (int)target * X_TILE;                                // frames from here on belong to silent tiles
          ready_clip = c;
          asm volatile("fence.proxy.async.global;" ::: "memory");   // the features come in through the async proxy
        }
        const int b0 = sh * VEC_PER_CLIP / X_FSHARES, b1 = (sh + 1) * VEC_PER_CLIP / X_FSHARES;     // < 2^31
        Item r;
        r.base = reinterpret_cast<float4*>(A.out + (size_t)c * (W_NMEL * W_NFRAME)) + b0;
        r.cnt = b1 - b0; r.col = b0 % VEC_PER_ROW; r.thr = rthr; r.fs = rfs;
        if (lane == 0) {
          fence_proxy_async();                                       // the buffer was last read through the generic proxy
          x_expect_tx(fbar + 8u * (unsigned)b, (unsigned)r.cnt * 16u);
          x_bulk_g2s(fbuf + (unsigned)b * X_FBUF_BYTES, r.base, (unsigned)r.cnt * 16u, fbar + 8u * (unsigned)b);
        }
        return r;
      };
      constexpr int X_FITER = (X_FBUF_BYTES / 16 + 31) / 32;         // 6 float4 per lane at most
      Item cur = {nullptr, 0, 0, 0.0f, 0}, nxt = cur;
      int ncl = 0, nk = 0;                                           // the item after `cur`: clip-list index, share index
      if (nitems > 0) { cur = open_item(0, 0, 0); nk = 1; if (nk == my_shares) { nk = 0; ncl = 1; } }
      for (int it = 0; it < nitems; ++it) {
        const int b = it & 1;
        if (it + 1 < nitems) {                                       // the next item's copy runs under this item's clamp
          nxt = open_item(ncl, nk, b ^ 1);
          if (++nk == my_shares) { nk = 0; ++ncl; }
        }
        const long long tw0 = tm ? clock64() : 0;
        x_wait(fbar + 8u * (unsigned)b, (unsigned)(it >> 1) & 1u);
        const long long tw1 = tm ? clock64() : 0;
        t_wait += tw1 - tw0;
        const float4* src = fbuf_p + b * (X_FBUF_BYTES / 16) + lane;
        float4 q[X_FITER];
        float4* dst = cur.base + lane;
        const int n_mine = cur.cnt - lane;
#pragma unroll
        for (int e = 0; e < X_FITER; ++e) q[e] = 32 * e < n_mine ? src[32 * e] : make_float4(0.f, 0.f, 0.f, 0.f);
        int col = cur.col + lane; col -= col >= VEC_PER_ROW ? VEC_PER_ROW : 0;
#pragma unroll
        for (int e = 0; e < X_FITER; ++e) {
          if (32 * e < n_mine) {
            if (col * 4 >= cur.fs) {                                 // a frame of a silent tile: nothing was stored there
              dst[32 * e] = make_float4(cur.thr, cur.thr, cur.thr, cur.thr);
            } else if (fminf(fminf(q[e].x, q[e].y), fminf(q[e].z, q[e].w)) < cur.thr) {
              dst[32 * e] = make_float4(fmaxf(q[e].x, cur.thr), fmaxf(q[e].y, cur.thr), fmaxf(q[e].z, cur.thr), fmaxf(q[e].w, cur.thr));
            }
          }
          col += 32; col -= col >= VEC_PER_ROW ? VEC_PER_ROW : 0;
        }
        __syncwarp();                                                // everybody has read the buffer before it is refilled
        if (tm) t_clamp += clock64() - tw1;
        cur = nxt;
      }
      if (tm && lane == 0) {
        long long* dbg = reinterpret_cast<long long*>(A.ws + A.batch + 2) + 8 * (2 * blockIdx.x + wk);
        dbg[0] = clock64() - t_start; dbg[1] = t_gate; dbg[2] = t_poll; dbg[3] = n_spin; dbg[4] = nitems; dbg[5] = t_wait; dbg[6] = t_clamp;
      }
      // the last CTA to get here leaves the workspace zero for the next launch
      x_bar_dyn(6, 64);
      if (wk == 0 && !(A.debug & 1)) {
        unsigned old = 0;
        if (lane == 0) old = atomicInc(reinterpret_cast<unsigned*>(A.ws + A.batch), gridDim.x - 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == gridDim.x - 1)
          for (int c = lane; c < A.batch; c += 32) A.ws[c] = 0ull;
      }
    }
  }
}
