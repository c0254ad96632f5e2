// urban_packed.cuh -- urban preset: the fused mel kernel (persistent CTAs, two frames per lane).
#pragma once
// ------------------------------------------------------------------------------------------------
// Replaces TA:transforms/_transforms.py:621-631 (MelSpectrogram.forward) and the torch.log(mel + 1e-9) of
// REF:urban_sounds/dataset.py:56:
//
//  * the frames of the whole batch are one flat list g = clip * n_frames + t; a tile is 32 consecutive
//    entries (no padding slots except in the very last tile), and a persistent CTA per SM walks tiles
//    blockIdx.x, blockIdx.x + gridDim.x, ...
//  * two frames per lane as float2 (FFMA2 / FADD2 / FMUL2), as in the Whisper kernel
//  * pass 1 has lane = r (the residue of the sample index mod 32) and task = frame pair: the 32 samples
//    x[r + 32 j] of a lane are strided by 128 B while the 32 lanes of a warp read consecutive addresses, so the
//    audio goes global -> registers in fully coalesced loads with no shared-memory staging at all.  The two
//    frames of a pair overlap by half (hop 512 = 16 * 32): 48 loads feed both.  Reflect padding and pairs that
//    straddle two clips take a generic per-sample path (2-3 tasks per clip).
//    Output: Y_r[k2], k2 = 0..16 (real 32-point DFT over j)  ->  E[r][row][frame], row pitch chosen so that
//    both the lane = r stores here and the lane = frame loads of pass 2 are bank-conflict free.
//  * pass 2 has task = k2 and lane = (h, q): q = frame pair, h = parity of k1.  One decimation-in-frequency
//    step splits the complex 32-point DFT over r into two independent 16-point DFTs (even / odd k1):
//        X[k2 + 32 (2m + h)] = sum_{r<16} W16^(r m) * ( Y_r T_h[k2][r] + Y_{r+16} T_h[k2][r+16] )
//        T_h[k2][r]      = W1024^(r k2) W32^(r h),   T_h[k2][r+16] = (-1)^h W1024^((r+16) k2) W32^(r h)
//    so both half-warps run the same instructions on different table rows and never exchange data.
//    k2 = 0 and k2 = 16 have real inputs and need only k1 <= 16 / k1 < 16: one warp runs both, 15 warps run
//    k2 = 1..15.  |X|^2 goes to P[bin][frame].
//  * mel: 32 half-warp groups share the 64 HTK filters (998 taps, cost-balanced on the host), FFMA2 over the
//    frame pair, optional log(. + eps), stores straight to out[clip][mel][t].
//  * schedule per tile: { pass 2; issue the next tile's audio loads } sync { mel of this tile and pass-1 DFT
//    of the next tile, in either order: two warps per scheduler run the (LSU-bound) mel first and two the
//    (FMA-bound) DFT first } sync.  The loads are in flight across the barrier and the first half-phase.
// All tables come from one image built by b200mel_create (u2_build_image) and copied to shared memory once
// per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int U2_THREADS = 512, U2_WARPS = 16;
constexpr int U2_EP = 32 * 32 + 2;                          // E floats per r: 32 rows x 32 frames + 2 (bank skew)
constexpr int U2_E = 32 * U2_EP;
constexpr int U2_P = U_NBIN * 32;
constexpr int U2_NNZ = B200MEL_U_NNZ;                       // 998
constexpr int U2_IMG_T4 = 0;                                // float4 (c, c, s, s) [17][32][2]
constexpr int U2_IMG_WIN2 = U2_IMG_T4 + 17 * 32 * 2 * 4;    // float2 (w, w) [1024]
constexpr int U2_MW_MAX = 1536;                             // filter weights, every filter padded with zeros to 8 k taps
constexpr int U2_IMG_MW = U2_IMG_WIN2 + 2048;               // float [U2_MW_MAX]
constexpr int U2_IMG_GOFF = U2_IMG_MW + U2_MW_MAX;          // int [33] (+ padding to 64)
constexpr int U2_IMG_GENT = U2_IMG_GOFF + 64;               // int4 (first bin, padded taps, weight offset, mel) [64]
constexpr int U2_IMG = U2_IMG_GENT + 64 * 4;
constexpr int U2_SMEM_BYTES = (U2_E + U2_P + U2_IMG) * 4;
static_assert(U2_SMEM_BYTES <= 227 * 1024, "urban packed kernel: shared memory");
static_assert(U2_E % 4 == 0 && (U2_E + U2_P) % 4 == 0 && U2_IMG % 4 == 0 && U2_IMG_GENT % 4 == 0, "16-byte carve-up");

// E row of Re/Im Y[k2]: k2 = 0 and k2 = 16 are real and share rows 0 and 1
__device__ __forceinline__ constexpr int u2_row_re(int k2) { return k2 == 0 ? 0 : (k2 == 16 ? 1 : 2 * k2); }

struct U2Geom {
  const float* __restrict__ wave;
  long long stride;
  unsigned total_frames, n_frames;     // flat frame indices fit 32 bits (checked by the host)
  int n_samples;
};

// L2 prefetch of the 6 KB a frame pair reads (fast-path pairs only; the edge pairs are few)
__device__ __forceinline__ void u2_prefetch_pair(const U2Geom& G, unsigned g) {
  if (g + 1 >= G.total_frames) return;
  const unsigned clip = g / G.n_frames;
  const unsigned t = g - clip * G.n_frames;
  const int base = (int)t * U_HOP - U_NFFT / 2;
  if (t + 1 < G.n_frames && base >= 0 && base + U_HOP + U_NFFT <= G.n_samples)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G.wave + (size_t)clip * G.stride + base), "r"((U_HOP + U_NFFT) * 4) : "memory");
}

// audio of one frame pair (flat frames g, g + 1) -> x[j] = (frame g, frame g + 1) sample r + 32 j
__device__ __forceinline__ void u2_load_pair(const U2Geom& G, unsigned g, int r, float2 x[32]) {
  const unsigned clip = g / G.n_frames;
  const unsigned t = g - clip * G.n_frames;
  const int base = (int)t * U_HOP - U_NFFT / 2;
  const bool fast = (g + 1 < G.total_frames) && (t + 1 < G.n_frames) && (base >= 0) && (base + U_HOP + U_NFFT <= G.n_samples);
  if (fast) {
    const float* __restrict__ p = G.wave + (size_t)clip * G.stride + base + r;
    float v[48];
#pragma unroll
    for (int j = 0; j < 48; ++j) v[j] = __ldg(p + 32 * j);
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = make_float2(v[j], v[j + 16]);
    return;
  }
  // clip edges (reflect padding, no edge repeat), pairs that straddle two clips, the tail of the last tile:
  // 2-3 pairs per clip
  float v[2][32];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const unsigned gs = g + s;
    const unsigned c = gs / G.n_frames;
    const int b = (int)(gs - c * G.n_frames) * U_HOP - U_NFFT / 2 + r;
    const bool valid = gs < G.total_frames;
    const float* __restrict__ src = G.wave + (size_t)c * G.stride;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      int i = b + 32 * j;
      i = i < 0 ? -i : (i >= G.n_samples ? 2 * (G.n_samples - 1) - i : i);
      v[s][j] = (valid && i >= 0 && i < G.n_samples) ? __ldg(src + i) : 0.0f;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = make_float2(v[0][j], v[1][j]);
}

// window, real 32-point DFT over j, store Y_r[k2] for the pair p (frames 2p, 2p + 1 of the tile)
__device__ __forceinline__ void u2_pass1_dft(const float2 x[32], int r, int p, const float* __restrict__ s_img, float* __restrict__ s_e) {
  const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_img + U2_IMG_WIN2) + r;
  float2 xw[32], Xr[17], Xi[17];
#pragma unroll
  for (int j = 0; j < 32; ++j) xw[j] = b2::vmul(x[j], w2[32 * j]);
  b2::real_dft32_windowed(xw, Xr, Xi);
  float2* __restrict__ dst = reinterpret_cast<float2*>(s_e + r * U2_EP) + p;     // + row * 16
  dst[0] = Xr[0];
  dst[16] = Xr[16];
#pragma unroll
  for (int k = 1; k < 16; ++k) { dst[(2 * k) * 16] = Xr[k]; dst[(2 * k + 1) * 16] = Xi[k]; }
}

// bin of output m (k1 = 2m + h) of task k2; m >= 8 is the conjugate mirror 1024 - k
template <int M>
__device__ __forceinline__ int u2_bin(int k2, int h) { return M < 8 ? k2 + 32 * h + 64 * M : 1024 - 64 * M - 32 * h - k2; }

template <int M>
__device__ __forceinline__ void u2_store_power(const float2 Xr[16], const float2 Xi[16], float2* __restrict__ p2, int k2, int h) {
  p2[u2_bin<M>(k2, h) * 16] = b2::vfma(Xr[M], Xr[M], b2::vmul(Xi[M], Xi[M]));
}

// pass 2, complex inputs (k2 = 1..15): all 32 outputs are wanted
__device__ __forceinline__ void u2_pass2(int k2, int h, int q, const float* __restrict__ s_img, const float* __restrict__ s_e, float* __restrict__ s_p) {
  const float2* __restrict__ e2 = reinterpret_cast<const float2*>(s_e) + (2 * k2) * 16 + q;     // + r * (U2_EP / 2); Im = +16
  const float4* __restrict__ t4 = reinterpret_cast<const float4*>(s_img + U2_IMG_T4) + (k2 * 32) * 2 + h;   // + r * 2
  float2 ur[16], ui[16], Xr[16], Xi[16];
  // T_h[k2][r + 16] = T_h[k2][r] * C with C = T_h[k2][16] (T_h[k2][0] = 1), so u = T[r] (Y_r + C Y_{r+16}): the same
  // eight packed operations per r with one table load instead of two
  const float4 tc = t4[16 * 2];
  const float2 cc = make_float2(tc.x, tc.y), cs = make_float2(tc.z, tc.w);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 ar = e2[r * (U2_EP / 2)], ai = e2[r * (U2_EP / 2) + 16];
    const float2 br = e2[(r + 16) * (U2_EP / 2)], bi = e2[(r + 16) * (U2_EP / 2) + 16];
    const float4 ta = t4[r * 2];
    const float2 ac = make_float2(ta.x, ta.y), as = make_float2(ta.z, ta.w);
    // (yr + i yi)(c - i s) = (yr c + yi s) + i (yi c - yr s)
    const float2 tr = b2::vfma(br, cc, b2::vfma(bi, cs, ar));
    const float2 ti = b2::vfma(bi, cc, b2::vfma(b2::vneg(br), cs, ai));
    ur[r] = b2::vfma(tr, ac, b2::vmul(ti, as));
    ui[r] = b2::vfma(ti, ac, b2::vmul(b2::vneg(tr), as));
  }
  b2::cplx_dft16(ur, ui, Xr, Xi);
  float2* __restrict__ p2 = reinterpret_cast<float2*>(s_p) + q;
  u2_store_power<0>(Xr, Xi, p2, k2, h);   u2_store_power<1>(Xr, Xi, p2, k2, h);
  u2_store_power<2>(Xr, Xi, p2, k2, h);   u2_store_power<3>(Xr, Xi, p2, k2, h);
  u2_store_power<4>(Xr, Xi, p2, k2, h);   u2_store_power<5>(Xr, Xi, p2, k2, h);
  u2_store_power<6>(Xr, Xi, p2, k2, h);   u2_store_power<7>(Xr, Xi, p2, k2, h);
  u2_store_power<8>(Xr, Xi, p2, k2, h);   u2_store_power<9>(Xr, Xi, p2, k2, h);
  u2_store_power<10>(Xr, Xi, p2, k2, h);  u2_store_power<11>(Xr, Xi, p2, k2, h);
  u2_store_power<12>(Xr, Xi, p2, k2, h);  u2_store_power<13>(Xr, Xi, p2, k2, h);
  u2_store_power<14>(Xr, Xi, p2, k2, h);  u2_store_power<15>(Xr, Xi, p2, k2, h);
}

// pass 2, real inputs (K2 = 0 or 16).  K2 = 0 gives bins 32 k1, k1 = 0..16; K2 = 16 gives bins 16 + 32 k1, k1 = 0..15;
// the remaining k1 are mirrors of those and are neither computed nor stored.
template <int K2>
__device__ __forceinline__ void u2_pass2_real(int h, int q, const float* __restrict__ s_img, const float* __restrict__ s_e, float* __restrict__ s_p) {
  const float2* __restrict__ e2 = reinterpret_cast<const float2*>(s_e) + u2_row_re(K2) * 16 + q;
  const float4* __restrict__ t4 = reinterpret_cast<const float4*>(s_img + U2_IMG_T4) + (K2 * 32) * 2 + h;
  float2 ur[16], ui[16], Xr[16], Xi[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 ar = e2[r * (U2_EP / 2)], br = e2[(r + 16) * (U2_EP / 2)];
    const float4 ta = t4[r * 2], tb = t4[(r + 16) * 2];
    ur[r] = b2::vfma(ar, make_float2(ta.x, ta.y), b2::vmul(br, make_float2(tb.x, tb.y)));
    ui[r] = b2::vfma(b2::vneg(ar), make_float2(ta.z, ta.w), b2::vmul(b2::vneg(br), make_float2(tb.z, tb.w)));
  }
  b2::cplx_dft16(ur, ui, Xr, Xi);
  float2* __restrict__ p2 = reinterpret_cast<float2*>(s_p) + q;
  u2_store_power<0>(Xr, Xi, p2, K2, h);  u2_store_power<1>(Xr, Xi, p2, K2, h);
  u2_store_power<2>(Xr, Xi, p2, K2, h);  u2_store_power<3>(Xr, Xi, p2, K2, h);
  u2_store_power<4>(Xr, Xi, p2, K2, h);  u2_store_power<5>(Xr, Xi, p2, K2, h);
  u2_store_power<6>(Xr, Xi, p2, K2, h);  u2_store_power<7>(Xr, Xi, p2, K2, h);
  if (K2 == 0 && h == 0) u2_store_power<8>(Xr, Xi, p2, K2, h);                    // k1 = 16 -> bin 512
}

// mel filters of this half-warp's group over the frame pair q of tile `tile`
__device__ __forceinline__ void u2_mel(const U2Geom& G, int tile, int group, int q, float log_eps,
                                       const float* __restrict__ s_img, const float* __restrict__ s_p, float* __restrict__ out) {
  const int* __restrict__ goff = reinterpret_cast<const int*>(s_img + U2_IMG_GOFF);
  const int4* __restrict__ gent = reinterpret_cast<const int4*>(s_img + U2_IMG_GENT);
  const float* __restrict__ mw = s_img + U2_IMG_MW;
  const float2* __restrict__ p2 = reinterpret_cast<const float2*>(s_p) + q;
  const unsigned g0 = (unsigned)tile * 32u + 2u * q;
  unsigned c0 = g0 / G.n_frames, t0 = g0 - c0 * G.n_frames;
  unsigned c1 = c0, t1 = t0 + 1;
  if (t1 == G.n_frames) { t1 = 0; ++c1; }
  const bool v0 = g0 < G.total_frames, v1 = g0 + 1 < G.total_frames;
  float* __restrict__ o0 = out + (size_t)c0 * ((size_t)U_NMEL * G.n_frames) + t0;
  float* __restrict__ o1 = out + (size_t)c1 * ((size_t)U_NMEL * G.n_frames) + t1;
  const bool take_log = log_eps >= 0.0f;
  const int e1 = goff[group + 1];
#pragma unroll 1
  for (int e = goff[group]; e < e1; ++e) {
    const int4 f = gent[e];
    const float2* __restrict__ pp = p2 + f.x * 16;
    const float4* __restrict__ ww = reinterpret_cast<const float4*>(mw + f.z);
    // 8 taps per trip over a run that u2_build_image keeps inside P (zero weights pad it to 8 k rows of the same
    // frame pair), two accumulation chains
    float2 a0 = make_float2(0.0f, 0.0f), a1 = a0;
#pragma unroll 1
    for (int j = 0; j < f.y; j += 8) {
      const float4 wa = ww[0], wb = ww[1];
      float2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = pp[i * 16];
      a0.x = __fmaf_rn(v[0].x, wa.x, a0.x); a0.y = __fmaf_rn(v[0].y, wa.x, a0.y);
      a1.x = __fmaf_rn(v[1].x, wa.y, a1.x); a1.y = __fmaf_rn(v[1].y, wa.y, a1.y);
      a0.x = __fmaf_rn(v[2].x, wa.z, a0.x); a0.y = __fmaf_rn(v[2].y, wa.z, a0.y);
      a1.x = __fmaf_rn(v[3].x, wa.w, a1.x); a1.y = __fmaf_rn(v[3].y, wa.w, a1.y);
      a0.x = __fmaf_rn(v[4].x, wb.x, a0.x); a0.y = __fmaf_rn(v[4].y, wb.x, a0.y);
      a1.x = __fmaf_rn(v[5].x, wb.y, a1.x); a1.y = __fmaf_rn(v[5].y, wb.y, a1.y);
      a0.x = __fmaf_rn(v[6].x, wb.z, a0.x); a0.y = __fmaf_rn(v[6].y, wb.z, a0.y);
      a1.x = __fmaf_rn(v[7].x, wb.w, a1.x); a1.y = __fmaf_rn(v[7].y, wb.w, a1.y);
      pp += 8 * 16; ww += 2;
    }
    float2 acc = make_float2(a0.x + a1.x, a0.y + a1.y);
    if (take_log) { acc.x = __logf(acc.x + log_eps); acc.y = __logf(acc.y + log_eps); }
    const unsigned mo = (unsigned)f.w * G.n_frames;
    if (v0) o0[mo] = acc.x;
    if (v1) o1[mo] = acc.y;
  }
}

__global__ void __launch_bounds__(U2_THREADS, 1)
urban_mel_packed_kernel(const float* __restrict__ wave, long long stride, int n_samples, int n_frames, unsigned total_frames,
                        int n_tiles, float log_eps, const float* __restrict__ image, float* __restrict__ out) {
  extern __shared__ __align__(1024) float smem[];
  float* s_e = smem;
  float* s_p = smem + U2_E;
  float* s_img = smem + U2_E + U2_P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = lane >> 4, q = lane & 15;
  const bool dft_first = (warp >> 2) & 1;             // two warps of either kind per scheduler
  for (int i = tid; i < U2_IMG / 4; i += U2_THREADS)
    reinterpret_cast<float4*>(s_img)[i] = __ldg(reinterpret_cast<const float4*>(image) + i);
  U2Geom G{wave, stride, total_frames, (unsigned)n_frames, n_samples};
  int cur = -1, nxt = blockIdx.x;                     // the first trip has no current tile: it only runs pass 1
  float2 x[32];
#pragma unroll 1
  for (;;) {
    if (cur >= 0) {
      if (warp == 0) {
        u2_pass2_real<0>(h, q, s_img, s_e, s_p);
        u2_pass2_real<16>(h, q, s_img, s_e, s_p);
      } else {
        u2_pass2(warp, h, q, s_img, s_e, s_p);
      }
    }
    const bool more = nxt < n_tiles;
    if (more) u2_load_pair(G, (unsigned)nxt * 32u + 2u * warp, lane, x);     // in flight across the barrier
    if (lane == 0 && nxt + (int)gridDim.x < n_tiles) u2_prefetch_pair(G, (unsigned)(nxt + gridDim.x) * 32u + 2u * warp);
    __syncthreads();                                  // P complete, E free (first trip: table image in place)
    if (dft_first && more) u2_pass1_dft(x, lane, warp, s_img, s_e);
    if (cur >= 0) u2_mel(G, cur, 2 * warp + h, q, log_eps, s_img, s_p, out);
    if (!dft_first && more) u2_pass1_dft(x, lane, warp, s_img, s_e);
    if (!more) break;
    __syncthreads();                                  // E complete, P free
    cur = nxt;
    nxt += gridDim.x;
  }
}

// Host: the table image (see the U2_IMG_* layout).  Twiddles are evaluated in double and rounded once.
static void u2_build_image(std::vector<float>& img) {
  img.assign(U2_IMG, 0.0f);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k2 = 0; k2 < 17; ++k2)
    for (int r = 0; r < 32; ++r)
      for (int h = 0; h < 2; ++h) {
        const int e = (r * k2 + 32 * (r & 15) * h + (r >= 16 ? 512 * h : 0)) & 1023;
        const float c = (float)cos(two_pi * e / 1024.0), s = (float)sin(two_pi * e / 1024.0);
        float* d = img.data() + U2_IMG_T4 + ((k2 * 32 + r) * 2 + h) * 4;
        d[0] = c; d[1] = c; d[2] = s; d[3] = s;
      }
  for (int n = 0; n < 1024; ++n) img[U2_IMG_WIN2 + 2 * n] = img[U2_IMG_WIN2 + 2 * n + 1] = host_tab::c_win1024[n];
  // 64 filters -> 32 half-warp groups, longest-processing-time first (cost = padded taps + 6 per filter)
  int order[64], load[32] = {0}, owner[64], pad[64];
  for (int m = 0; m < 64; ++m) { order[m] = m; pad[m] = (host_tab::kUMelLen[m] + 7) / 8 * 8; }
  std::sort(order, order + 64, [&](int a, int b) { return pad[a] != pad[b] ? pad[a] > pad[b] : a < b; });
  for (int i = 0; i < 64; ++i) {
    int best = 0;
    for (int g = 1; g < 32; ++g) if (load[g] < load[best]) best = g;
    owner[order[i]] = best;
    load[best] += pad[order[i]] + 6;
  }
  // the two groups of a warp run in lock step: put groups of similar load side by side
  int gorder[32];
  for (int g = 0; g < 32; ++g) gorder[g] = g;
  std::sort(gorder, gorder + 32, [&](int a, int b) { return load[a] != load[b] ? load[a] > load[b] : a < b; });
  int* goff = reinterpret_cast<int*>(img.data() + U2_IMG_GOFF);
  int* gent = reinterpret_cast<int*>(img.data() + U2_IMG_GENT);
  int n = 0, woff = 0;
  for (int gi = 0; gi < 32; ++gi) {
    goff[gi] = n;
    for (int m = 0; m < 64; ++m)
      if (owner[m] == gorder[gi]) {
        // the padded run of 8 k taps must stay inside P's 513 rows: a filter that ends near the last bin starts its run
        // earlier and leads with zero weights instead of trailing with them
        int first = host_tab::kUMelStart[m];
        const int lead = first + pad[m] > U_NBIN ? first + pad[m] - U_NBIN : 0;
        first -= lead;
        if (first < 0) { img.clear(); return; }
        for (int j = 0; j < host_tab::kUMelLen[m]; ++j) img[U2_IMG_MW + woff + lead + j] = host_tab::c_umelw[host_tab::kUMelOff[m] + j];
        gent[4 * n + 0] = first; gent[4 * n + 1] = pad[m];
        gent[4 * n + 2] = woff;                    gent[4 * n + 3] = m;
        woff += pad[m];
        ++n;
      }
  }
  if (woff > U2_MW_MAX) { img.clear(); return; }
  goff[32] = n;
}
