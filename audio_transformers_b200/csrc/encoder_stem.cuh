// encoder_stem.cuh -- the Whisper encoder stem on the 5th-generation tensor cores.
#pragma once
// ================================================================================================
// Replaces HF:models/whisper/modeling_whisper.py:619-625 for the tiny model (80 mels, d_model 384):
//     x = gelu(conv1(input_features))          Conv1d(80 -> 384, k 3, pad 1)               (B, 384, 3000)
//     x = gelu(conv2(x))                       Conv1d(384 -> 384, k 3, stride 2, pad 1)    (B, 384, 1500)
//     hidden_states = x.permute(0, 2, 1) + embed_positions.weight                          (B, 1500, 384)
// consuming the (B, 80, 3000) feature map of the front end as it lies in HBM / L2.
//
// Both convolutions are GEMMs C[m][n] = sum_k A[m][k] W[n][k] with m = time, n = output channel and
// k = (tap, input channel), run by ONE kernel template (es_gemm_kernel) on tcgen05.mma with BF16 operands and FP32
// accumulation in tensor memory:
//   conv1   A1[t][tap * 80 + ci] = x[ci][t + tap - 1]: an im2col image in BF16 (K padded 240 -> 256 with zeros), written
//           by es_im2col_kernel (the features are frames-fastest, the MMA wants K-fastest rows: the transpose goes
//           through shared memory once);
//   conv2   A2[t'][tap * 384 + ci] = h[2 t' + tap - 1][ci] needs no im2col: h is stored time-major with one zero row in
//           front of and behind every clip, and a 4-D tensor map (channel, row parity, row pair, clip) lets TMA fetch the
//           stride-2 rows of a tap as one box.
// A CTA owns 128 (time) x 192 (channel) output tiles.  Warp 0: TMA producer (one lane) through a 4-stage ring of
// {A 128 x 64, W 192 x 64} BF16 tiles in the 128-byte-swizzled K-major layout; warp 1: tensor-memory allocation and the
// single thread that issues tcgen05.mma (M 128, N 192, K 16) and commits to mbarriers; warps 2..13: epilogue
// (tcgen05.ld, bias, exact GELU, conv1: BF16 rows of h; conv2: + positional embedding, FP32 rows of the output).  Two
// accumulators of 192 columns alternate, so the epilogue of a tile runs under the MMAs of the next one.
// ================================================================================================


constexpr int ES_NMEL = 80, ES_T = 3000, ES_D = 384, ES_T2 = 1500;
constexpr int ES_K1 = 256;                       // 3 taps x 80 mels, zero padded to a multiple of 64
constexpr int ES_K2 = 3 * ES_D;                  // 1152
constexpr int ES_HROWS = ES_T + 2;               // rows of h per clip: a zero row, 3000 frames, a zero row
constexpr int ES_BM = 128, ES_BN = 192, ES_BK = 64, ES_UK = 16;
constexpr int ES_STAGES = 4;
constexpr int ES_A_BYTES = ES_BM * ES_BK * 2, ES_B_BYTES = ES_BN * ES_BK * 2, ES_STAGE_BYTES = ES_A_BYTES + ES_B_BYTES;
constexpr int ES_EPI_WARPS = 12, ES_THREADS = (2 + ES_EPI_WARPS) * 32;   // three epilogue warps per TMEM lane quarter, two 32-column chunks each
constexpr int ES_ACC_COLS = 256, ES_TMEM_COLS = 512;     // two accumulators, 192 columns used of each 256
constexpr int ES_SMEM_BYTES = ES_STAGES * ES_STAGE_BYTES + ES_EPI_WARPS * 4096 /* epilogue staging */ + 1024 /* alignment slack */ + 256 /* barriers, tmem address */;
static_assert(ES_A_BYTES % 1024 == 0 && ES_B_BYTES % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(ES_D % ES_BN == 0 && ES_K1 % ES_BK == 0 && ES_K2 % ES_BK == 0 && ES_D % ES_BK == 0, "tiling");

// ---- im2col of the feature map for conv1 (and the zero rows of h) ------------------------------------------------------
constexpr int ES_IC_FRAMES = 64, ES_IC_THREADS = 256, ES_IC_PITCH = ES_IC_FRAMES + 3;
__global__ void __launch_bounds__(ES_IC_THREADS)
es_im2col_kernel(const float* __restrict__ feat, __nv_bfloat16* __restrict__ a1, __nv_bfloat16* __restrict__ h) {
  __shared__ float s[ES_NMEL][ES_IC_PITCH];                       // frames t0 - 1 .. t0 + 64
  const int clip = blockIdx.y, t0 = blockIdx.x * ES_IC_FRAMES, tid = threadIdx.x;
  const float* __restrict__ src = feat + (size_t)clip * (ES_NMEL * ES_T);
  for (int i = tid; i < ES_NMEL * (ES_IC_FRAMES + 2); i += ES_IC_THREADS) {
    const int ci = i / (ES_IC_FRAMES + 2), c = i - ci * (ES_IC_FRAMES + 2), t = t0 - 1 + c;
    s[ci][c] = (t >= 0 && t < ES_T) ? __ldg(src + ci * ES_T + t) : 0.0f;
  }
  if (blockIdx.x == 0) {                                          // rows -1 and 3000 of h are the convolution's zero padding
    uint4* z0 = reinterpret_cast<uint4*>(h + (size_t)clip * ES_HROWS * ES_D);
    uint4* z1 = reinterpret_cast<uint4*>(h + ((size_t)clip * ES_HROWS + ES_HROWS - 1) * ES_D);
    for (int i = tid; i < ES_D * 2 / 16; i += ES_IC_THREADS) { z0[i] = make_uint4(0, 0, 0, 0); z1[i] = make_uint4(0, 0, 0, 0); }
  }
  __syncthreads();
  // a row of A1 is 32 chunks of 8 BF16 (16 bytes): chunk j < 30 holds tap j / 10, mels 8 (j % 10) .. + 7; chunks 30, 31 zeros
  for (int u = tid; u < ES_IC_FRAMES * 32; u += ES_IC_THREADS) {
    const int row = u >> 5, j = u & 31, t = t0 + row;
    if (t >= ES_T) continue;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < 30) {
      const int tap = j / 10, ci0 = (j - tap * 10) * 8;
      const float* p = &s[ci0][row + tap];
      __nv_bfloat162 b0 = __floats2bfloat162_rn(p[0], p[ES_IC_PITCH]);
      __nv_bfloat162 b1 = __floats2bfloat162_rn(p[2 * ES_IC_PITCH], p[3 * ES_IC_PITCH]);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(p[4 * ES_IC_PITCH], p[5 * ES_IC_PITCH]);
      __nv_bfloat162 b3 = __floats2bfloat162_rn(p[6 * ES_IC_PITCH], p[7 * ES_IC_PITCH]);
      v.x = *reinterpret_cast<unsigned*>(&b0); v.y = *reinterpret_cast<unsigned*>(&b1);
      v.z = *reinterpret_cast<unsigned*>(&b2); v.w = *reinterpret_cast<unsigned*>(&b3);
    }
    reinterpret_cast<uint4*>(a1 + ((size_t)clip * ES_T + t) * ES_K1)[j] = v;
  }
}

// ---- tcgen05 / TMA primitives -------------------------------------------------------------------------------------------
__device__ __forceinline__ void es_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void es_tma_load_4d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void es_tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void es_prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void es_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void es_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor: K-major rows of 128 bytes, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long es_smem_desc(unsigned saddr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);        // start address, bits [0, 14)
  d |= (unsigned long long)1 << 16;                            // leading byte offset (unused with swizzled K-major: 1)
  d |= (unsigned long long)(1024 >> 4) << 32;                  // stride byte offset, bits [32, 46)
  d |= (unsigned long long)1 << 46;                            // descriptor version of sm_100
  d |= (unsigned long long)2 << 61;                            // SWIZZLE_128B
  return d;
}
// instruction descriptor: D FP32, A / B BF16, both K-major, N = 192, M = 128
constexpr unsigned ES_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(ES_BN >> 3) << 17) | ((unsigned)(ES_BM >> 4) << 24);
__device__ __forceinline__ void es_umma(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(ES_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void es_umma_commit(unsigned long long* bar) {   // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void es_tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// torch.nn.functional.gelu(approximate="none") = x Phi(x) for two values at once (packed FP32 pairs: FFMA2 / FMUL2).
// Phi through erfc(z) = t (a1 + t (a2 + ... a5 t)) exp(-z^2), t = 1 / (1 + p z), z = |x| / sqrt 2 (Abramowitz & Stegun
// 7.1.26, |error| <= 1.5e-7): 11 packed operations, two MUFU (rcp, ex2) and a select per value instead of ~25 scalar
// instructions for erff; max |error| of the GELU against FP64 over [-12, 12]: 4.2e-7 (numpy restatement of these lines).
#ifndef ES_GELU
#define ES_GELU 1            // 0: erff (CUDA math library), 1: the packed form; 2: identity (timing experiments only)
#endif
__device__ __forceinline__ float2 es_gelu2(float2 x) {
#if ES_GELU == 0
  return make_float2(0.5f * x.x * (1.0f + erff(x.x * 0.70710678118654752440f)), 0.5f * x.y * (1.0f + erff(x.y * 0.70710678118654752440f)));
#elif ES_GELU == 2
  return x;
#else
  constexpr float P = 0.3275911f * 0.70710678118654752440f, C2 = 0.84932180028801904272f;   // sqrt(log2(e) / 2)
  constexpr float A1 = 0.5f * 0.254829592f, A2 = 0.5f * -0.284496736f, A3 = 0.5f * 1.421413741f, A4 = 0.5f * -1.453152027f,
                  A5 = 0.5f * 1.061405429f;
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = __ffma2_rn(ax, make_float2(P, P), make_float2(1.0f, 1.0f));
  const float2 u = __fmul2_rn(ax, make_float2(C2, C2));
  const float2 w = __fmul2_rn(u, u);
  float2 t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(-w.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(-w.y));
  float2 q = __ffma2_rn(t, make_float2(A5, A5), make_float2(A4, A4));
  q = __ffma2_rn(q, t, make_float2(A3, A3));
  q = __ffma2_rn(q, t, make_float2(A2, A2));
  q = __ffma2_rn(q, t, make_float2(A1, A1));
  const float2 h = __fmul2_rn(__fmul2_rn(q, t), e);                      // Phi(-|x|)
  const float2 phi = make_float2(x.x >= 0.0f ? 1.0f - h.x : h.x, x.y >= 0.0f ? 1.0f - h.y : h.y);
  return __fmul2_rn(x, phi);
#endif
}

struct EsGemm {
  int batch;          // clips
  int mtiles;         // 128-row tiles per clip (conv1: 24, conv2: 12)
  int m_valid;        // rows per clip (3000 / 1500)
  int kblocks;        // 64-wide K blocks (4 / 18)
  int kb_per_tap;     // K blocks per tap (conv1: 4 -> a single "tap" whose rows are the im2col rows; conv2: 6)
  const float* bias;  // [384]
  const float* pos;   // conv2: [1500][384] positional embedding; conv1: unused
  void* out;          // conv1: h, BF16 [batch][3002][384]; conv2: FP32 [batch][1500][384]
};

// Shared memory.  conv2 (MODE 1): a ring of ES_STAGES x {A 16 KB, W 24 KB}.  conv1 (MODE 0): K is only four blocks, so the
// CTA keeps its channel half of W1 (4 x 24 KB) resident for the whole kernel and the ring holds A tiles only.  Then one
// staging buffer per epilogue warp (a 32 x 32 chunk of the output: rows leave as full 128-byte / 64-byte segments), the
// mbarriers and the tensor-memory address.
constexpr int ES_RING1 = ES_STAGES * ES_STAGE_BYTES;                         // MODE 1
constexpr int ES_W1_BYTES = (ES_K1 / ES_BK) * ES_B_BYTES;                    // MODE 0: resident W1 half, 96 KB
constexpr int ES_RING0 = ES_W1_BYTES + ES_STAGES * ES_A_BYTES;
constexpr int ES_STG_OFF = ES_RING1 > ES_RING0 ? ES_RING1 : ES_RING0;
constexpr int ES_STG_BYTES = 32 * 32 * 4;
constexpr int ES_BAR_OFF = ES_STG_OFF + ES_EPI_WARPS * ES_STG_BYTES;
static_assert(ES_BAR_OFF + 256 + 1024 <= ES_SMEM_BYTES, "shared-memory budget of es_gemm_kernel");

template <int MODE>   // 0: conv1 -> h (BF16), 1: conv2 -> hidden states (FP32, + positions)
__global__ void __launch_bounds__(ES_THREADS, 1)
es_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const EsGemm p) {
  extern __shared__ unsigned char es_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)es_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + ES_BAR_OFF);
  unsigned long long* full = bars;                         // [ES_STAGES] TMA -> MMA
  unsigned long long* empty = bars + ES_STAGES;            // [ES_STAGES] MMA -> TMA
  unsigned long long* acc_full = bars + 2 * ES_STAGES;     // [2] MMA -> epilogue
  unsigned long long* acc_empty = bars + 2 * ES_STAGES + 2;   // [2] epilogue -> MMA
  unsigned long long* w_full = bars + 2 * ES_STAGES + 4;   // MODE 0: the resident W1 half has landed
  unsigned* s_tmem = reinterpret_cast<unsigned*>(bars + 2 * ES_STAGES + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // MODE 1: tiles (clip, 128 rows, channel half) round robin over the CTAs, the two halves of a row block side by side.
  // MODE 0: the CTA's channel half is fixed (blockIdx & 1; the grid is even) and it walks row blocks only.
  constexpr int NH = ES_D / ES_BN;
  const int t_first = MODE == 0 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = MODE == 0 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_count = MODE == 0 ? p.batch * p.mtiles : p.batch * p.mtiles * NH;
  auto decode = [&](int it, int& nt, int& mt, int& clip) {
    const int lin = MODE == 0 ? it : it / NH;
    nt = MODE == 0 ? (int)(blockIdx.x & 1) : it - lin * NH;
    clip = lin / p.mtiles;
    mt = lin - clip * p.mtiles;
  };
  constexpr int A_STRIDE = MODE == 0 ? ES_A_BYTES : ES_STAGE_BYTES;          // distance between ring slots
  unsigned char* ring = smem + (MODE == 0 ? ES_W1_BYTES : 0);

  if (warp == 0 && lane == 0) {
    es_prefetch_tmap(&tm_a);
    es_prefetch_tmap(&tm_w);
    for (int i = 0; i < ES_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, ES_EPI_WARPS); }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(ES_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  es_tc_fence_before();
  __syncthreads();
  es_tc_fence_after();
  const unsigned tmem_base = *s_tmem;
  // everything above overlaps the tail of the previous kernel of the stream; its results are needed from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (MODE == 0) {                                  // weights first: they do not depend on the previous kernel's output
        mbar_arrive_expect_tx(w_full, ES_W1_BYTES);
        for (int kb = 0; kb < ES_K1 / ES_BK; ++kb)
          es_tma_load_2d(smem + kb * ES_B_BYTES, &tm_w, kb * ES_BK, (int)(blockIdx.x & 1) * ES_BN, w_full);
      }
      int stage = 0; unsigned phase = 0;
      for (int it = t_first; it < t_count; it += t_step) {
        int nt, mt, clip;
        decode(it, nt, mt, clip);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty + stage, phase ^ 1u);
          unsigned char* sa = ring + stage * A_STRIDE;
          mbar_arrive_expect_tx(full + stage, MODE == 0 ? ES_A_BYTES : ES_STAGE_BYTES);
          const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * ES_BK;
          es_tma_load_4d(sa, &tm_a, c0, tap & 1, mt * ES_BM + (tap >> 1), clip, full + stage);
          if (MODE == 1) es_tma_load_2d(sa + ES_A_BYTES, &tm_w, kb * ES_BK, nt * ES_BN, full + stage);
          if (++stage == ES_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int stage = 0; unsigned phase = 0;
    int acc = 0; unsigned acc_phase = 0;
    if (MODE == 0) mbar_wait(w_full, 0);
    for (int it = t_first; it < t_count; it += t_step) {
      mbar_wait(acc_empty + acc, acc_phase ^ 1u);
      es_tc_fence_after();
      const unsigned d_tmem = tmem_base + (unsigned)(acc * ES_ACC_COLS);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(full + stage, phase);
        es_tc_fence_after();
        if (lane == 0) {
          const unsigned sa = smem_u32(ring + stage * A_STRIDE);
          const unsigned sb = MODE == 0 ? smem_u32(smem + kb * ES_B_BYTES) : sa + ES_A_BYTES;
          const unsigned long long adesc = es_smem_desc(sa), bdesc = es_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < ES_BK / ES_UK; ++k)     // 32 bytes further along K inside the swizzle atom: + 2 in the address field
            es_umma(d_tmem, adesc + 2u * k, bdesc + 2u * k, (kb | k) != 0);
          es_umma_commit(empty + stage);              // the stage is free once these MMAs have read it
        }
        __syncwarp();
        if (++stage == ES_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (lane == 0) es_umma_commit(acc_full + acc);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31, columns 64 ((w - 2) / 4) .. + 63 of the accumulator =====
    // A lane owns a ROW of the accumulator; rows are 768 / 1536 bytes apart in memory, so a chunk of 32 rows x 32 columns
    // goes through the warp's staging buffer (16-byte slots, XOR-swizzled so that neither side has bank conflicts) and
    // leaves with a quarter-warp (conv2) / four lanes (conv1) per row: full 128-byte / 64-byte segments, and the
    // positional embedding is read the same way.
    const int q = warp & 3, part = (warp - 2) >> 2;
    unsigned char* stg = smem + ES_STG_OFF + (warp - 2) * ES_STG_BYTES;
    int acc = 0; unsigned acc_phase = 0;
    for (int it = t_first; it < t_count; it += t_step) {
      int nt, mt, clip;
      decode(it, nt, mt, clip);
      const int m0 = mt * ES_BM + q * 32;                      // first row of this warp's chunk
      mbar_wait(acc_full + acc, acc_phase);
      es_tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col = part * 64 + c * 32, n0 = nt * ES_BN + col;
        float v[32];
        es_tmem_ld32(tmem_base + (unsigned)(acc * ES_ACC_COLS + col) + ((unsigned)(q * 32) << 16), v);
        if (c == 1) {                               // the accumulator is in registers: hand it back to the MMA warp
          es_tc_fence_before();
          __syncwarp();
          if (lane == 0) es_mbar_arrive(acc_empty + acc);
        }
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
        if (MODE == 0) {
          uint4* s4 = reinterpret_cast<uint4*>(stg);            // [32 rows][4 slots of 8 BF16]
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
            const float2 g0 = es_gelu2(__fadd2_rn(make_float2(v[8 * g], v[8 * g + 1]), make_float2(ba.x, ba.y)));
            const float2 g1 = es_gelu2(__fadd2_rn(make_float2(v[8 * g + 2], v[8 * g + 3]), make_float2(ba.z, ba.w)));
            const float2 g2 = es_gelu2(__fadd2_rn(make_float2(v[8 * g + 4], v[8 * g + 5]), make_float2(bb.x, bb.y)));
            const float2 g3 = es_gelu2(__fadd2_rn(make_float2(v[8 * g + 6], v[8 * g + 7]), make_float2(bb.z, bb.w)));
            __nv_bfloat162 o0 = __floats2bfloat162_rn(g0.x, g0.y), o1 = __floats2bfloat162_rn(g1.x, g1.y);
            __nv_bfloat162 o2 = __floats2bfloat162_rn(g2.x, g2.y), o3 = __floats2bfloat162_rn(g3.x, g3.y);
            uint4 o;
            o.x = *reinterpret_cast<unsigned*>(&o0); o.y = *reinterpret_cast<unsigned*>(&o1);
            o.z = *reinterpret_cast<unsigned*>(&o2); o.w = *reinterpret_cast<unsigned*>(&o3);
            s4[lane * 4 + (g ^ ((lane >> 1) & 3))] = o;
          }
          __syncwarp();
          __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(p.out) + ((size_t)clip * ES_HROWS + 1) * ES_D + n0;
#pragma unroll
          for (int r8 = 0; r8 < 4; ++r8) {
            const int row = r8 * 8 + (lane >> 2), j = lane & 3;
            const uint4 o = s4[row * 4 + (j ^ ((row >> 1) & 3))];
            if (m0 + row < p.m_valid) reinterpret_cast<uint4*>(hb + (size_t)(m0 + row) * ES_D)[j] = o;
          }
          __syncwarp();
        } else {
          float4* s4 = reinterpret_cast<float4*>(stg);          // [32 rows][8 slots of 4 floats]
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b = __ldg(b4 + g);
            const float2 g0 = es_gelu2(__fadd2_rn(make_float2(v[4 * g], v[4 * g + 1]), make_float2(b.x, b.y)));
            const float2 g1 = es_gelu2(__fadd2_rn(make_float2(v[4 * g + 2], v[4 * g + 3]), make_float2(b.z, b.w)));
            s4[lane * 8 + (g ^ (lane & 7))] = make_float4(g0.x, g0.y, g1.x, g1.y);
          }
          __syncwarp();
          float* ob = reinterpret_cast<float*>(p.out) + (size_t)clip * ES_T2 * ES_D + n0;
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4) {
            const int row = r4 * 4 + (lane >> 3), j = lane & 7;
            const float4 o = s4[row * 8 + (j ^ (row & 7))];
            if (m0 + row < p.m_valid) {
              const float4 e = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)(m0 + row) * ES_D + n0) + j);
              reinterpret_cast<float4*>(ob + (size_t)(m0 + row) * ES_D)[j] = make_float4(o.x + e.x, o.y + e.y, o.z + e.z, o.w + e.w);
            }
          }
          __syncwarp();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  es_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    es_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ES_TMEM_COLS) : "memory");
  }
}
