// fft_codelets.cuh -- straight-line, register-resident DFT codelets for the log-mel kernels.
//
// Everything here is a pure function of its arguments, usable from host code (the CPU unit test
// tests/cpu/test_codelets.cpp builds it with g++) and from device code.  `V` is the value type a
// thread carries per logical real number: `float` (one frame per lane) -- the kernels keep one
// *frame* per lane, so every index-dependent quantity (twiddles, window taps, filter weights) is
// warp-uniform and is encoded as an immediate / constant-bank operand, never loaded per lane.
//
// Transform sign convention: forward DFT, X[k] = sum_n x[n] exp(-2 pi i n k / N), the one
// torch.stft uses (HF:models/whisper/feature_extraction_whisper.py:149, TA:functional/functional.py:123).
//
// Whisper frame (N = 400 = 16 x 25, coprime) uses the Good-Thomas prime-factor map, which needs NO
// twiddle factors between the two passes:
//     n = (25 a + 16 b) mod 400,  a in [0,16), b in [0,25)
//     k = (225 k1 + 176 k2) mod 400,  k1 = k mod 16, k2 = k mod 25
//     X[k] = sum_a W16^(a k1) * ( sum_b W25^(b k2) x[n(a,b)] )
// Pass 1 (per a):   real 25-point DFT, keeping k2 = 0..12 (the input is real, the rest is the
//                   conjugate mirror)                                          -> real_dft25
// Pass 2 (per k2):  complex 16-point DFT over a                                -> cplx_dft16
// Since only |X|^2 is needed and |X[400-k]| = |X[k]|, pass 2 for k2 = 1..12 (all k1) plus
// k2 = 0 (k1 = 0..8) covers bins 0..200 exactly once.
#pragma once

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_CX __host__ __device__ constexpr
#else
#define B2_HD inline
#define B2_CX constexpr
#endif

#include <math.h>

namespace b2 {

// ---- value-type primitives (float) ---------------------------------------------------------
B2_HD float vadd(float a, float b) { return a + b; }
B2_HD float vsub(float a, float b) { return a - b; }
B2_HD float vneg(float a) { return -a; }
B2_HD float vmulc(float a, float c) { return a * c; }
#if defined(__CUDA_ARCH__)
B2_HD float vfmac(float a, float c, float b) { return __fmaf_rn(a, c, b); }   // a*c + b
#else
B2_HD float vfmac(float a, float c, float b) { return fmaf(a, c, b); }
#endif
B2_HD float vfmsc(float a, float c, float b) { return vfmac(a, c, -b); }      // a*c - b
B2_HD float vmul(float a, float b) { return a * b; }
B2_HD float vfma(float a, float b, float c) { return vfmac(a, b, c); }

// double overloads: the CPU unit test runs the same codelets in FP64 to separate "wrong index
// map" from "FP32 round-off".
B2_HD double vadd(double a, double b) { return a + b; }
B2_HD double vsub(double a, double b) { return a - b; }
B2_HD double vneg(double a) { return -a; }
B2_HD double vmulc(double a, double c) { return a * c; }
B2_HD double vfmac(double a, double c, double b) { return a * c + b; }
B2_HD double vfmsc(double a, double c, double b) { return a * c - b; }
B2_HD double vmul(double a, double b) { return a * b; }
B2_HD double vfma(double a, double b, double c) { return a * b + c; }

// packed pair of FP32 values (two frames per lane): Blackwell's FADD2 / FMUL2 / FFMA2 process both
// halves in one issue slot and accept a scalar immediate / uniform register broadcast to both halves
// and a negate modifier, so the packed codelets cost exactly half the issue slots of the scalar ones.
#if defined(__CUDACC__)
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ float2 vmulc(float2 a, float c) { return __fmul2_rn(a, make_float2(c, c)); }
__device__ __forceinline__ float2 vfmac(float2 a, float c, float2 b) { return __ffma2_rn(a, make_float2(c, c), b); }
__device__ __forceinline__ float2 vfmsc(float2 a, float c, float2 b) { return __ffma2_rn(a, make_float2(c, c), vneg(b)); }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif

// ---- constants ---------------------------------------------------------------------------------
#define B2_C5_1 0.30901699437494742410    // cos(2 pi / 5)
#define B2_C5_2 (-0.80901699437494742410) // cos(4 pi / 5)
#define B2_S5_1 0.95105651629515357212    // sin(2 pi / 5)
#define B2_S5_2 0.58778525229247312917    // sin(4 pi / 5)
#define B2_SQRT1_2 0.70710678118654752440

// cos/sin(2 pi m / 25), m = 1..8
#define B2_C25_1 0.96858316112863108
#define B2_S25_1 0.24868988716485479
#define B2_C25_2 0.87630668004386358
#define B2_S25_2 0.48175367410171532
#define B2_C25_3 0.72896862742141155
#define B2_S25_3 0.68454710592868873
#define B2_C25_4 0.53582679497899666
#define B2_S25_4 0.84432792550201508
#define B2_C25_6 0.06279051952931337
#define B2_S25_6 0.99802672842827156
#define B2_C25_8 (-0.42577929156507272)
#define B2_S25_8 0.90482705246601958

// cos/sin(2 pi m / 16), m = 1, 3
#define B2_C16_1 0.92387953251128675613
#define B2_S16_1 0.38268343236508977173

template <class V> struct real_t { typedef float type; };
template <> struct real_t<double> { typedef double type; };
#define B2_K(V, x) ((typename real_t<V>::type)(x))

// ---- 5-point DFT of REAL input ----------------------------------------------------------------
// y0 real; Y1 = y1r - i*u1; Y2 = y2r - i*u2; (Y3 = conj Y2, Y4 = conj Y1).   14 flops-instr.
template <class V>
B2_HD void rdft5(V a0, V a1, V a2, V a3, V a4, V& y0, V& y1r, V& u1, V& y2r, V& u2) {
  V t1 = vadd(a1, a4), t2 = vadd(a2, a3), t3 = vsub(a1, a4), t4 = vsub(a2, a3);
  y0 = vadd(a0, vadd(t1, t2));
  y1r = vfmac(t2, B2_K(V, B2_C5_2), vfmac(t1, B2_K(V, B2_C5_1), a0));
  y2r = vfmac(t2, B2_K(V, B2_C5_1), vfmac(t1, B2_K(V, B2_C5_2), a0));
  u1 = vfmac(t4, B2_K(V, B2_S5_2), vmulc(t3, B2_K(V, B2_S5_1)));
  u2 = vfmac(t4, B2_K(V, -B2_S5_1), vmulc(t3, B2_K(V, B2_S5_2)));
}

// Same, with a window tap folded into each input (x_j * w_j): 17 instr instead of 14 + 5.
template <class V, class W>
B2_HD void rdft5w(V x0, V x1, V x2, V x3, V x4, W w0, W w1, W w2, W w3, W w4,
                  V& y0, V& y1r, V& u1, V& y2r, V& u2) {
  V p0 = vmulc(x0, w0), p4 = vmulc(x4, w4), p3 = vmulc(x3, w3);
  V t1 = vfmac(x1, w1, p4), t3 = vfmsc(x1, w1, p4);
  V t2 = vfmac(x2, w2, p3), t4 = vfmsc(x2, w2, p3);
  y0 = vadd(p0, vadd(t1, t2));
  y1r = vfmac(t2, B2_K(V, B2_C5_2), vfmac(t1, B2_K(V, B2_C5_1), p0));
  y2r = vfmac(t2, B2_K(V, B2_C5_1), vfmac(t1, B2_K(V, B2_C5_2), p0));
  u1 = vfmac(t4, B2_K(V, B2_S5_2), vmulc(t3, B2_K(V, B2_S5_1)));
  u2 = vfmac(t4, B2_K(V, -B2_S5_1), vmulc(t3, B2_K(V, B2_S5_2)));
}

// ---- 5-point DFT of COMPLEX input (zr + i zi) -> (Yr + i Yi), all five outputs ------------------
template <class V>
B2_HD void cdft5(const V zr[5], const V zi[5], V Yr[5], V Yi[5]) {
  V a0, a1r, ua1, a2r, ua2, b0, b1r, ub1, b2r, ub2;
  rdft5(zr[0], zr[1], zr[2], zr[3], zr[4], a0, a1r, ua1, a2r, ua2);
  rdft5(zi[0], zi[1], zi[2], zi[3], zi[4], b0, b1r, ub1, b2r, ub2);
  Yr[0] = a0;               Yi[0] = b0;
  Yr[1] = vadd(a1r, ub1);   Yi[1] = vsub(b1r, ua1);
  Yr[4] = vsub(a1r, ub1);   Yi[4] = vadd(b1r, ua1);
  Yr[2] = vadd(a2r, ub2);   Yi[2] = vsub(b2r, ua2);
  Yr[3] = vsub(a2r, ub2);   Yi[3] = vadd(b2r, ua2);
}

// ---- 25-point DFT of REAL (windowed) input, outputs k = 0..12 ----------------------------------
// x[b], b = 0..24 in natural DFT order; w[b] the window taps for those samples.
// out[0] = X0 (real); out[2k-1] = Re X_k, out[2k] = Im X_k for k = 1..12.
// second stage, given the five first-stage 5-point transforms (see real_dft25)
template <class V>
B2_HD void real_dft25_stage2(const V Y0[5], const V Y1r[5], const V U1[5], const V Y2r[5], const V U2[5], V out[25]) {
  // s = 0: real 5-point DFT of Y0 -> X0, X5, X10
  {
    V y0, y1r, u1, y2r, u2;
    rdft5(Y0[0], Y0[1], Y0[2], Y0[3], Y0[4], y0, y1r, u1, y2r, u2);
    out[0] = y0;
    out[2 * 5 - 1] = y1r;  out[2 * 5] = vneg(u1);
    out[2 * 10 - 1] = y2r; out[2 * 10] = vneg(u2);
  }
  // s = 1: Z_r = W25^r * (Y1r - i U1);  W25^m = c - i s  ->  Z = (c*yr - s*u) - i (c*u + s*yr)
  {
    const typename real_t<V>::type c[5] = {1, B2_K(V, B2_C25_1), B2_K(V, B2_C25_2), B2_K(V, B2_C25_3), B2_K(V, B2_C25_4)};
    const typename real_t<V>::type s[5] = {0, B2_K(V, B2_S25_1), B2_K(V, B2_S25_2), B2_K(V, B2_S25_3), B2_K(V, B2_S25_4)};
    V zr[5], zi[5], Xr[5], Xi[5];
    zr[0] = Y1r[0]; zi[0] = vneg(U1[0]);
#pragma unroll
    for (int r = 1; r < 5; ++r) {
      zr[r] = vfmac(U1[r], -s[r], vmulc(Y1r[r], c[r]));
      zi[r] = vneg(vfmac(Y1r[r], s[r], vmulc(U1[r], c[r])));
    }
    cdft5(zr, zi, Xr, Xi);     // X[1], X[6], X[11], X[16], X[21]
    out[2 * 1 - 1] = Xr[0];  out[2 * 1] = Xi[0];
    out[2 * 6 - 1] = Xr[1];  out[2 * 6] = Xi[1];
    out[2 * 11 - 1] = Xr[2]; out[2 * 11] = Xi[2];
    out[2 * 9 - 1] = Xr[3];  out[2 * 9] = vneg(Xi[3]);   // X9 = conj X16
    out[2 * 4 - 1] = Xr[4];  out[2 * 4] = vneg(Xi[4]);   // X4 = conj X21
  }
  // s = 2: Z_r = W25^(2r) * (Y2r - i U2)
  {
    const typename real_t<V>::type c[5] = {1, B2_K(V, B2_C25_2), B2_K(V, B2_C25_4), B2_K(V, B2_C25_6), B2_K(V, B2_C25_8)};
    const typename real_t<V>::type s[5] = {0, B2_K(V, B2_S25_2), B2_K(V, B2_S25_4), B2_K(V, B2_S25_6), B2_K(V, B2_S25_8)};
    V zr[5], zi[5], Xr[5], Xi[5];
    zr[0] = Y2r[0]; zi[0] = vneg(U2[0]);
#pragma unroll
    for (int r = 1; r < 5; ++r) {
      zr[r] = vfmac(U2[r], -s[r], vmulc(Y2r[r], c[r]));
      zi[r] = vneg(vfmac(Y2r[r], s[r], vmulc(U2[r], c[r])));
    }
    cdft5(zr, zi, Xr, Xi);     // X[2], X[7], X[12], X[17], X[22]
    out[2 * 2 - 1] = Xr[0];  out[2 * 2] = Xi[0];
    out[2 * 7 - 1] = Xr[1];  out[2 * 7] = Xi[1];
    out[2 * 12 - 1] = Xr[2]; out[2 * 12] = Xi[2];
    out[2 * 8 - 1] = Xr[3];  out[2 * 8] = vneg(Xi[3]);   // X8 = conj X17
    out[2 * 3 - 1] = Xr[4];  out[2 * 3] = vneg(Xi[4]);   // X3 = conj X22
  }
}

template <class V, class W>
B2_HD void real_dft25(const V x[25], const W w[25], V out[25]) {
  V Y0[5], Y1r[5], U1[5], Y2r[5], U2[5];
#pragma unroll
  for (int r = 0; r < 5; ++r)
    rdft5w(x[r], x[r + 5], x[r + 10], x[r + 15], x[r + 20], w[r], w[r + 5], w[r + 10], w[r + 15], w[r + 20],
           Y0[r], Y1r[r], U1[r], Y2r[r], U2[r]);
  real_dft25_stage2(Y0, Y1r, U1, Y2r, U2, out);
}

// ---- 4-point complex DFT (in place on 4 re/im pairs) --------------------------------------------
template <class V>
B2_HD void cdft4(V& r0, V& i0, V& r1, V& i1, V& r2, V& i2, V& r3, V& i3) {
  V p0r = vadd(r0, r2), p0i = vadd(i0, i2), p1r = vsub(r0, r2), p1i = vsub(i0, i2);
  V p2r = vadd(r1, r3), p2i = vadd(i1, i3), p3r = vsub(r1, r3), p3i = vsub(i1, i3);
  r0 = vadd(p0r, p2r); i0 = vadd(p0i, p2i);
  r2 = vsub(p0r, p2r); i2 = vsub(p0i, p2i);
  r1 = vadd(p1r, p3i); i1 = vsub(p1i, p3r);      // p1 - i p3
  r3 = vsub(p1r, p3i); i3 = vadd(p1i, p3r);      // p1 + i p3
}

// multiply (r + i m) by exp(-2 pi i e / 16), e compile-time
template <int E, class V>
B2_HD void twiddle16(V& r, V& m) {
  typedef typename real_t<V>::type R;
  constexpr int e = ((E % 16) + 16) % 16;
  if (e == 0) return;
  if (e == 4) { V t = r; r = m; m = vneg(t); return; }                         // * -i
  if (e == 8) { r = vneg(r); m = vneg(m); return; }
  if (e == 12) { V t = r; r = vneg(m); m = t; return; }                        // * +i
  if (e == 2) { V t = vadd(r, m), u = vsub(m, r); r = vmulc(t, (R)B2_SQRT1_2); m = vmulc(u, (R)B2_SQRT1_2); return; }
  if (e == 6) { V t = vsub(m, r), u = vadd(r, m); r = vmulc(t, (R)B2_SQRT1_2); m = vmulc(u, (R)-B2_SQRT1_2); return; }
  // generic: (c - i s)(r + i m) = (c r + s m) + i (c m - s r)
  R c = 0, s = 0;
  if (e == 1) { c = (R)B2_C16_1; s = (R)B2_S16_1; }
  if (e == 3) { c = (R)B2_S16_1; s = (R)B2_C16_1; }
  if (e == 9) { c = (R)-B2_C16_1; s = (R)-B2_S16_1; }
  if (e == 5) { c = (R)-B2_S16_1; s = (R)B2_C16_1; }
  if (e == 7) { c = (R)-B2_C16_1; s = (R)B2_S16_1; }
  V nr = vfmac(m, s, vmulc(r, c));
  V nm = vfmac(r, -s, vmulc(m, c));
  r = nr; m = nm;
}

// ---- 16-point complex DFT, radix 4 x 4 ---------------------------------------------------------
// in: yr[a], yi[a], a = 0..15 (natural order); out: Xr[k], Xi[k], k = 0..15 (natural order).
template <class V>
B2_HD void cplx_dft16(const V yr[16], const V yi[16], V Xr[16], V Xi[16]) {
  V tr[4][4], ti[4][4];   // [a2][c1]
#pragma unroll
  for (int a2 = 0; a2 < 4; ++a2) {
    V r0 = yr[a2], i0 = yi[a2], r1 = yr[4 + a2], i1 = yi[4 + a2];
    V r2 = yr[8 + a2], i2 = yi[8 + a2], r3 = yr[12 + a2], i3 = yi[12 + a2];
    cdft4(r0, i0, r1, i1, r2, i2, r3, i3);
    tr[a2][0] = r0; ti[a2][0] = i0; tr[a2][1] = r1; ti[a2][1] = i1;
    tr[a2][2] = r2; ti[a2][2] = i2; tr[a2][3] = r3; ti[a2][3] = i3;
  }
  twiddle16<1>(tr[1][1], ti[1][1]); twiddle16<2>(tr[1][2], ti[1][2]); twiddle16<3>(tr[1][3], ti[1][3]);
  twiddle16<2>(tr[2][1], ti[2][1]); twiddle16<4>(tr[2][2], ti[2][2]); twiddle16<6>(tr[2][3], ti[2][3]);
  twiddle16<3>(tr[3][1], ti[3][1]); twiddle16<6>(tr[3][2], ti[3][2]); twiddle16<9>(tr[3][3], ti[3][3]);
#pragma unroll
  for (int c1 = 0; c1 < 4; ++c1) {
    V r0 = tr[0][c1], i0 = ti[0][c1], r1 = tr[1][c1], i1 = ti[1][c1];
    V r2 = tr[2][c1], i2 = ti[2][c1], r3 = tr[3][c1], i3 = ti[3][c1];
    cdft4(r0, i0, r1, i1, r2, i2, r3, i3);
    Xr[c1] = r0; Xi[c1] = i0; Xr[c1 + 4] = r1; Xi[c1 + 4] = i1;
    Xr[c1 + 8] = r2; Xi[c1 + 8] = i2; Xr[c1 + 12] = r3; Xi[c1 + 12] = i3;
  }
}

// ---- 8-point DFT of REAL input, outputs k = 0..4 -------------------------------------------------
// X0, X4 real; X1, X2, X3 complex (Xr, Xi); X[8-k] = conj X[k].
template <class V>
B2_HD void rdft8(const V x[8], V& X0, V& X4, V& X1r, V& X1i, V& X2r, V& X2i, V& X3r, V& X3i) {
  typedef typename real_t<V>::type R;
  V s04 = vadd(x[0], x[4]), d04 = vsub(x[0], x[4]), s26 = vadd(x[2], x[6]), d26 = vsub(x[2], x[6]);
  V s15 = vadd(x[1], x[5]), d15 = vsub(x[1], x[5]), s37 = vadd(x[3], x[7]), d37 = vsub(x[3], x[7]);
  V ee = vadd(s04, s26), eo = vadd(s15, s37);
  X0 = vadd(ee, eo); X4 = vsub(ee, eo);
  X2r = vsub(s04, s26); X2i = vsub(s37, s15);
  V p = vsub(d15, d37), q = vadd(d15, d37);
  X1r = vfmac(p, (R)B2_SQRT1_2, d04);  X1i = vneg(vfmac(q, (R)B2_SQRT1_2, d26));
  X3r = vfmac(p, (R)-B2_SQRT1_2, d04); X3i = vfmac(q, (R)-B2_SQRT1_2, d26);
}

// ---- |X[k]|^2, k = 0..8, of the 16-point DFT of REAL input --------------------------------------
// (pass 2 of the Whisper frame for k2 = 0, where the pass-1 outputs are real).  Even/odd split:
// X[k] = E[k] + W16^k O[k], X[8-k] = conj(E[k] - W16^k O[k]) for k = 1..3.
template <class V>
B2_HD void real_dft16_power(const V y[16], V P[9]) {
  typedef typename real_t<V>::type R;
  V e[8], o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { e[j] = y[2 * j]; o[j] = y[2 * j + 1]; }
  V E0, E4, E1r, E1i, E2r, E2i, E3r, E3i, O0, O4, O1r, O1i, O2r, O2i, O3r, O3i;
  rdft8(e, E0, E4, E1r, E1i, E2r, E2i, E3r, E3i);
  rdft8(o, O0, O4, O1r, O1i, O2r, O2i, O3r, O3i);
  V x0 = vadd(E0, O0), x8 = vsub(E0, O0);
  P[0] = vmul(x0, x0); P[8] = vmul(x8, x8);
  P[4] = vfma(E4, E4, vmul(O4, O4));                 // X4 = E4 - i O4
  twiddle16<1>(O1r, O1i); twiddle16<2>(O2r, O2i); twiddle16<3>(O3r, O3i);
  { V ar = vadd(E1r, O1r), ai = vadd(E1i, O1i), br = vsub(E1r, O1r), bi = vsub(E1i, O1i);
    P[1] = vfma(ar, ar, vmul(ai, ai)); P[7] = vfma(br, br, vmul(bi, bi)); }
  { V ar = vadd(E2r, O2r), ai = vadd(E2i, O2i), br = vsub(E2r, O2r), bi = vsub(E2i, O2i);
    P[2] = vfma(ar, ar, vmul(ai, ai)); P[6] = vfma(br, br, vmul(bi, bi)); }
  { V ar = vadd(E3r, O3r), ai = vadd(E3i, O3i), br = vsub(E3r, O3r), bi = vsub(E3i, O3i);
    P[3] = vfma(ar, ar, vmul(ai, ai)); P[5] = vfma(br, br, vmul(bi, bi)); }
}

// multiply (r + i m) by exp(-2 pi i E / 32), E compile-time (even E reuse the /16 table)
#define B2_C32_1 0.98078528040323044913
#define B2_S32_1 0.19509032201612826785
#define B2_C32_3 0.83146961230254523708
#define B2_S32_3 0.55557023301960222474
template <int E, class V>
B2_HD void twiddle32(V& r, V& m) {
  typedef typename real_t<V>::type R;
  constexpr int e = ((E % 32) + 32) % 32;
  if (e % 2 == 0) { twiddle16<e / 2>(r, m); return; }
  // odd e: angle = e * 11.25 deg; fold into the first octant
  R c = 0, s = 0;
  if (e == 1)  { c = (R)B2_C32_1;  s = (R)B2_S32_1; }
  if (e == 3)  { c = (R)B2_C32_3;  s = (R)B2_S32_3; }
  if (e == 5)  { c = (R)B2_S32_3;  s = (R)B2_C32_3; }
  if (e == 7)  { c = (R)B2_S32_1;  s = (R)B2_C32_1; }
  if (e == 9)  { c = (R)-B2_S32_1; s = (R)B2_C32_1; }
  if (e == 11) { c = (R)-B2_S32_3; s = (R)B2_C32_3; }
  if (e == 13) { c = (R)-B2_C32_3; s = (R)B2_S32_3; }
  if (e == 15) { c = (R)-B2_C32_1; s = (R)B2_S32_1; }
  if (e > 16)  {   // W^(e) = -W^(e-16)
    constexpr int f = e - 16;
    if (f == 1)  { c = (R)-B2_C32_1; s = (R)-B2_S32_1; }
    if (f == 3)  { c = (R)-B2_C32_3; s = (R)-B2_S32_3; }
    if (f == 5)  { c = (R)-B2_S32_3; s = (R)-B2_C32_3; }
    if (f == 7)  { c = (R)-B2_S32_1; s = (R)-B2_C32_1; }
    if (f == 9)  { c = (R)B2_S32_1;  s = (R)-B2_C32_1; }
    if (f == 11) { c = (R)B2_S32_3;  s = (R)-B2_C32_3; }
    if (f == 13) { c = (R)B2_C32_3;  s = (R)-B2_S32_3; }
    if (f == 15) { c = (R)B2_C32_1;  s = (R)-B2_S32_1; }
  }
  V nr = vfmac(m, s, vmulc(r, c));       // (c - i s)(r + i m) = (c r + s m) + i (c m - s r)
  V nm = vfmac(r, -s, vmulc(m, c));
  r = nr; m = nm;
}

template <int K, class V>
B2_HD void dft32_combine(const V Er[16], const V Ei[16], const V Or[16], const V Oi[16], V Xr[32], V Xi[32]) {
  V tr = Or[K], ti = Oi[K];
  twiddle32<K>(tr, ti);
  Xr[K] = vadd(Er[K], tr);      Xi[K] = vadd(Ei[K], ti);
  Xr[K + 16] = vsub(Er[K], tr); Xi[K + 16] = vsub(Ei[K], ti);
  if constexpr (K + 1 < 16) dft32_combine<K + 1>(Er, Ei, Or, Oi, Xr, Xi);
}

// ---- 32-point complex DFT: radix 2 x 16 (decimation in time) -----------------------------------
template <class V>
B2_HD void cplx_dft32(const V zr[32], const V zi[32], V Xr[32], V Xi[32]) {
  V er[16], ei[16], orr[16], oi[16], Er[16], Ei[16], Or[16], Oi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { er[j] = zr[2 * j]; ei[j] = zi[2 * j]; orr[j] = zr[2 * j + 1]; oi[j] = zi[2 * j + 1]; }
  cplx_dft16(er, ei, Er, Ei);
  cplx_dft16(orr, oi, Or, Oi);
  dft32_combine<0>(Er, Ei, Or, Oi, Xr, Xi);
}

template <int K, class V>
B2_HD void rdft32_untangle(const V Zr[16], const V Zi[16], V Xr[17], V Xi[17]) {
  // X[k] = A + W32^k * B,  A = (Z[k] + conj Z[16-k]) / 2,  B = (Z[k] - conj Z[16-k]) / (2i)
  typedef typename real_t<V>::type R;
  constexpr int K2 = (16 - K) % 16;
  V ar = vmulc(vadd(Zr[K], Zr[K2]), (R)0.5), ai = vmulc(vsub(Zi[K], Zi[K2]), (R)0.5);
  V br = vmulc(vadd(Zi[K], Zi[K2]), (R)0.5), bi = vmulc(vsub(Zr[K2], Zr[K]), (R)0.5);
  twiddle32<K>(br, bi);
  Xr[K] = vadd(ar, br); Xi[K] = vadd(ai, bi);
  if constexpr (K + 1 < 16) rdft32_untangle<K + 1>(Zr, Zi, Xr, Xi);
}

// ---- 32-point DFT of REAL (windowed) input, outputs k = 0..16 -----------------------------------
// Packs z[j] = x[2j] + i x[2j+1], one complex 16-point DFT, then the standard untangling step.
// xw: the 32 samples with the window already applied
template <class V>
B2_HD void real_dft32_windowed(const V xw[32], V Xr[17], V Xi[17]) {
  V zr[16], zi[16], Zr[16], Zi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { zr[j] = xw[2 * j]; zi[j] = xw[2 * j + 1]; }
  cplx_dft16(zr, zi, Zr, Zi);
  rdft32_untangle<0>(Zr, Zi, Xr, Xi);
  Xr[16] = vsub(Zr[0], Zi[0]);           // X[16] = sum_even - sum_odd (real)
  Xi[16] = vsub(Zr[0], Zr[0]);           // exact zero of the right type
}
template <class V, class W>
B2_HD void real_dft32(const V x[32], const W w[32], V Xr[17], V Xi[17]) {
  V xw[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) xw[j] = vmulc(x[j], w[j]);
  real_dft32_windowed(xw, Xr, Xi);
}

// ---- Good-Thomas index maps for N = 400 = 16 x 25 ----------------------------------------------
B2_CX int pfa400_n(int a, int b) { return (25 * a + 16 * b) % 400; }
B2_CX int pfa400_k(int k1, int k2) { return (225 * k1 + 176 * k2) % 400; }
B2_CX int pfa400_bin(int k1, int k2) { return pfa400_k(k1, k2) <= 200 ? pfa400_k(k1, k2) : 400 - pfa400_k(k1, k2); }

}  // namespace b2
