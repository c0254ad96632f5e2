// CPU unit test of csrc/fft_codelets.cuh: the two-pass prime-factor 400-point real DFT built
// from real_dft25 + cplx_dft16 must reproduce |X[k]|^2 of a naive O(N^2) DFT for bins 0..200.
// Built and run by tests/test_codelets.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../audio_transformers_b200/csrc/fft_codelets.cuh"

template <class V>
static void frame_power_pfa(const V* x, const V* w, V* power /*201*/) {
  static V E[16][25];
  for (int a = 0; a < 16; ++a) {
    V xin[25], win[25];
    for (int b = 0; b < 25; ++b) { xin[b] = x[b2::pfa400_n(a, b)]; win[b] = w[b2::pfa400_n(a, b)]; }
    b2::real_dft25(xin, win, E[a]);
  }
  for (int k = 0; k <= 200; ++k) power[k] = (V)-1;
  for (int k2 = 0; k2 <= 12; ++k2) {
    V yr[16], yi[16], Xr[16], Xi[16];
    for (int a = 0; a < 16; ++a) {
      yr[a] = k2 == 0 ? E[a][0] : E[a][2 * k2 - 1];
      yi[a] = k2 == 0 ? (V)0 : E[a][2 * k2];
    }
    if (k2 == 0) {     // the kernel's real-input codelet
      V P[9];
      b2::real_dft16_power(yr, P);
      for (int k1 = 0; k1 < 9; ++k1) {
        int bin = b2::pfa400_bin(k1, 0);
        if (power[bin] != (V)-1) { printf("bin %d written twice\n", bin); exit(1); }
        power[bin] = P[k1];
      }
      continue;
    }
    b2::cplx_dft16(yr, yi, Xr, Xi);
    for (int k1 = 0; k1 < 16; ++k1) {
      int bin = b2::pfa400_bin(k1, k2);
      if (power[bin] != (V)-1) { printf("bin %d written twice\n", bin); exit(1); }
      power[bin] = Xr[k1] * Xr[k1] + Xi[k1] * Xi[k1];
    }
  }
  for (int k = 0; k <= 200; ++k) if (power[k] == (V)-1) { printf("bin %d never written\n", k); exit(1); }
}

int main() {
  const int N = 400;
  std::vector<double> x(N), w(N), ref(201);
  srand(1234);
  double worst32 = 0, worst64 = 0;
  for (int trial = 0; trial < 6; ++trial) {
    for (int n = 0; n < N; ++n) {
      w[n] = 0.5 - 0.5 * cos(2 * M_PI * n / N);
      double noise = (rand() / (double)RAND_MAX - 0.5);
      x[n] = trial == 0 ? noise : trial == 1 ? sin(2 * M_PI * 440.0 * n / 16000.0) : trial == 2 ? (n == 137) :
             trial == 3 ? 1.0 : trial == 4 ? 0.5 * sin(2 * M_PI * 3999.0 * n / 16000.0) + 1e-3 * noise : noise * noise * noise;
    }
    double pmax = 0;
    for (int k = 0; k <= 200; ++k) {
      double re = 0, im = 0;
      for (int n = 0; n < N; ++n) { double a = -2 * M_PI * (double)((long)n * k % N) / N; re += x[n] * w[n] * cos(a); im += x[n] * w[n] * sin(a); }
      ref[k] = re * re + im * im; pmax = fmax(pmax, ref[k]);
    }
    std::vector<double> p64(201); frame_power_pfa<double>(x.data(), w.data(), p64.data());
    std::vector<float> xf(x.begin(), x.end()), wf(w.begin(), w.end()), p32(201);
    frame_power_pfa<float>(xf.data(), wf.data(), p32.data());
    for (int k = 0; k <= 200; ++k) {
      worst64 = fmax(worst64, fabs(p64[k] - ref[k]) / pmax);
      worst32 = fmax(worst32, fabs((double)p32[k] - ref[k]) / pmax);
    }
  }
  printf("worst |dP|/Pmax  fp64 %.3e  fp32 %.3e\n", worst64, worst32);
  if (worst64 > 1e-13 || worst32 > 2e-6) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
