// CPU unit test of the 32-point codelets (urban preset): two-pass Cooley-Tukey 1024-point real DFT
// built from real_dft32 + twiddles + cplx_dft32 vs a naive DFT.  Built/run by tests/test_codelets.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../audio_transformers_b200/csrc/fft_codelets.cuh"

template <class V>
static void frame_power_1024(const V* x, const V* w, V* power /*513*/) {
  static V Er[32][17], Ei[32][17];
  for (int r = 0; r < 32; ++r) {
    V xin[32], win[32];
    for (int j = 0; j < 32; ++j) { xin[j] = x[r + 32 * j]; win[j] = w[r + 32 * j]; }
    b2::real_dft32(xin, win, Er[r], Ei[r]);
  }
  for (int k = 0; k <= 512; ++k) power[k] = (V)-1;
  for (int k2 = 0; k2 <= 16; ++k2) {
    V zr[32], zi[32], Xr[32], Xi[32];
    for (int r = 0; r < 32; ++r) {
      double ang = -2.0 * M_PI * (double)(r * k2) / 1024.0;
      V c = (V)cos(ang), s = (V)sin(ang);
      zr[r] = Er[r][k2] * c - Ei[r][k2] * s;
      zi[r] = Er[r][k2] * s + Ei[r][k2] * c;
    }
    b2::cplx_dft32(zr, zi, Xr, Xi);
    for (int k1 = 0; k1 < 32; ++k1) {
      int k = k2 + 32 * k1;
      int bin = k <= 512 ? k : 1024 - k;
      V p = Xr[k1] * Xr[k1] + Xi[k1] * Xi[k1];
      if (power[bin] != (V)-1 && fabs((double)(power[bin] - p)) > 1e-3 * fabs((double)p) + 1e-6) { printf("bin %d mismatch on duplicate\n", bin); exit(1); }
      power[bin] = p;
    }
  }
  for (int k = 0; k <= 512; ++k) if (power[k] == (V)-1) { printf("bin %d never written\n", k); exit(1); }
}

int main() {
  const int N = 1024;
  std::vector<double> x(N), w(N), ref(513);
  srand(99);
  double worst32 = 0, worst64 = 0;
  for (int trial = 0; trial < 4; ++trial) {
    for (int n = 0; n < N; ++n) {
      w[n] = 0.5 - 0.5 * cos(2 * M_PI * n / N);
      double noise = (rand() / (double)RAND_MAX - 0.5);
      x[n] = trial == 0 ? noise : trial == 1 ? sin(2 * M_PI * 1234.5 * n / 22050.0) : trial == 2 ? (n == 700) : 1.0 + 0.1 * noise;
    }
    double pmax = 0;
    for (int k = 0; k <= 512; ++k) {
      double re = 0, im = 0;
      for (int n = 0; n < N; ++n) { double a = -2 * M_PI * (double)((long)n * k % N) / N; re += x[n] * w[n] * cos(a); im += x[n] * w[n] * sin(a); }
      ref[k] = re * re + im * im; pmax = fmax(pmax, ref[k]);
    }
    std::vector<double> p64(513); frame_power_1024<double>(x.data(), w.data(), p64.data());
    std::vector<float> xf(x.begin(), x.end()), wf(w.begin(), w.end()), p32(513);
    frame_power_1024<float>(xf.data(), wf.data(), p32.data());
    for (int k = 0; k <= 512; ++k) {
      worst64 = fmax(worst64, fabs(p64[k] - ref[k]) / pmax);
      worst32 = fmax(worst32, fabs((double)p32[k] - ref[k]) / pmax);
    }
  }
  printf("worst |dP|/Pmax  fp64 %.3e  fp32 %.3e\n", worst64, worst32);
  if (worst64 > 1e-12 || worst32 > 3e-6) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
