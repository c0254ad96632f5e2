"""CPU checks of the index algebra the CUDA kernels rely on (pure numpy restatements, no GPU):

* Whisper frame, N = 400 = 16 x 25 (Good-Thomas, no twiddles): csrc/fft_codelets.cuh pfa400_n / pfa400_k
* urban frame, N = 1024 = 32 x 32 (Cooley-Tukey) with the decimation-in-frequency split of pass 2 into two 16-point
  DFTs per k2 and the twiddle table T_h[k2][r] of csrc/urban_packed.cuh (u2_build_image), incl. the rule which of the
  32 outputs of a task are stored where, and the one-load form T[r + 16] = T[r] * C.
"""
import numpy as np


def _w(e, n):
    return np.exp(-2j * np.pi * e / n)


def test_whisper_prime_factor_map_reaches_every_bin_once():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(400)
    ref = np.fft.fft(x)
    # pass 1: per class a, 25-point DFT over b of x[(25 a + 16 b) mod 400]; pass 2: per k2, 16-point DFT over a
    y = np.array([np.fft.fft(x[(25 * a + 16 * np.arange(25)) % 400]) for a in range(16)])       # [a][k2]
    seen = np.zeros(201, int)
    for k2 in range(13):
        X = np.fft.fft(y[:, k2])                                                                  # [k1]
        for k1 in range(16 if k2 else 9):
            k = (225 * k1 + 176 * k2) % 400
            assert abs(X[k1] - ref[k]) < 1e-9 * np.abs(ref).max()
            seen[k if k <= 200 else 400 - k] += 1
    assert (seen == 1).all()


def _urban_T(k2, r, h):
    e = (r * k2 + 32 * (r & 15) * h + (512 * h if r >= 16 else 0)) & 1023
    return _w(e, 1024)


def test_urban_dif_split_and_bin_map():
    rng = np.random.default_rng(1)
    xw = rng.standard_normal(1024)
    ref = np.abs(np.fft.rfft(xw)) ** 2
    Y = np.array([np.fft.fft(xw[r::32]) for r in range(32)])                                     # pass 1: [r][k2]
    P = np.full(513, np.nan)
    writes = np.zeros(513, int)
    for k2 in range(17):
        for h in range(2):
            u = np.array([Y[r, k2] * _urban_T(k2, r, h) + Y[r + 16, k2] * _urban_T(k2, r + 16, h) for r in range(16)])
            # the kernel's one-load form
            C = _urban_T(k2, 16, h)
            u2 = np.array([_urban_T(k2, r, h) * (Y[r, k2] + C * Y[r + 16, k2]) for r in range(16)])
            assert np.abs(u - u2).max() < 1e-9 * np.abs(u).max()
            X = np.fft.fft(u)                                                                     # [m], k1 = 2 m + h
            for m in range(16):
                if k2 == 0 and not (m < 8 or (m == 8 and h == 0)):
                    continue                                                                      # mirrors of stored bins
                if k2 == 16 and m >= 8:
                    continue
                b = k2 + 32 * h + 64 * m if m < 8 else 1024 - 64 * m - 32 * h - k2
                P[b] = abs(X[m]) ** 2
                writes[b] += 1
    assert (writes == 1).all()
    assert np.abs(P - ref).max() < 1e-9 * ref.max()
    # k2 = 0 and 16 have real pass-1 outputs (their imaginary rows are not stored)
    assert np.abs(Y[:, 0].imag).max() < 1e-9 and np.abs(Y[:, 16].imag).max() < 1e-9


def test_urban_pass1_lane_layout_is_conflict_free():
    """E pitch 1026 floats per r: the 64-bit stores of pass 1 (lane = r) hit 16 different bank pairs per half-warp."""
    for row in range(32):
        for p in range(16):
            for half in range(2):
                pairs = {((r * 1026 + row * 32 + 2 * p) // 2) % 16 for r in range(16 * half, 16 * half + 16)}
                assert len(pairs) == 16
