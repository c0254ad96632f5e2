"""GPU parity tests for the urban preset (MelSpectrogram(22050, 1024, hop 512, 64 mels) + log(.+1e-9)).

Tolerance: max-abs <= 1e-4 on log(mel + 1e-9); the linear mel is checked
relative to the clip's peak mel value.
"""
import os

import numpy as np
import pytest
import torch

from audio_transformers_b200 import signals
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from audio_transformers_b200 import ops as _ops
    return _ops


def test_golden_vectors(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "urban_golden.npz"))
    wave = signals.urban_batch(int(g["batch"]), seed=int(g["seed"]))                  # (4, 1, 88200)
    x = torch.from_numpy(wave[:, 0]).cuda()
    logmel = ops.mel_power(x, 1e-9).cpu().numpy()
    mel = ops.mel_power(x, -1.0).cpu().numpy()
    assert logmel.shape == (4, 64, 173) and logmel.dtype == np.float32
    err = float(np.abs(logmel - g["logmel"][:, 0]).max())
    print(f"urban max-abs vs golden log-mel: {err:.2e}")
    assert err <= TOL
    assert np.abs(mel - g["mel"][:, 0]).max() <= 1e-5 * np.abs(g["mel"]).max()
    z = ops.mel_power(torch.zeros(1, 88200, device="cuda"), 1e-9).cpu().numpy()
    assert np.abs(z - g["zeros_logmel"][:, 0]).max() < 1e-5


def test_config1_batch32_vs_oracle(ops):
    """randn(32, 1, 88200) peak-normalised."""
    wave = signals.urban_batch(32, seed=0)
    out = ops.mel_power(torch.from_numpy(wave[:, 0]).cuda(), 1e-9).cpu().numpy()
    ref32 = O.urban_melspec(wave, dtype=np.float32)[:, 0]
    ref64 = O.urban_melspec(wave, dtype=np.float64)[:, 0]
    e32, e64 = float(np.abs(out - ref32).max()), float(np.abs(out - ref64).max())
    print(f"urban batch 32: |cuda-oracle32| {e32:.2e}  |cuda-oracle64| {e64:.2e}")
    assert e32 <= TOL and e64 <= TOL


@pytest.mark.parametrize("n", [88200, 513, 1000, 1024, 4096, 44100, 100003, 16384 + 512 * 31])
def test_lengths(ops, n):
    rng = np.random.default_rng(n)
    wave = (0.3 * rng.standard_normal((3, n))).astype(np.float32)
    wave[1] = (0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / 22050.0)).astype(np.float32)
    x = torch.from_numpy(wave).cuda()
    out = ops.mel_power(x, 1e-9).cpu().numpy()
    ref = O.urban_melspec(wave, dtype=np.float32)
    assert out.shape == ref.shape == (3, 64, 1 + n // 512)
    # broadband clips: the contract tolerance on log(mel + 1e-9)
    assert np.abs(out[[0, 2]] - ref[[0, 2]]).max() <= TOL
    # pure tone: far-off mels sit at the FP32 round-off floor (~1e-14 of the peak power), below the 1e-9
    # epsilon, where any two FP32 FFTs (including torch's own on different hardware) disagree in the log.
    # Check the linear mel against the FP64 restatement relative to the clip's peak, and the log where the
    # mel is comfortably above that floor.
    lin = ops.mel_power(x, -1.0).cpu().numpy()[1]
    lin64 = O.urban_melspec(wave[1:2], log_eps=None, dtype=np.float64)[0]
    assert np.abs(lin - lin64).max() <= 1e-6 * lin64.max()
    strong = lin64 > 1e-4 * lin64.max()
    assert np.abs(out[1][strong] - np.log(lin64[strong] + 1e-9)).max() <= TOL


def test_pure_tone_weak_bins_against_live_torchaudio(ops):
    """The carve-out of test_lengths with numbers next to it: on the bins it excludes (mel below 1e-4 of the clip's peak
    for a pure tone) the log-mel of the CUDA path, of live torchaudio on this host's CPU and of the FP64 restatement are
    compared pairwise.  torchaudio itself misses FP64 there by more than the contract tolerance; the CUDA path must be
    no further from FP64 than a small multiple of that."""
    ta = pytest.importorskip("torchaudio")
    n = 88200
    wave = (0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / 22050.0)).astype(np.float32)[None]
    ours = ops.mel_power(torch.from_numpy(wave).cuda(), 1e-9).cpu().numpy()[0]
    lib = torch.log(ta.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)(torch.from_numpy(wave)) + 1e-9).numpy()[0]
    lin64 = O.urban_melspec(wave, log_eps=None, dtype=np.float64)[0]
    truth = np.log(lin64 + 1e-9)
    weak = lin64 <= 1e-4 * lin64.max()
    assert weak.any() and (~weak).any()
    e_ours, e_lib, e_pair = (float(np.abs(a[weak] - b[weak]).max()) for a, b in ((ours, truth), (lib, truth), (ours, lib)))
    print(f"pure tone, {int(weak.sum())} weak bins of {weak.size}: |cuda - fp64| {e_ours:.2e}  |torchaudio cpu - fp64| {e_lib:.2e}  "
          f"|cuda - torchaudio cpu| {e_pair:.2e};  strong bins: |cuda - torchaudio cpu| {float(np.abs(ours[~weak] - lib[~weak]).max()):.2e}")
    assert np.abs(ours[~weak] - lib[~weak]).max() <= TOL
    assert e_ours <= max(4.0 * e_lib, TOL)


def test_module_matches_live_torchaudio(ops):
    ta = pytest.importorskip("torchaudio")
    from audio_transformers_b200 import B200MelSpectrogram
    ref_tf = ta.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
    mine = B200MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
    assert set(mine.state_dict().keys()) == set(ref_tf.state_dict().keys()) == {"spectrogram.window", "mel_scale.fb"}
    for k, v in ref_tf.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k
    wave = torch.from_numpy(signals.urban_batch(6, seed=3))                              # (6, 1, 88200)
    ref = torch.log(ref_tf(wave) + 1e-9)
    out = torch.log(mine(wave.cuda()) + 1e-9)                                           # REF:urban_sounds/dataset.py:55-56
    assert out.shape == ref.shape == (6, 1, 64, 173) and out.is_cuda
    assert (out.cpu() - ref).abs().max().item() <= TOL
    fused = B200MelSpectrogram(log_eps=1e-9)(wave.cuda())
    assert (fused.cpu() - ref).abs().max().item() <= TOL
    # a single (1, T) clip, as REF:urban_sounds/dataset.py:55 passes it
    one = mine(wave[0].cuda())
    assert one.shape == (1, 64, 173)
    assert torch.equal(one, mine(wave.cuda())[0])


def test_prep_front_end_vs_golden_and_oracle(ops, golden_dir):
    """REF:urban_sounds/dataset.py:26-58 (process_audio) on the GPU, batched over mixed rates and channel counts,
    against the torch/torchaudio golden vectors and the oracle."""
    from audio_transformers_b200.urban import B200UrbanFrontEnd
    g = np.load(os.path.join(golden_dir, "urban_prep_golden.npz"))
    names = [str(n) for n in g["names"]]
    metas = [tuple(int(v) for v in g[f"{n}/meta"]) for n in names]
    arrays = [signals.urban_raw_clip(n, *m) for n, m in zip(names, metas)]
    rates = [m[0] for m in metas]
    fe = B200UrbanFrontEnd()
    waves = fe.waveforms(arrays, rates)
    assert waves.shape == (len(names), 88200) and waves.is_cuda and waves.dtype == torch.float32
    feats = fe.process_batch(arrays, rates)
    assert feats.shape == (len(names), 1, 64, 173) and feats.is_cuda
    w, f = waves.cpu().numpy(), feats.cpu().numpy()
    for i, name in enumerate(names):
        assert np.abs(w[i, ::37] - g[f"{name}/wave_sub"]).max() <= 2e-6, name
        assert abs(float(np.abs(w[i]).max()) - g[f"{name}/wave_stats"][1]) <= 1e-6, name
        ref_w = O.urban_preprocess(arrays[i], orig_sr=rates[i])
        assert np.abs(w[i] - ref_w[0]).max() <= 2e-6, name
        ref = g[f"{name}/logmel_sub"]
        lm = f[i, 0][:, ::9]
        strong = (ref > ref.max() - 9.2) & (ref > np.log(1e-6))      # see tests/test_oracle.py::_strong
        assert np.abs(lm[strong] - ref[strong]).max(initial=0.0) <= TOL, name
        assert np.abs(np.exp(lm) - np.exp(ref)).max() <= 1e-5 * np.exp(ref).max() + 1e-9, name
    # the reference's per-sample signature
    one = fe.process_audio(arrays[0], rates[0])
    assert one.shape == (1, 64, 173) and torch.equal(one, feats[0])


def test_errors(ops):
    from audio_transformers_b200 import B200MelSpectrogram
    with pytest.raises(NotImplementedError):
        B200MelSpectrogram(n_mels=128)
    with pytest.raises(RuntimeError):
        B200MelSpectrogram()(torch.zeros(1, 88200))
    with pytest.raises(ValueError):
        ops.mel_power(torch.zeros(1, 512, device="cuda"), 1e-9)


@pytest.mark.parametrize("batch,n", [(37, 2048 + 17), (150, 1537), (5, 513), (64, 88200)])
def test_flat_frame_list_straddles_clips(ops, batch, n):
    """The packed kernel walks one flat list of frames: 32-frame tiles and frame pairs straddle clip boundaries at
    every alignment (odd frame counts), clips shorter than a tile, many tiles per CTA."""
    rng = np.random.default_rng(batch * 100003 + n)
    wave = (0.3 * rng.standard_normal((batch, n))).astype(np.float32)
    out = ops.mel_power(torch.from_numpy(wave).cuda(), 1e-9).cpu().numpy()
    pick = sorted({0, 1, batch // 2, batch - 2, batch - 1})
    ref = O.urban_melspec(wave[pick], dtype=np.float32)
    assert out.shape == (batch, 64, 1 + n // 512)
    assert np.abs(out[pick] - ref).max() <= TOL


def test_large_batch_equals_single_clip_calls(ops):
    """Size-independent property at a batch far above one tile per SM (600 clips = 3244 tiles on 148 CTAs): every clip's
    features are bit-identical to the same clip processed alone, whatever tile / lane pair / CTA it lands in."""
    g = torch.Generator().manual_seed(7)
    wave = torch.randn(600, 88200, generator=g).cuda()
    out = ops.mel_power(wave, 1e-9)
    assert torch.isfinite(out).all()
    for c in (0, 1, 172, 311, 598, 599):
        assert torch.equal(out[c], ops.mel_power(wave[c:c + 1].contiguous(), 1e-9)[0]), c
