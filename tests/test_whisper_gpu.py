"""GPU parity tests for the Whisper preset: CUDA path (through the torch custom op -> C ABI) vs the
CPU oracle, the committed golden vectors and, when importable, the live HF extractor.

Tolerance (BASELINE.md section 5 / north_star): max-abs <= 1e-4 on the normalised log-mel.
"""
import os

import numpy as np
import pytest
import torch

from audio_transformers_b200 import signals
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    from audio_transformers_b200 import ops as _ops
    return _ops


def _run(ops, clips, lengths=None):
    """clips: list of 1-D float32 arrays (ragged).  Returns (B, 80, 3000) numpy."""
    width = (max(len(c) for c in clips) + 3) // 4 * 4
    host = np.zeros((len(clips), max(width, 4)), dtype=np.float32)
    for i, c in enumerate(clips):
        host[i, :len(c)] = c
    wave = torch.from_numpy(host).cuda()
    lens = torch.tensor([len(c) for c in clips] if lengths is None else lengths, dtype=torch.int32).cuda()
    out = ops.whisper_logmel(wave, lens)
    torch.cuda.synchronize()
    assert out.shape == (len(clips), 80, 3000) and out.dtype == torch.float32 and out.is_cuda and out.is_contiguous()
    return out.cpu().numpy()


def test_golden_vectors(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "whisper_golden.npz"))
    names = [str(n) for n in g["names"]]
    report = {}
    for name in names:
        index, length = (int(v) for v in g[f"{name}/meta"])
        wav = signals.whisper_clip(index, seed=0, n_samples=length, kind=str(g[f"{name}/kind"]))
        out = _run(ops, [wav])[0]
        frames, values, stats = g[f"{name}/frames"], g[f"{name}/values"], g[f"{name}/stats"]
        err = float(np.abs(out[:, frames] - values).max())
        report[name] = err
        assert np.isfinite(out).all()
        assert err <= TOL, (name, err)
        assert abs(float(out.max()) - stats[1]) <= TOL and abs(float(out.min()) - stats[2]) <= TOL, name
    print("max-abs vs golden:", {k: f"{v:.2e}" for k, v in report.items()})


def test_signal_classes_vs_oracle(ops):
    """Per signal class, against the FP32 oracle (contract) and the FP64 restatement (truth)."""
    clips = [signals.whisper_clip(i, seed=11) for i in range(8)]           # 2 of each class
    out = _run(ops, clips)
    ref32 = O.whisper_logmel(clips, dtype=np.float32)
    ref64 = O.whisper_logmel(clips, dtype=np.float64)
    for i in range(8):
        e32 = float(np.abs(out[i] - ref32[i]).max())
        e64 = float(np.abs(out[i] - ref64[i]).max())
        print(f"class {signals.CLASSES[i % 4]:10s} |cuda-oracle32| {e32:.2e}  |cuda-oracle64| {e64:.2e}")
        assert e32 <= TOL and e64 <= TOL


def test_ragged_and_edge_lengths(ops):
    lengths = list(signals.EDGE_LENGTHS) + [int(v) for v in signals.ragged_lengths(5, seed=1)]
    clips = [signals.whisper_clip(100 + i, seed=2, n_samples=L) for i, L in enumerate(lengths)]
    out = _run(ops, clips)
    ref = O.whisper_logmel(clips, dtype=np.float32)
    assert np.abs(out - ref).max() <= TOL
    # clips longer than 30 s are truncated exactly like the extractor does
    long = clips[lengths.index(600000)]
    assert np.array_equal(out[lengths.index(600000)], _run(ops, [long[:480000]])[0])


def test_zeros_and_floor(ops):
    out = _run(ops, [np.zeros(480000, np.float32), np.zeros(7, np.float32)])
    assert np.abs(out + 1.5).max() < 1e-6                                      # (-10 + 4) / 4


def test_floor_reaches_exactly_the_blocks_below_it(ops):
    """The clip-floor pass skips every (tile, mel range) block whose smallest energy is already above the floor: clips
    whose dynamic range crosses the 80 dB line only in places -- a loud burst over very quiet noise, digital silence in
    the middle, a pure tone (far bins below the floor everywhere), quiet noise alone (nothing to floor) -- must still
    equal the oracle everywhere, and the floored set must be the oracle's."""
    rng = np.random.default_rng(77)
    n = 480000
    burst = (1e-6 * rng.standard_normal(n)).astype(np.float32)
    burst[16000:32000] += (0.5 * rng.standard_normal(16000)).astype(np.float32)
    gap = (0.1 * rng.standard_normal(n)).astype(np.float32)
    gap[200000:260000] = 0.0
    t = np.arange(n, dtype=np.float64) / 16000.0
    tone = (0.8 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    quiet = (1e-3 * rng.standard_normal(n)).astype(np.float32)
    ramp = (rng.standard_normal(n) * np.logspace(-7, 0, n)).astype(np.float32)
    clips = [burst, gap, tone, quiet, ramp, ramp[::-1].copy(), burst[:123457]]
    out = _run(ops, clips)
    ref = O.whisper_logmel(clips, dtype=np.float32)
    assert np.abs(out - ref).max() <= TOL
    for i in range(len(clips)):
        floor = ref[i].min()
        assert abs(float(out[i].min()) - float(floor)) <= TOL
        # nothing stays under the clip's floor, and what the oracle floored is floored here
        assert (out[i] >= out[i].max() - 2.0 - 1e-6).all()
        assert np.abs(out[i][ref[i] <= floor + 1e-7] - floor).max() <= TOL
    assert (ref[3] > ref[3].min() + 1e-3).mean() > 0.99                         # the "nothing to floor" clip is one


def test_lengths_none_means_full_stride(ops):
    clips = signals.whisper_batch(3, seed=5)
    wave = torch.from_numpy(clips).cuda()
    a = ops.whisper_logmel(wave, None).cpu().numpy()
    b = _run(ops, list(clips))
    assert np.array_equal(a, b)
    assert np.abs(a - O.whisper_logmel(list(clips))).max() <= TOL


def test_segment_chunks_like_inference(ops):
    """REF:whisper_finetune/inference.py:176-200: a 30 s clip cut into six 5 s chunks, each padded to 30 s."""
    clip = signals.whisper_clip(1, seed=9)
    chunks = [clip[i * 80000:(i + 1) * 80000] for i in range(6)]
    out = _run(ops, chunks)
    ref = O.whisper_logmel(chunks)
    assert np.abs(out - ref).max() <= TOL


def test_segment_features_zero_copy(ops):
    """The segment mode of the shim (one H2D, strided rows, one launch) equals the per-piece calls of the reference."""
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    fe = B200WhisperFeatureExtractor(device="cuda")
    for n in (480000, 333333, 80000, 79999, 5, 700001):
        clip = signals.whisper_clip(3, seed=12, n_samples=n)
        seg = fe.segment_features(clip, 80000, sampling_rate=16000)
        chunks = [clip[i:i + 80000] for i in range(0, n, 80000)]
        assert seg.shape == (len(chunks), 80, 3000) and seg.is_cuda
        assert np.array_equal(seg.cpu().numpy(), _run(ops, chunks))
        assert np.abs(seg.cpu().numpy() - O.whisper_logmel(chunks)).max() <= TOL
    assert fe.segment_features(np.zeros(0, np.float32), 80000).shape == (0, 80, 3000)
    with pytest.raises(ValueError):
        fe.segment_features(np.zeros(10, np.float32), 80000, sampling_rate=8000)


def test_collated_batch_path(ops):
    """__getitem__ keeps the waveform, the collate stacks the ragged clips into one pinned
    buffer, features_on_device does one H2D and one launch (replaces REF:whisper_finetune/dataset.py:53-110)."""
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    from audio_transformers_b200.collate import WaveformCollator, features_on_device
    lens = (480000, 91234, 16000, 520000)
    items = [{"waveform": signals.whisper_clip(i, seed=8, n_samples=n), "labels": torch.arange(3 + i),
              "emotion_label": i % 3} for i, n in enumerate(lens)]
    batch = WaveformCollator(pad_token_id=50257)(items)
    assert batch["waveform"].is_pinned() and batch["waveform"].shape == (4, 480000) and batch["labels"].shape == (4, 6)
    out = features_on_device(batch, B200WhisperFeatureExtractor(device="cuda"))
    feats = out["input_features"]
    assert feats.is_cuda and feats.shape == (4, 80, 3000)
    assert feats.to("cuda") is feats                                   # REF:whisper_finetune/train.py:188 becomes a no-op
    ref = O.whisper_logmel([it["waveform"] for it in items])
    assert np.abs(feats.cpu().numpy() - ref).max() <= TOL


def test_batch_independence_and_determinism(ops):
    clips = [signals.whisper_clip(i, seed=4) for i in range(5)]
    a = _run(ops, clips)
    b = _run(ops, clips[::-1])[::-1]
    assert np.array_equal(a, b)
    assert np.array_equal(a, _run(ops, clips))


def test_live_hf_extractor(ops):
    tr = pytest.importorskip("transformers")
    fe = tr.WhisperFeatureExtractor()
    clips = [signals.whisper_clip(i, seed=21, n_samples=n) for i, n in enumerate((480000, 123457, 480000, 32000))]
    ref = fe([c.astype(np.float64) for c in clips], sampling_rate=16000, return_tensors="pt").input_features.numpy()
    out = _run(ops, clips)
    err = np.abs(out - ref).reshape(len(clips), -1).max(axis=1)
    print("max-abs vs live HF per clip:", err)
    assert err.max() <= TOL


_STRIDES = [480000, 480004, 500000, 20600, 20604, 31000, 284, 280]


def _stride_case(ops, stride):
    rng = np.random.default_rng(stride)
    n = min(stride, 480000)
    host = (0.1 * rng.standard_normal((3, stride))).astype(np.float32)
    lens = [n, max(1, n - 777), max(1, n // 2)]
    out = ops.whisper_logmel(torch.from_numpy(host).cuda(), torch.tensor(lens, dtype=torch.int32).cuda()).cpu().numpy()
    return host, lens, out


def test_tma_and_generic_staging_agree(ops, tmp_path):
    """Interior tiles arrive through one TMA box per tile, edge tiles through ordinary stores into the same
    layout.  Both must give the same bits for every row stride the ABI accepts (the tensor map's extents depend
    on it; 20600/20604 straddle the first stride that has an interior tile at all, 284/280 the smallest map).
    The all-stores variant is a property of the handle (B200MEL_DEBUG_NO_TMA=1 when it is created, a test hook), so
    it runs in a child process."""
    import subprocess
    import sys
    script = tmp_path / "no_tma.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        "from audio_transformers_b200 import ops\n"
        "import test_whisper_gpu as T\n"
        "np.savez(sys.argv[1], **{str(s): T._stride_case(ops, s)[2] for s in T._STRIDES})\n")
    env = dict(os.environ, B200MEL_DEBUG_NO_TMA="1")
    dump = str(tmp_path / "no_tma.npz")
    subprocess.run([sys.executable, str(script), dump], check=True, env=env, timeout=600)
    other = np.load(dump)
    for stride in _STRIDES:
        host, lens, a = _stride_case(ops, stride)
        assert np.array_equal(a, other[str(stride)]), stride
        ref = O.whisper_logmel([host[i, :lens[i]] for i in range(3)])
        assert np.abs(a - ref).max() <= TOL, stride


def test_cuda_graph_capture_and_replay(ops):
    """include/b200mel.h promises stream-ordered, allocation-free, sync-free calls: a call (TMA descriptor as a kernel
    parameter, programmatic launch attribute and all) can be captured once and replayed on new audio."""
    clips = signals.whisper_batch(4, seed=17)
    static_in = torch.from_numpy(clips).cuda()
    ops.whisper_logmel(static_in, None)                      # warm-up outside capture (handle creation)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_out = ops.whisper_logmel(static_in, None)
    other = signals.whisper_batch(4, seed=18)
    static_in.copy_(torch.from_numpy(other))
    g.replay()
    torch.cuda.synchronize()
    assert np.abs(static_out.cpu().numpy() - O.whisper_logmel(list(other))).max() <= TOL
    static_in.copy_(torch.from_numpy(clips))
    g.replay()
    torch.cuda.synchronize()
    assert np.abs(static_out.cpu().numpy() - O.whisper_logmel(list(clips))).max() <= TOL


def test_large_ragged_batch(ops):
    """Many more tiles than resident CTAs, random lengths (tiles of pure padding are skipped, edge tiles use the
    store path, interior tiles TMA): spot-check clips against the oracle."""
    B = 400
    rng = np.random.default_rng(5)
    base = signals.whisper_batch(12, seed=9)
    idx = rng.integers(0, 12, size=B)
    lens = rng.integers(1, 480001, size=B).astype(np.int32)
    lens[::7] = 480000
    wave = torch.from_numpy(base).cuda()[torch.from_numpy(idx).cuda()].contiguous()
    out = ops.whisper_logmel(wave, torch.from_numpy(lens).cuda())
    assert bool(torch.isfinite(out).all())
    for b in (0, 1, 7, 233, B - 1, int(np.argmin(lens)), int(np.argmax(lens))):
        ref = O.whisper_logmel([base[idx[b]][:lens[b]]])[0]
        assert np.abs(out[b].cpu().numpy() - ref).max() <= TOL, b


def test_frame_mask(ops):
    lens = torch.tensor([1, 160, 161, 480000, 600000], dtype=torch.int32).cuda()
    m = ops.whisper_frame_mask(lens).cpu().numpy()
    assert np.array_equal(m, O.whisper_attention_mask([1, 160, 161, 480000, 600000]))


def test_unaligned_inputs_are_repacked(ops):
    clip = signals.whisper_clip(2, seed=6, n_samples=100003)
    wave = torch.from_numpy(clip[None, :]).cuda()                              # T % 4 != 0
    out = ops.whisper_logmel(wave, None).cpu().numpy()
    assert np.abs(out - O.whisper_logmel([clip])).max() <= TOL


def test_cpu_tensor_fails_loudly(ops):
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.whisper_logmel(torch.zeros(1, 480000), None)


def test_column_sliced_view_without_lengths(ops):
    """A row-sliced view of a wider buffer (aligned pointer, stride(0) != T) and lengths=None: the clips are the
    T visible samples of each row, not the whole row stride."""
    big = torch.from_numpy(signals.whisper_batch(3, seed=23)).cuda()            # (3, 480000)
    view = big[:, :160000]
    assert view.stride(0) == 480000 and view.data_ptr() % 16 == 0
    out = ops.whisper_logmel(view, None).cpu().numpy()
    ref = O.whisper_logmel([c[:160000] for c in big.cpu().numpy()])
    assert np.abs(out - ref).max() <= TOL
    # lengths longer than the view are clamped to it; lengths on another device type are rejected
    lens = torch.tensor([999999, 160000, 5], dtype=torch.int32).cuda()
    out2 = ops.whisper_logmel(view, lens).cpu().numpy()
    assert np.array_equal(out2[:2], out[:2])
    with pytest.raises((ValueError, RuntimeError)):
        ops.whisper_logmel(view, torch.tensor([1, 2, 3], dtype=torch.int32))


# ---- the reference's call shapes, end to end through the shims, against the live HF extractor ---------------------
def _hf():
    tr = pytest.importorskip("transformers")
    return tr.WhisperFeatureExtractor()


def test_input_canonicalisation_vs_live_hf():
    """HF:models/whisper/feature_extraction_whisper.py:274-292: float64 arrays (what `datasets` yields and the reference
    passes, REF:whisper_finetune/dataset.py:57-62), lists of them, a 2-D array, a list of Python floats, a mixed
    float32 / float64 list, a tuple -- all through the native cast + H2D + kernel route."""
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    hf, fe = _hf(), B200WhisperFeatureExtractor(device="cuda")
    c = [signals.whisper_clip(i, seed=41, n_samples=n) for i, n in enumerate((480000, 77777, 160000, 480000))]
    cases = {
        "one float64 array": c[1].astype(np.float64),
        "one float32 array": c[0],
        "list of float64 arrays": [x.astype(np.float64) for x in c],
        "mixed float32 / float64 list": [c[0].astype(np.float64), c[1], c[2].astype(np.float64), c[3]],
        "tuple of arrays": tuple(x.astype(np.float64) for x in c[:2]),
        "2-D float64 array": np.stack([c[0], c[3]]).astype(np.float64),
        "2-D float32 array": np.stack([c[0], c[3]]),
        "list of Python floats": [float(v) for v in c[1][:4000]],
        "int16-valued integer array": (c[2][:30000] * 1000).astype(np.int32),
    }
    for name, raw in cases.items():
        ref = hf(raw, sampling_rate=16000, return_tensors="pt").input_features.numpy()
        got = fe(raw, sampling_rate=16000, return_tensors="pt").input_features
        assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == ref.shape, name
        err = float(np.abs(got.cpu().numpy() - ref).max())
        print(f"{name:32s} max-abs vs live HF {err:.2e}")
        assert err <= TOL, (name, err)
    # return_tensors=None / "np": numpy on the host, like HF
    got = fe(c[1].astype(np.float64), sampling_rate=16000)
    assert isinstance(got["input_features"], np.ndarray) and got["input_features"].shape == (1, 80, 3000)
    # repeated calls reuse the two staging slots: the third call must not disturb the first result
    a = fe(c[0].astype(np.float64), sampling_rate=16000, return_tensors="pt").input_features
    fe(c[1].astype(np.float64), sampling_rate=16000, return_tensors="pt")
    fe(c[2].astype(np.float64), sampling_rate=16000, return_tensors="pt")
    ref0 = hf(c[0].astype(np.float64), sampling_rate=16000, return_tensors="pt").input_features.numpy()
    assert np.abs(a.cpu().numpy() - ref0).max() <= TOL


def test_stereo_and_wrong_rate_are_rejected_like_hf():
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    hf, fe = _hf(), B200WhisperFeatureExtractor(device="cuda")
    stereo = np.zeros((2, 2, 100), np.float32)
    with pytest.raises(ValueError, match="Only mono-channel audio"):
        hf(stereo, sampling_rate=16000)
    with pytest.raises(ValueError, match="Only mono-channel audio"):
        fe(stereo, sampling_rate=16000)
    for bad in (hf, fe):
        with pytest.raises(ValueError, match="sampling rate of 16000"):
            bad(np.zeros(100, np.float32), sampling_rate=8000)
    with pytest.raises(NotImplementedError):
        fe(np.zeros(100, np.float32), sampling_rate=16000, padding=True)        # LONGEST in HF: not what this computes
    with pytest.raises(NotImplementedError):
        fe(np.zeros(100, np.float32), sampling_rate=16000, padding="longest")


def test_attention_mask_vs_live_hf():
    """HF:models/whisper/feature_extraction_whisper.py:328-337 (mask[:, ::160]) for the edge lengths."""
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    hf, fe = _hf(), B200WhisperFeatureExtractor(device="cuda")
    lens = (1, 159, 160, 161, 480000, 600000)
    clips = [signals.whisper_clip(i, seed=43, n_samples=n).astype(np.float64) for i, n in enumerate(lens)]
    ref = hf(clips, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
    got = fe(clips, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
    assert tuple(got["attention_mask"].shape) == tuple(ref["attention_mask"].shape) == (len(lens), 3000)
    assert np.array_equal(got["attention_mask"].cpu().numpy(), ref["attention_mask"].numpy())
    assert np.abs(got["input_features"].cpu().numpy() - ref["input_features"].numpy()).max() <= TOL
    # one clip per call, as the reference calls it
    for c in clips[:4]:
        r = hf(c, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
        g = fe(c, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
        assert np.array_equal(g["attention_mask"].cpu().numpy(), r["attention_mask"].numpy())


def test_processor_audio_path_vs_live_hf():
    """HF:models/whisper/processing_whisper.py:31-54 with audio: positional, audio=, audio + text (labels), and the
    reference's exact line REF:whisper_finetune/dataset.py:58-62."""
    from audio_transformers_b200 import B200WhisperProcessor

    class Tok:                                        # the tokenizer is passthrough; a stub keeps the test offline
        pad_token_id, eos_token_id = 50257, 50256

        def __call__(self, text=None, **kw):
            return {"input_ids": [[7, 8, 9]]}

    hf = _hf()
    proc = B200WhisperProcessor(tokenizer=Tok(), device="cuda")
    audio = signals.whisper_clip(5, seed=47, n_samples=200000).astype(np.float64)
    ref = hf(audio, sampling_rate=16000, return_tensors="pt").input_features
    a = proc(audio, sampling_rate=16000, return_tensors="pt").input_features.squeeze(0)      # dataset.py:58-62
    assert a.is_cuda and tuple(a.shape) == (80, 3000)
    assert float((a.cpu() - ref[0]).abs().max()) <= TOL
    b = proc(audio=audio, sampling_rate=16000, return_tensors="pt")
    assert torch.equal(b["input_features"].squeeze(0), a) and torch.equal(b.input_features.squeeze(0), a)
    both = proc(audio=audio, text="hello", sampling_rate=16000, return_tensors="pt")
    assert both["labels"] == [[7, 8, 9]] and torch.equal(both["input_features"].squeeze(0), a)
    assert torch.equal(proc.feature_extractor(audio, sampling_rate=16000, return_tensors="pt").input_features.squeeze(0), a)
    assert a.to("cuda") is a                          # REF:whisper_finetune/inference.py:154 `.to(device)` is a no-op


def test_two_streams_interleaved_match_serial(ops):
    """The handle is immutable and the workspace is per (stream, batch): calls enqueued alternately on two streams, with
    different batches in flight at once, give the bits of the same calls run one after the other."""
    waves = [torch.from_numpy(signals.whisper_batch(n, seed=40 + i)).cuda() for i, n in enumerate((5, 9, 5, 12, 9, 3))]
    lens = [torch.tensor([int(v) for v in signals.ragged_lengths(w.shape[0], seed=50 + i)], dtype=torch.int32).cuda()
            for i, w in enumerate(waves)]
    serial = [ops.whisper_logmel(w, l).clone() for w, l in zip(waves, lens)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [None] * len(waves)
    for rep in range(3):
        for i, (w, l) in enumerate(zip(waves, lens)):
            with torch.cuda.stream(streams[i % 2]):
                outs[i] = ops.whisper_logmel(w, l)
    torch.cuda.synchronize()
    for a, b in zip(serial, outs):
        assert torch.equal(a, b)
