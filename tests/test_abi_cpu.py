"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, its lookup tables equal the oracle's, and the host layer fails loudly without a GPU."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT
from oracle import denoise, features


def test_package_imports_without_gpu(pkg):
    assert pkg.FEATURE_LEN == 149
    alias = importlib.import_module("dysb200")
    assert alias is pkg


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "dysfluency_b200.h")).read()
    declared = set(re.findall(r"DYS_API\s+[\w\s\*]+?\b(dys_\w+)\s*\(", header))
    assert len(declared) >= 13
    assert declared == set(pkg._lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = pkg._lib.load()
    assert loaded.dys_version() == 200
    assert loaded.dys_last_error() == b""


def test_workspace_queries(pkg):
    lib = pkg._lib.load()
    need = lib.dys_workspace_bytes(64, 48000, 1)
    least = lib.dys_workspace_min_bytes(64, 48000, 1)
    assert 0 < least <= need
    assert lib.dys_workspace_bytes(64, 48000, 0) < need
    assert lib.dys_workspace_bytes(-1, 10, 0) == -1


def test_tables_match_oracle(pkg):
    fe = pkg.frontend
    assert np.array_equal(fe.get_table(0), features.mel_filterbank())
    np.testing.assert_allclose(fe.get_table(1), features.dct_matrix(), atol=1e-8)
    for ti in (0, 13, 50, 99):
        np.testing.assert_array_equal(fe.get_table(2, ti), features.chroma_filterbank(float(features.TUNING_EDGES[ti])).T)
    assert np.array_equal(fe.get_table(3), features.hann_periodic(2048).astype(np.float32))
    assert np.array_equal(fe.get_table(4), features.TUNING_EDGES)
    taps = fe.get_table(5)
    np.testing.assert_allclose(np.outer(taps[:33], taps[33:]), denoise.smoothing_filter(), atol=1e-17)
    assert np.array_equal(fe.get_table(6), denoise._window_sumsquare(20)[1024:1280])
    assert fe.get_table(7)[0] == denoise.iir_coefficient()


def test_mel_runs_are_conflict_free_and_exact(pkg):
    """The lane-per-filter layout of the sparse mel bank: every group's 32 run starts differ mod 32 (no shared-memory bank
    conflicts on the power reads), the leading padding is zero weights only, no step was added, and the runs reproduce
    librosa's filterbank exactly."""
    fe = pkg.frontend
    runs, wt, dense = fe.get_table(8), fe.get_table(9), fe.get_table(0)
    start, length, goff = runs[:128], runs[128:256], runs[256:]
    rebuilt = np.zeros_like(dense)
    steps = 0
    for g in range(4):
        s, l = start[32 * g:32 * g + 32], length[32 * g:32 * g + 32]
        assert len(set(int(v) % 32 for v in s)) == 32
        assert s.min() >= 0 and (s + l).max() <= 1025
        steps += int(l.max())
        for lane in range(32):
            for j in range(int(l[lane])):
                rebuilt[32 * g + lane, s[lane] + j] = wt[goff[g] + 32 * j + lane]
    assert np.array_equal(rebuilt, dense)
    assert steps == 86


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    fe = pkg.frontend
    y = np.zeros(48000, np.float32)
    with pytest.raises(pkg.DysError):
        fe.extract_features_batch([y])
    with pytest.raises(pkg.DysError):
        fe.extract_features(y, 16000)
    # the C ABI itself refuses too
    assert pkg._lib.load().dys_init() != 0
    assert b"CUDA" in pkg._lib.load().dys_last_error() or b"device" in pkg._lib.load().dys_last_error()


def test_text_features_and_errors(pkg):
    fe = pkg.frontend
    assert not fe.extract_text_features("").any()
    v = fe.extract_text_features("the the cat sat sat sat")
    assert v.dtype == np.float32 and v.tolist()[:3] == [23.0, 6.0, 3.0]
    assert fe.load_audio("/nonexistent/file.wav") == (None, None)
    with pytest.raises(ValueError):
        fe.extract_features_batch([np.zeros(10, np.float32)], sr=22050)


def test_sharding_and_windows(pkg):
    sh = pkg.sharding
    for n, w in ((10000, 8), (7, 3), (5, 8)):
        ranges = [sh.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert len(sh.sliding_windows(57_600_000)) == 2399          # BASELINE config 4
    assert sh.sliding_windows(47999) == []


def test_rank_to_gpu_spreading(pkg):
    f = pkg.sharding.spread_device_index
    assert [f(r, 4, 8) for r in range(4)] == [0, 4, 1, 5] and [f(r, 2, 8) for r in range(2)] == [0, 4]
    assert [f(r, 8, 8) for r in range(8)] == list(range(8)) and f(0, 1, 8) == 0 and [f(r, 2, 3) for r in range(2)] == [0, 1]
    assert sorted(f(r, 6, 8) for r in range(6)) == [0, 1, 2, 4, 5, 6]


def test_wav_io_round_trip(pkg, tmp_path):
    q = (np.random.default_rng(0).integers(-32768, 32767, 1000)).astype(np.int16)
    p = tmp_path / "a.wav"
    pkg.wavio.write_wav_pcm16(str(p), q)
    y, sr = pkg.wavio.read_wav(str(p))
    assert sr == 16000 and np.array_equal(y, q.astype(np.float32) / 32768)
    from oracle import wavio as owav
    back, _ = owav.read_wav_pcm16(str(p))
    assert np.array_equal(back, q)


def test_on_disk_formats_are_byte_identical_to_the_reference(pkg, golden_dir, tmp_path):
    """the WAV the package writes for the reference's PCM and the .npy it writes for the reference's
    vector are byte-for-byte the files the reference committed (soundfile PCM_16 header, np.save v1 header)."""
    import io
    g = np.load(os.path.join(golden_dir, "ref_file_bytes.npz"))
    wav_bytes, npy_bytes = g["wav"].tobytes(), g["npy"].tobytes()
    src = tmp_path / "ref.wav"
    src.write_bytes(wav_bytes)
    y, sr = pkg.wavio.read_wav(str(src))
    assert sr == 16000 and y.dtype == np.float32
    pcm = np.round(y * 32768.0).astype(np.int16)
    out = tmp_path / "ours.wav"
    pkg.wavio.write_wav_pcm16(str(out), pcm, sr)
    assert out.read_bytes() == wav_bytes
    vec = np.load(io.BytesIO(npy_bytes))
    assert vec.shape == (149,) and vec.dtype == np.float32
    np.save(tmp_path / "ours.npy", vec)
    assert (tmp_path / "ours.npy").read_bytes() == npy_bytes and len(npy_bytes) == 724


def test_torch_operators_are_registered_and_have_no_cpu_kernel(pkg):
    """The PyTorch-op face of the boundary: torch.ops.dysb200.* exist, infer shapes on fake tensors, and refuse CPU
    tensors instead of falling back."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    pkg.torch_ops                                                   # registers the operators
    for name in ("features_raw", "features_raw_clean", "qc_metrics", "resample_to_16k"):
        assert hasattr(torch.ops.dysb200, name)
    with FakeTensorMode():
        a = torch.empty(1000, device="cuda")
        s = torch.empty(4, dtype=torch.int64, device="cuda")
        ln = torch.empty(4, dtype=torch.int32, device="cuda")
        raw, clean, st = torch.ops.dysb200.features_raw_clean(a, s, ln, 250, 1.0)
        assert raw.shape == clean.shape == (4, 149) and st.shape == (8,) and st.dtype == torch.int32
        assert torch.ops.dysb200.qc_metrics(a, s, ln, 250).shape == (4, 3)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.dysb200.features_raw(torch.zeros(10), torch.zeros(1, dtype=torch.int64), torch.ones(1, dtype=torch.int32), 10)


def test_numa_binding_helper_never_raises(pkg):
    """bench.py calls it on every rank before allocating pinned staging; without a GPU / NUMA information it must
    report why it did nothing instead of failing."""
    before = os.sched_getaffinity(0)
    r = pkg.sharding.bind_to_gpu_numa_node(0)
    assert isinstance(r, dict) and ("skipped" in r or "node" in r)
    if "skipped" in r:
        assert os.sched_getaffinity(0) == before


def test_wav_reader_covers_libsndfile_encodings_and_mono_mixdown(pkg, tmp_path):
    """load_audio's WAV side (pipeline1.py:102: soundfile -> float32, librosa mono=True = mean over channels)."""
    import struct
    rng = np.random.default_rng(0)

    def write(path, codec, bits, channels, payload, sr=16000):
        blk = channels * bits // 8
        hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, codec, channels, sr, sr * blk, blk, bits)
        open(path, "wb").write(hdr + b"data" + struct.pack("<I", len(payload)) + payload)

    q = rng.integers(-32768, 32767, size=(500, 2)).astype("<i2")
    write(tmp_path / "st16.wav", 1, 16, 2, q.tobytes())
    y, sr = pkg.wavio.read_wav(str(tmp_path / "st16.wav"))
    assert sr == 16000 and y.dtype == np.float32
    np.testing.assert_array_equal(y, (q.astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32))
    with pytest.raises(ValueError):
        pkg.wavio.read_wav_pcm16(str(tmp_path / "st16.wav"))
    f = rng.standard_normal(300).astype("<f4")
    write(tmp_path / "f32.wav", 3, 32, 1, f.tobytes(), sr=22050)
    y, sr = pkg.wavio.read_wav(str(tmp_path / "f32.wav"))
    assert sr == 22050 and np.array_equal(y, f)
    q24 = rng.integers(-(1 << 23), (1 << 23) - 1, size=200)
    payload = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in q24)
    write(tmp_path / "p24.wav", 1, 24, 1, payload)
    np.testing.assert_array_equal(pkg.wavio.read_wav(str(tmp_path / "p24.wav"))[0], (q24 / float(1 << 23)).astype(np.float32))
    u8 = rng.integers(0, 255, size=100).astype(np.uint8)
    write(tmp_path / "u8.wav", 1, 8, 1, u8.tobytes())
    np.testing.assert_array_equal(pkg.wavio.read_wav(str(tmp_path / "u8.wav"))[0], (u8.astype(np.float32) - 128) / 128)
