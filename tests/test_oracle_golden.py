"""Pins the CPU oracle to the reference's own artefacts.

The reference ships no tests; its committed ``clear_audio/*.wav`` ->
``cache_features/*_clean_feats.npy`` pairs and ``output_results/scaler_after.pkl`` are
the de-facto golden vectors of the hot path (pipeline1.py:206-265, :429-440, :470-473).
"""
import glob
import os

import numpy as np
import pytest

from conftest import REFERENCE_DIR, have_reference
from oracle import cmvn, denoise, features, wavio

ATOL, RTOL = 1e-3, 1e-4          # BASELINE.json north_star tolerance


def test_committed_golden_pairs(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_clean_pairs.npz"))
    offs = g["offsets"]
    assert len(offs) - 1 == 24
    for i in range(len(offs) - 1):
        y = wavio.dequantize_pcm16(g["pcm"][offs[i]:offs[i + 1]])
        got = features.extract_features(y, 16000)
        assert got.dtype == np.float32 and got.shape == (149,)
        np.testing.assert_allclose(got, g["feats"][i], atol=ATOL, rtol=RTOL, err_msg=str(g["names"][i]))
        assert not got[144:].any()


@pytest.mark.skipif(not have_reference(), reason="/root/reference only exists in the build container")
def test_all_888_reference_pairs():
    files = sorted(glob.glob(f"{REFERENCE_DIR}/cache_features/*_clean_feats.npy"))
    assert len(files) == 888
    worst = 0.0
    for f in files:
        stem = os.path.basename(f)[:-len("_clean_feats.npy")]
        q, sr = wavio.read_wav_pcm16(f"{REFERENCE_DIR}/clear_audio/{stem}.wav")
        got = features.extract_features(wavio.dequantize_pcm16(q), sr)
        ref = np.load(f)
        np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL, err_msg=stem)
        worst = max(worst, float(np.abs(got - ref).max()))
    assert worst < 2e-4


def test_scaler_golden_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_scaler_after.npz"))
    mean, var, scale, n = cmvn.fit(g["X"])
    assert n == int(g["n"]) == 905
    assert np.array_equal(mean, g["mean"])
    assert np.array_equal(var, g["var"])
    assert np.array_equal(scale, g["scale"])
    assert np.all(scale[144:] == 1.0)           # zero-variance text tail -> scale 1
    m2, v2, s2, _ = cmvn.from_moments(cmvn.moments(g["X"]))
    np.testing.assert_allclose(m2, mean, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(s2, scale, rtol=1e-9, atol=1e-12)


def _published_rf_metrics(Z, labels):
    """The reference's 'after' RandomForest leg (pipeline1.py:462-531): LabelEncoder, stratified 80/20 split with
    random_state=42, RandomForestClassifier(200, random_state=42) -> (accuracy %, log-loss)."""
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.metrics import accuracy_score, log_loss
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import LabelEncoder
    y = LabelEncoder().fit_transform(labels)
    Xtr, Xte, ytr, yte = train_test_split(Z, y, test_size=0.2, stratify=y, random_state=42)
    rf = RandomForestClassifier(n_estimators=200, random_state=42).fit(Xtr, ytr)
    return float(accuracy_score(yte, rf.predict(Xte)) * 100.0), float(log_loss(yte, rf.predict_proba(Xte)))


def test_classifier_feed_reproduces_published_metrics(golden_dir):
    """Scaler apply + the reference classifier on the reference's own cached features reproduces the accuracy and
    log-loss the reference published (output_results/metrics_summary.csv): pins the classifier input loader."""
    from sklearn.preprocessing import StandardScaler
    g = np.load(os.path.join(golden_dir, "ref_scaler_after.npz"))
    c = np.load(os.path.join(golden_dir, "ref_classifier_after.npz"))
    mean, var, scale, _ = cmvn.fit(g["X"])
    Z = cmvn.transform(g["X"], mean, scale)
    assert Z.dtype == np.float32
    assert np.array_equal(Z, StandardScaler().fit(g["X"]).transform(g["X"]))         # bit-for-bit sklearn
    acc, loss = _published_rf_metrics(Z, c["labels"])
    i = list(c["models"]).index("RandomForest")
    assert abs(acc - float(c["accuracy"][i])) < 1e-9 and abs(loss - float(c["test_loss"][i])) < 1e-12
    assert len(c["labels"]) == 905 and sorted(set(c["labels"])) == ["Prolongatio sample", "syllable repetition",
                                                                    "word repetition"]


def test_qc_metrics_match_reference_csv(golden_dir):
    """snr_db / spectral_flatness_mean / high_freq_energy_ratio (pipeline1.py:151-186) on the committed cleaned WAVs
    against the *_after columns the reference wrote to per_file_analysis.csv."""
    from oracle import qc
    pairs = np.load(os.path.join(golden_dir, "ref_clean_pairs.npz"))
    gq = np.load(os.path.join(golden_dir, "ref_qc_after.npz"))
    assert list(gq["names"]) == list(pairs["names"])
    offs = pairs["offsets"]
    for i in range(len(offs) - 1):
        y = wavio.dequantize_pcm16(pairs["pcm"][offs[i]:offs[i + 1]])
        m = qc.qc_metrics(y)
        assert abs(m[0] - gq["snr"][i]) < 1e-4, (i, m[0], gq["snr"][i])                 # dB
        assert abs(m[1] - gq["flat"][i]) <= 5e-4 * gq["flat"][i], (i, m[1], gq["flat"][i])
        assert abs(m[2] - gq["hf"][i]) <= 1e-5 * gq["hf"][i], (i, m[2], gq["hf"][i])
    assert qc.snr_db(np.zeros(399, np.float32)) == 0.0 and qc.snr_db(None) == 0.0       # pipeline1.py:155-156


def test_wav_round_trip_and_full_scale_property(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_clean_pairs.npz"))
    offs = g["offsets"]
    for i in range(len(offs) - 1):
        q = g["pcm"][offs[i]:offs[i + 1]]
        # every WAV the reference wrote peaks at full scale: normalise -> clip(rint(x * 32768))
        assert max(int(q.max()), -int(q.min())) >= 32767
    q = g["pcm"][offs[0]:offs[1]]
    p = tmp_path / "x.wav"
    wavio.write_wav_pcm16(str(p), q)
    back, sr = wavio.read_wav_pcm16(str(p))
    assert sr == 16000 and np.array_equal(back, q)
    y = wavio.dequantize_pcm16(q)
    assert np.array_equal(wavio.quantize_pcm16(y), q)
    assert wavio.quantize_pcm16(np.array([1.0, -1.0, 0.99999, 1.5e-5, 4.6e-5], np.float32)).tolist() == [
        32767, -32768, 32767, 0, 2]


def test_frame_counts_and_short_clip_zeros(synth):
    for n in (4096, 4607, 4608, 47999, 48000, 48001):
        y = synth.synth_clip(7, n)
        it = features.intermediates(y)
        T = 1 + n // 512
        assert it["power"].shape == (1025, T) and it["mfcc"].shape == (20, T) and it["chroma"].shape == (12, T)
    # T < 9 frames: librosa.feature.delta raises -> reference returns zeros(144)  (pipeline_errors.log)
    assert not features.extract_features(synth.synth_clip(3, 4095)).any()
    assert not features.extract_features(np.zeros(0, np.float32)).any()
    assert not features.extract_features(None).any()
    bad = synth.synth_clip(3, 8000)
    bad[17] = np.nan
    assert not features.extract_features(bad).any()


def test_tables_match_closed_forms():
    W = features.mel_filterbank()
    assert W.shape == (128, 1025) and W.dtype == np.float32
    assert int((W != 0).sum()) == 2020
    assert ((W != 0).sum(axis=0) <= 2).all() and not W[:, 0].any() and not W[:, 1024].any()
    L = np.random.default_rng(0).normal(size=(128, 11)).astype(np.float32)
    np.testing.assert_allclose(features.dct_matrix() @ L.astype(np.float64), features.mfcc_from_logmel(L),
                               atol=2e-5)
    assert abs(denoise.iir_coefficient() - 0.007968063999744) < 1e-14
    F = denoise.smoothing_filter()
    assert F.shape == (33, 7) and abs(F.sum() - 1) < 1e-15


def test_delta_matches_scipy_savgol():
    import scipy.signal
    x = np.random.default_rng(1).normal(size=(20, 37)).astype(np.float32)
    for order in (1, 2):
        ref = scipy.signal.savgol_filter(x, 9, deriv=order, polyorder=order, axis=-1, mode="interp")
        np.testing.assert_allclose(features.delta(x, order), ref, atol=1e-6)
    with pytest.raises(ValueError):
        features.delta(x[:, :8], 1)


def test_denoise_properties(synth):
    y = synth.synth_clip(0)
    out = denoise.reduce_noise(y)
    assert out.dtype == np.float32 and out.shape == y.shape          # length preserved (per_file_analysis.csv)
    q = denoise.clean_audio(y)
    assert q.dtype == np.int16 and max(int(q.max()), -int(q.min())) >= 32767
    # the gate must actually attenuate the silent gaps of the synthetic clip
    e_in = float(np.mean(y.astype(np.float64) ** 2))
    e_out = float(np.mean(out.astype(np.float64) ** 2))
    assert 0 < e_out < e_in
    # prop_decrease = 0 leaves the signal (STFT/ISTFT is a perfect reconstruction in the interior)
    np.testing.assert_allclose(denoise.reduce_noise(y, prop_decrease=0.0), y, atol=1e-6)
    # all-zero clip: 0/0 -> NaN -> normalize raises -> reference falls back to the raw file
    z = np.zeros(48000, np.float32)
    assert denoise.clean_audio(z) is None
    assert np.array_equal(denoise.clean_then_load(z), z)


def test_denoise_iir_closed_form():
    """filtfilt([b],[1,b-1], padtype=None) == forward then backward one-pole recursions with
    steady-state initial conditions (SURVEY Appendix A.6 iv) -- the form the CUDA kernel uses."""
    import scipy.signal
    rng = np.random.default_rng(2)
    A = np.abs(rng.normal(size=(5, 60)))
    b = denoise.iir_coefficient()
    ref = scipy.signal.filtfilt([b], [1, b - 1], A, axis=-1, padtype=None)
    f = np.empty_like(A)
    prev = A[:, 0].copy()
    for t in range(A.shape[1]):
        prev = b * A[:, t] + (1 - b) * prev
        f[:, t] = prev
    S = np.empty_like(A)
    nxt = f[:, -1].copy()
    for t in range(A.shape[1] - 1, -1, -1):
        nxt = b * f[:, t] + (1 - b) * nxt
        S[:, t] = nxt
    np.testing.assert_allclose(S, ref, rtol=1e-12)


def test_denoise_chunking_long_clip(synth):
    """> 600 000 samples: chunked with 30 000-sample real overlap (SpectralGate.get_traces)."""
    y = np.tile(synth.synth_clip(1), 13)[:610000]
    out = denoise.reduce_noise(y)
    assert out.shape == y.shape and np.isfinite(out).all()
    # far from the chunk seam the chunked result equals gating the first chunk alone with its padding
    chunk0 = np.zeros(660000)
    chunk0[30000:30000 + 610000] = y            # 30 000 zeros | samples 0..609 999 | zeros
    ref = denoise.spectral_gate_chunk(chunk0)
    np.testing.assert_array_equal(out[:600000], ref[30000:630000].astype(np.float32))
    chunk1 = np.zeros(660000)
    chunk1[:40000] = y[570000:]                 # samples 570 000..609 999 | zeros
    ref1 = denoise.spectral_gate_chunk(chunk1)
    np.testing.assert_array_equal(out[600000:], ref1[30000:40000].astype(np.float32))
