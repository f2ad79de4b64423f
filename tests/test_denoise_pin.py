"""Pins the oracle's LOAD + DENOISE chain to the reference's own golden pairs.

The reference's goldens for ``clean_audio_and_cache`` (pipeline1.py:126-146) are its 888
``segrigated_samples/**/<stem>.mp3 -> clear_audio/<stem>.wav`` pairs; for ``load_audio`` + ``extract_features`` on
the raw branch (pipeline1.py:100-106, :449) the 888 ``<stem>.mp3 -> cache_features/<stem>_raw_feats.npy`` pairs.
Both are committed (tests/golden/ref_corpus_*.npz).  Chain under test, CPU only:

    mp3io.decode_mp3 (FFmpeg mp3float, libmpg123's gapless conventions)  ->  oracle.resample (soxr_hq restatement)
    ->  oracle.denoise.clean_audio (noisereduce restatement + normalise + PCM-16)   vs the reference's WAV samples.

Bit-exactness is out of reach for two named reasons -- a different MP3 decoder implementation (float rounding) and a
restated (not linked) soxr filter whose 7.3-8 kHz skirt is not pinned -- so the comparison is statistical, with the
thresholds written here, and it is shown to DISCRIMINATE: every single spectral-gate parameter moved off its
noisereduce default drops the agreement by 25-50 dB on every clip.
Whole-corpus numbers (888 clips): profiles/r02_denoise_pin_corpus.json (tools/pin_denoise_corpus.py).
"""
import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import corpus  # noqa: E402

from conftest import PKG_NAME
from oracle import denoise, features, resample, wavio

mp3io = importlib.import_module(PKG_NAME + ".mp3io")

needs_fixtures = pytest.mark.skipif(not (corpus.have_wav() and corpus.have_mp3()), reason="corpus fixtures missing")
needs_decoder = pytest.mark.skipif(not mp3io.available(), reason="no libavcodec with mp3float in this image")


@pytest.fixture(scope="module")
def wav_corpus():
    return corpus.load_wav_corpus()


@pytest.fixture(scope="module")
def mp3_corpus():
    return corpus.load_mp3_corpus()


def snr_db(ref: np.ndarray, got: np.ndarray, drop_tail: int = 64) -> float:
    """SNR of the difference over all but the last 4 ms: clips are cut mid-word, and the ringing of the resampler's
    low-pass at that discontinuity depends on the un-pinned filter skirt."""
    m = max(1, len(ref) - drop_tail)
    r, g = ref[:m].astype(np.float64), got[:m].astype(np.float64)
    return float(10 * np.log10((r ** 2).sum() / max(((g - r) ** 2).sum(), 1e-9)))


@needs_fixtures
def test_decoded_lengths_reproduce_all_888_reference_lengths(wav_corpus, mp3_corpus):
    """len(clear_audio/<stem>.wav) == ceil(n_decoded * 16000 / sr) for every file: pins the gapless trimming
    (delay + 529 in front, padding - 529 at the end, nothing without an Info frame) and librosa's ceil."""
    _, _, offs, _ = wav_corpus
    names, blobs, _, _ = mp3_corpus
    kinds = {"info_plain": 0, "info_with_delay": 0, "untagged": 0}
    for i, blob in enumerate(blobs):
        frames = mp3io.parse_frames(blob)
        assert frames and frames[0][4] == 22050 and frames[0][5] == 3, names[i]          # 22 050 Hz mono, like SURVEY 2.1 says
        tag = mp3io.info_tag(blob, frames[0])
        kinds["untagged" if tag is None else ("info_with_delay" if tag["delay"] else "info_plain")] += 1
        n22 = mp3io.decoded_length(blob)
        assert -(-n22 * 320 // 441) == offs[i + 1] - offs[i], names[i]
    assert kinds == {"info_plain": 771, "info_with_delay": 111, "untagged": 6}


@needs_fixtures
@needs_decoder
def test_denoise_oracle_is_pinned_to_reference_wavs(wav_corpus, mp3_corpus):
    names, pcm, offs, clean_feats = wav_corpus
    _, blobs, raw_feats, _ = mp3_corpus
    picks = corpus.stratified(len(names), 96, np.diff(offs))
    snrs, mismatch, raw_err, clean_err = [], [], [], []
    for i in picks:
        y22, sr = mp3io.decode_mp3(blobs[i])
        y = resample.resample(y22, sr, 16000)
        q_ref = pcm[offs[i]:offs[i + 1]]
        assert len(y) == len(q_ref)
        q = denoise.clean_audio(y)
        assert q is not None and np.abs(q.astype(np.int32)).max() >= 32767                   # peak-normalised to full scale like every reference WAV
        snrs.append(snr_db(q_ref, q))
        mismatch.append(float(np.mean(q != q_ref)))
        raw_err.append(np.abs(features.extract_features(y) - raw_feats[i]))
        clean_err.append(np.abs(features.extract_features(wavio.dequantize_pcm16(q)) - clean_feats[i]))
    snrs, raw_err, clean_err = np.asarray(snrs), np.asarray(raw_err), np.asarray(clean_err)
    # whole corpus: median 70.9 dB, 5th percentile 57.2 dB, minimum 42.6 dB (profiles/r02_denoise_pin_corpus.json)
    assert np.median(snrs) >= 66.0, np.median(snrs)
    assert np.percentile(snrs, 5) >= 52.0, np.percentile(snrs, 5)
    assert snrs.min() >= 40.0, snrs.min()
    # at 70 dB the two PCM streams differ by about half an LSB rms: roughly every second sample is off by one
    assert 0.3 <= np.median(mismatch) <= 0.7
    # raw branch (decode + resample + features) against *_raw_feats.npy: MFCC values reach 500, the residual is the
    # top mel band sitting on the low-pass skirt (alternating sign over the coefficients)
    assert np.median(raw_err[:, :40].max(1)) <= 0.03 and np.median(raw_err[:, 40:120].max(1)) <= 0.003
    assert np.median(raw_err[:, 120:144].max(1)) <= 1e-3
    # clean branch end to end (decode + resample + gate + normalise + PCM-16 + features) against *_clean_feats.npy
    assert np.median(clean_err[:, :40].max(1)) <= 0.06 and np.median(clean_err[:, 40:120].max(1)) <= 0.01


PERTURBATIONS = {                                 # SURVEY A.6 items (i)-(ii) and the reference's own prop_decrease (pipeline1.py:140)
    "n_grad_freq 16->8": (1.0, {"n_grad_freq": 8}),
    "n_grad_time 3->1": (1.0, {"n_grad_time": 1}),
    "thresh 2->1.5": (1.0, {"thresh": 1.5}),
    "slope 10->5": (1.0, {"slope": 5.0}),
    "time_constant 2->1 s": (1.0, {"time_constant_s": 1.0}),
    "prop_decrease 1.0->0.8 (main1.py:605)": (0.8, None),
}


@needs_fixtures
@needs_decoder
def test_the_pin_discriminates(wav_corpus, mp3_corpus):
    """The same comparison with one noisereduce default changed must fail clearly on EVERY clip."""
    names, pcm, offs, _ = wav_corpus
    _, blobs, _, _ = mp3_corpus
    picks = corpus.stratified(len(names), 24, np.diff(offs))
    base, pert = [], {k: [] for k in PERTURBATIONS}
    for i in picks:
        y22, sr = mp3io.decode_mp3(blobs[i])
        y = resample.resample(y22, sr, 16000)
        q_ref = pcm[offs[i]:offs[i + 1]]
        base.append(snr_db(q_ref, denoise.clean_audio(y)))
        for k, (prop, pt) in PERTURBATIONS.items():
            pert[k].append(snr_db(q_ref, denoise.clean_audio(y, prop, pt)))
    base = np.asarray(base)
    for k, v in pert.items():
        v = np.asarray(v)
        assert np.all(v < base - 8.0), (k, v, base)                 # worse on every clip ...
        assert np.median(v) <= np.median(base) - 20.0, (k, np.median(v))   # ... and by a wide margin overall


@needs_fixtures
def test_wav_corpus_fixture_is_the_reference_corpus(wav_corpus):
    names, pcm, offs, feats = wav_corpus
    assert len(names) == 888 and feats.shape == (888, 149) and offs[-1] == len(pcm)
    lens = np.diff(offs)
    assert lens.min() == 7140 and lens.max() == 161367                 # 0.45 s .. 10.09 s (per_file_analysis.csv)
    for i in (0, 443, 887):                                            # every reference WAV peaks at full scale
        assert np.abs(pcm[offs[i]:offs[i + 1]].astype(np.int32)).max() >= 32767
