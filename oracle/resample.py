"""Oracle: sample-rate conversion of ``librosa.load(path, sr=16000)``.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pipeline1.py:100-106: ``librosa.load`` decodes at the file's own rate (the corpus:
MPEG-2 Layer III, 22 050 Hz mono), then ``librosa.resample(..., res_type="soxr_hq")`` -> python-soxr ``HQ`` and
``util.fix_length(ceil(n * 16000 / 22050))``.  libsoxr is un-vendored and not installable here, so its *published
filter specification* is restated instead of its code (PARITY: statistical, see tests/test_oracle_golden.py):

  soxr_quality_spec(SOXR_HQ): precision 20 bit, linear phase, stopband_begin = 1.0 (the lower Nyquist),
      passband_end = 1 - 0.05 / TO_3dB(rej),  rej = 20 * 6.0206 dB,  TO_3dB(a) = (1.6e-6 a - 7.5e-4) a + 0.646  -> 0.91363
  one steep low-pass decides the response: a Kaiser-windowed sinc designed on the grid of twice the lower rate, -6 dB
      point midway between pass- and stop-band edge (7 654.5 Hz for 16 kHz), attenuation (20 + 1) * 6.0206 = 126.4 dB,
      beta = 0.1102 (att - 8.7), taps = (att - 7.95) / (2.285 * transition) + 1 rounded up to 1 mod 4 (385 at 32 kHz);
      every other stage of soxr is flat over the band this filter passes.
  output sample m is taken at input time m * sr_in / sr_out (delay compensated), the input is zero outside [0, n).

The low-pass is applied here as ONE continuous-time kernel evaluated at the exact fractional offsets (float64), which
is what soxr's stage cascade approximates to its 20-bit accuracy.  What this restatement cannot pin is the exact shape
of the 7.3 - 8.0 kHz skirt: energy there (the corpus' MP3 encoder cuts at ~7.8 kHz) lands in the top mel band, which is
where the residual against the reference's ``*_raw_feats.npy`` sits (alternating-sign MFCC error ~1e-2).
"""
from __future__ import annotations

import functools
from math import ceil, gcd

import numpy as np
from scipy.special import i0

SOXR_HQ_BITS = 20
_DB_PER_BIT = 20.0 * np.log10(2.0)


def soxr_hq_spec():
    """-> (passband_end, stopband_begin, attenuation dB) of soxr's HQ recipe."""
    rej = SOXR_HQ_BITS * _DB_PER_BIT
    to_3db = (1.6e-6 * rej - 7.5e-4) * rej + 0.646
    return 1.0 - 0.05 / to_3db, 1.0, (SOXR_HQ_BITS + 1) * _DB_PER_BIT


def lowpass_design(sr_in: int, sr_out: int):
    """-> (cutoff Hz, window half-width seconds, Kaiser beta) of the steep low-pass."""
    fp, fs, att = soxr_hq_spec()
    low = min(sr_in, sr_out)
    grid = 2.0 * low                               # rate the filter is designed at
    tr = 0.5 * (fs - fp) * (low / 2.0) / (grid / 2.0)            # 6 dB -> stop, as a fraction of the grid's Nyquist
    fc = (fs * (low / 2.0)) / (grid / 2.0) - tr
    beta = 0.1102 * (att - 8.7)
    taps = int(ceil((att - 7.95) / (2.285 * 2.0 * np.pi * tr) + 1))
    taps = (taps + 2) // 4 * 4 + 1
    half_width_s = (0.5 * (taps - 1) + 0.5) / grid
    return fc * grid / 2.0, half_width_s, beta


@functools.lru_cache(maxsize=None)
def phase_table(sr_in: int, sr_out: int):
    """Polyphase form of the kernel: -> (h float64 [phases, 2 * half + 1], half, up, down) with
    y[m] = sum_j h[(m * down) % up, j] * x[(m * down) // up + j - half]."""
    g = gcd(sr_in, sr_out)
    up, down = sr_out // g, sr_in // g
    fc, hw, beta = lowpass_design(sr_in, sr_out)
    t_in = hw * sr_in                                     # window half-width in input samples
    half = int(ceil(t_in))
    frac = np.arange(up, dtype=np.float64) / up
    d = np.arange(-half, half + 1, dtype=np.float64)[None, :] - frac[:, None]      # tap position relative to the output instant
    u = d / t_in
    win = np.where(np.abs(u) < 1.0, i0(beta * np.sqrt(np.clip(1.0 - u * u, 0.0, 1.0))) / i0(beta), 0.0)
    f = fc / (sr_in / 2.0)
    return f * np.sinc(f * d) * win, half, up, down


def resample(x: np.ndarray, sr_in: int, sr_out: int = 16000) -> np.ndarray:
    """librosa.resample(x, orig_sr=sr_in, target_sr=sr_out, res_type="soxr_hq") -> float32 [ceil(n * sr_out / sr_in)]."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if sr_in == sr_out:
        return x.astype(np.float32)
    h, half, up, down = phase_table(int(sr_in), int(sr_out))
    n = x.shape[0]
    m = int(ceil(n * up / down))
    k = np.arange(m, dtype=np.int64) * down
    first, phase = k // up, k % up
    xp = np.concatenate([np.zeros(half), x, np.zeros(half + 2)])
    out = np.empty(m, dtype=np.float64)
    taps = np.arange(2 * half + 1)
    for s in range(0, m, 8192):
        idx = first[s:s + 8192, None] + taps[None, :]
        out[s:s + 8192] = np.einsum("ij,ij->i", xp[idx], h[phase[s:s + 8192]])
    return out.astype(np.float32)


def load_audio(blob_or_samples, sr_in: int | None = None, sr: int = 16000) -> np.ndarray:
    """``librosa.load(path, sr=16000, mono=True)`` for already-decoded mono samples at ``sr_in``."""
    return resample(np.asarray(blob_or_samples, dtype=np.float32), int(sr_in), sr)
