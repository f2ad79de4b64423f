"""Oracle: the reference's cleaning step.  TEST INFRASTRUCTURE ONLY.  PARITY: PINNED STATISTICALLY (see below).

Follows /root/reference/pipeline1.py:126-146 (``clean_audio_and_cache``):

    y_clean = nr.reduce_noise(y=y, sr=sr)          # :140  noisereduce defaults
    y_clean = librosa.util.normalize(y_clean)      # :141  peak normalise
    sf.write(out_path, y_clean, sr)                # :142  PCM-16 WAV
    ... later  librosa.load(cleaned)               # :389 / :437  int16/32768

``noisereduce`` (requirements.txt:6, unpinned; 3.x API) is NOT vendored in the
reference and NOT installable here.  This file therefore restates the published
algorithm of noisereduce 3.0.x ``SpectralGateNonStationary`` with the defaults that
``reduce_noise(y, sr)`` selects (stationary=False, prop_decrease=1.0, n_fft=1024,
hop=256, time_constant_s=2.0, freq_mask_smooth_hz=500, time_mask_smooth_ms=50,
thresh_n_mult_nonstationary=2, sigmoid_slope_nonstationary=10, chunk_size=600000,
padding=30000), calling the same scipy routines noisereduce calls
(scipy.signal.filtfilt, scipy.signal.fftconvolve) and librosa-equivalent STFT/ISTFT.
What IS checked against the reference's artefacts (tests/test_denoise_pin.py, whole corpus in
profiles/r02_denoise_pin_corpus.json): the reference's 888 ``segrigated_samples/**.mp3 ->
clear_audio/*.wav`` pairs.  Decoded (FFmpeg mp3float with libmpg123's trimming) and resampled
(oracle/resample.py), the output of this file agrees with the reference's WAV samples to a median
70.9 dB (5th percentile 57.2 dB) -- about half an LSB rms -- and every single default moved off its
value (n_grad_freq, n_grad_time, threshold, slope, time constant, prop_decrease) drops that by
25 - 50 dB on every clip.  Bit-exactness is out of reach for two named reasons: another MP3 decoder
implementation and a restated (not linked) soxr filter.  Also: length preservation and the
full-scale peak of every committed clear_audio/*.wav (normalise + quantise).
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.signal

from .features import hann_periodic, stft_frames
from .wavio import dequantize_pcm16, quantize_pcm16

SR = 16000
NR_N_FFT = 1024
NR_HOP = 256
NR_BINS = NR_N_FFT // 2 + 1       # 513
NR_PADDING = 30000
NR_CHUNK = 600000
NR_TIME_CONSTANT_S = 2.0
NR_THRESH = 2.0
NR_SLOPE = 10.0


@functools.lru_cache(maxsize=None)
def iir_coefficient() -> float:
    """noisereduce.spectralgate.utils.get_time_smoothed_representation."""
    t_frames = NR_TIME_CONSTANT_S * SR / float(NR_HOP)
    return float((np.sqrt(1 + 4 * t_frames ** 2) - 1) / (2 * t_frames ** 2))


def _tri(n: int) -> np.ndarray:
    return np.concatenate([np.linspace(0, 1, n + 1, endpoint=False), np.linspace(1, 0, n + 2)])[1:-1]


@functools.lru_cache(maxsize=None)
def smoothing_filter(n_grad_freq: int | None = None, n_grad_time: int | None = None) -> np.ndarray:
    """noisereduce _smoothing_filter(n_grad_freq=16, n_grad_time=3): [33, 7], sums to 1.
    (The arguments exist for the discrimination test only: tests perturb them and must FAIL against the goldens.)"""
    if n_grad_freq is None:
        n_grad_freq = int(500 / (SR / (NR_N_FFT / 2)))
    if n_grad_time is None:
        n_grad_time = int(50 / ((NR_HOP / SR) * 1000))
    f = np.outer(_tri(n_grad_freq), _tri(n_grad_time))
    return f / np.sum(f)


def nr_stft(x: np.ndarray) -> np.ndarray:
    """librosa.stft(x, n_fft=1024, hop_length=256) on float64 -> complex128 [513, T]."""
    frames = stft_frames(x, NR_N_FFT, NR_HOP)
    return np.fft.rfft(hann_periodic(NR_N_FFT)[None, :] * frames, axis=1).T


@functools.lru_cache(maxsize=None)
def _window_sumsquare(n_frames_: int) -> np.ndarray:
    """librosa.filters.window_sumsquare for hann/1024/256 (frames added in ascending order)."""
    n = NR_N_FFT + NR_HOP * (n_frames_ - 1)
    x = np.zeros(n, dtype=np.float64)
    wsq = hann_periodic(NR_N_FFT) ** 2
    for i in range(n_frames_):
        s = i * NR_HOP
        x[s:min(n, s + NR_N_FFT)] += wsq[:max(0, min(NR_N_FFT, n - s))]
    return x


def nr_istft(D: np.ndarray) -> np.ndarray:
    """librosa.istft(D, hop_length=256, win_length=1024) -> float64 [256 * (T - 1)]."""
    T = D.shape[1]
    frames = np.fft.irfft(D.T, n=NR_N_FFT, axis=1) * hann_periodic(NR_N_FFT)[None, :]
    full = np.zeros(NR_N_FFT + NR_HOP * (T - 1), dtype=np.float64)
    for t in range(T):
        full[t * NR_HOP:t * NR_HOP + NR_N_FFT] += frames[t]
    wss = _window_sumsquare(T)
    y = full[NR_N_FFT // 2:NR_N_FFT // 2 + NR_HOP * (T - 1)]
    w = wss[NR_N_FFT // 2:NR_N_FFT // 2 + NR_HOP * (T - 1)]
    nz = w > np.finfo(np.float64).tiny
    y = y.copy()
    y[nz] /= w[nz]
    return y


def spectral_gate_chunk(chunk: np.ndarray, prop_decrease: float = 1.0, return_parts: bool = False, perturb: dict | None = None):
    """SpectralGateNonStationary.spectral_gating_nonstationary on one padded float64 chunk.
    ``perturb`` (tests only) overrides thresh / slope / n_grad_freq / n_grad_time / time_constant_s."""
    pt = perturb or {}
    D = nr_stft(chunk)
    A = np.abs(D)
    b = iir_coefficient()
    if "time_constant_s" in pt:
        t_frames = pt["time_constant_s"] * SR / float(NR_HOP)
        b = float((np.sqrt(1 + 4 * t_frames ** 2) - 1) / (2 * t_frames ** 2))
    S = scipy.signal.filtfilt([b], [1, b - 1], A, axis=-1, padtype=None)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        above = (A - S) / S
        mask0 = 1 / (1 + np.exp(-(above + -pt.get("thresh", NR_THRESH)) * pt.get("slope", NR_SLOPE)))
    mask = scipy.signal.fftconvolve(mask0, smoothing_filter(pt.get("n_grad_freq"), pt.get("n_grad_time")), mode="same")
    mask = mask * prop_decrease + np.ones(np.shape(mask)) * (1.0 - prop_decrease)
    y = nr_istft(D * mask)
    out = np.zeros(chunk.shape, chunk.dtype)
    out[:len(y)] = y
    if return_parts:
        return out, dict(A=A, S=S, mask0=mask0, mask=mask)
    return out


def reduce_noise(y: np.ndarray, prop_decrease: float = 1.0, perturb: dict | None = None) -> np.ndarray:
    """nr.reduce_noise(y=y, sr=16000) for a 1-D float32 clip -> float32, same length.
    Inputs longer than 600 000 samples are gated in 600 000-sample chunks, each padded
    with 30 000 neighbouring samples (zeros outside the clip) -- SpectralGate.get_traces."""
    y = np.asarray(y)
    n = y.shape[0]

    def filter_chunk(start, end):
        i1, i2 = start - NR_PADDING, end + NR_PADDING
        a, bnd = max(i1, 0), min(i2, n)
        chunk = np.zeros(i2 - i1, dtype=np.float64)
        if bnd > a:
            chunk[a - i1:bnd - i1] = y[a:bnd]
        return spectral_gate_chunk(chunk, prop_decrease, perturb=perturb)[start - i1:end - i1]

    if n > NR_CHUNK:
        out = np.zeros(n, dtype=y.dtype)
        nchunks = int((n - 1) / NR_CHUNK) + 1
        for ich in range(nchunks):
            s0 = ich * NR_CHUNK
            e0 = min(n, s0 + NR_CHUNK)
            out[s0:e0] = filter_chunk(s0, s0 + NR_CHUNK)[:e0 - s0]
        return out.astype(y.dtype)
    return filter_chunk(0, n).astype(y.dtype)


def peak_normalize(y: np.ndarray) -> np.ndarray:
    """librosa.util.normalize(y) (norm=inf): raises on non-finite input; peak < tiny -> unchanged."""
    if not np.isfinite(y).all():
        raise ValueError("Input must be finite")
    mag = np.abs(y).astype(float)
    length = np.max(mag, axis=0, keepdims=True)
    length[length < np.finfo(y.dtype).tiny] = 1.0
    out = np.empty_like(y)
    out[:] = y / length
    return out


def clean_audio(y: np.ndarray, prop_decrease: float = 1.0, perturb: dict | None = None):
    """In-memory clean_audio_and_cache (pipeline1.py:136-143): returns the int16 PCM that
    the reference writes to clear_audio/<stem>.wav, or ``None`` where the reference's
    ``except`` branch fires (pipeline1.py:144-146), e.g. the all-zero clip (0/0 -> NaN)."""
    if y is None:
        return None
    y = np.asarray(y, dtype=np.float32)
    try:
        if y.size == 0:
            raise ValueError("empty")
        return quantize_pcm16(peak_normalize(reduce_noise(y, prop_decrease, perturb)))
    except Exception:
        return None


def clean_then_load(y: np.ndarray, prop_decrease: float = 1.0) -> np.ndarray:
    """What the feature function sees on the clean branch: the reloaded WAV, or -- when
    cleaning failed -- the raw clip itself (pipeline1.py:385-387)."""
    q = clean_audio(y, prop_decrease)
    if q is None:
        return np.asarray(y, dtype=np.float32)
    return dequantize_pcm16(q)
