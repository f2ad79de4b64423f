"""Loader of the committed whole-corpus fixtures (written by make_golden_corpus.py from the reference's artefacts).

  ref_corpus_wav.npz : the 888 ``clear_audio/<stem>.wav`` payloads (int16 PCM) + ``<stem>_clean_feats.npy``
  ref_corpus_mp3.npz : the 888 ``segrigated_samples/**/<stem>.mp3`` files (bytes) + ``<stem>_raw_feats.npy``
PCM is stored as LZMA over the two byte planes of its first difference (peak-normalised speech: 54 % of raw).
"""
import lzma
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
WAV_NPZ = os.path.join(HERE, "ref_corpus_wav.npz")
MP3_NPZ = os.path.join(HERE, "ref_corpus_mp3.npz")


def pack_pcm(pcm: np.ndarray) -> np.ndarray:
    d = np.diff(pcm.astype(np.int32), prepend=0).astype(np.int16).view(np.uint16)       # wraps mod 2^16: invertible
    planes = np.concatenate([(d & 255).astype(np.uint8), (d >> 8).astype(np.uint8)])
    return np.frombuffer(lzma.compress(planes.tobytes(), preset=6), dtype=np.uint8)


def unpack_pcm(blob: np.ndarray) -> np.ndarray:
    planes = np.frombuffer(lzma.decompress(blob.tobytes()), dtype=np.uint8)
    n = planes.size // 2
    d = planes[:n].astype(np.uint16) | (planes[n:].astype(np.uint16) << 8)
    return np.cumsum(d, dtype=np.uint16).view(np.int16)          # running sum mod 2^16 undoes the wrapped difference


def have_wav() -> bool:
    return os.path.exists(WAV_NPZ)


def have_mp3() -> bool:
    return os.path.exists(MP3_NPZ)


def load_wav_corpus():
    """-> (names [888], pcm int16 (concatenated), offsets int64 [889], clean_feats float32 [888, 149])"""
    z = np.load(WAV_NPZ)
    return z["names"], unpack_pcm(z["pcm_packed"]), z["offsets"], z["clean_feats"]


def load_mp3_corpus():
    """-> (names [888], list of bytes, raw_feats float32 [888, 149], labels [888])"""
    z = np.load(MP3_NPZ)
    blob, offs = z["mp3_bytes"].tobytes(), z["offsets"]
    return z["names"], [blob[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)], z["raw_feats"], z["labels"]


def stratified(n_total: int, count: int, lengths=None):
    """Indices of ``count`` clips spread evenly over the length-sorted corpus (shortest and longest included)."""
    order = np.argsort(lengths, kind="stable") if lengths is not None else np.arange(n_total)
    picks = sorted({int(order[int(round(i))]) for i in np.linspace(0, n_total - 1, count)})
    return np.asarray(picks, dtype=np.int64)
