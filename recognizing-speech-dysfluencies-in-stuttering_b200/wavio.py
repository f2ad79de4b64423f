"""Minimal RIFF/WAVE PCM-16 mono I/O for the cache artefacts.

The reference reads audio with librosa.load (pipeline1.py:100-106) and writes the cleaned
clip with soundfile as WAV/PCM_16 (pipeline1.py:142).  Mono PCM-16 WAV is what the reference's
clear_audio/ directory holds; other PCM widths, IEEE-float and multi-channel files are read like libsndfile + librosa's
mono=True would; MP3 inputs go through mp3io.py, other rates through the GPU resampler.
"""
from __future__ import annotations

import struct

import numpy as np


def _chunks(path: str):
    with open(path, "rb") as fh:
        blob = fh.read()
    if len(blob) < 12 or blob[0:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    fmt, data, pos = None, None, 12
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        if tag == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", blob, pos + 8)
            if fmt[0] == 0xFFFE and size >= 26:                       # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                fmt = (struct.unpack_from("<H", blob, pos + 8 + 24)[0],) + fmt[1:]
        elif tag == b"data":
            data = blob[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{path}: fmt or data chunk missing")
    return fmt, data


def read_wav(path: str):
    """-> (float32[n] mono, sr): what ``librosa.load(path, sr=None, mono=True)`` returns for a RIFF/WAVE file, i.e.
    libsndfile's float conversion (PCM-16: q / 32768, PCM-24: q / 2^23, PCM-32: q / 2^31, PCM-8: (u - 128) / 128, IEEE
    float as stored) followed by ``np.mean`` over the channels in float32."""
    fmt, data = _chunks(path)
    codec, channels, sr, _, _, bits = fmt
    if channels < 1:
        raise ValueError(f"{path}: no channels")
    if codec == 1 and bits == 16:
        y = np.frombuffer(data[:len(data) // 2 * 2], dtype="<i2").astype(np.float32) / np.float32(32768.0)
    elif codec == 1 and bits == 8:
        y = (np.frombuffer(data, dtype=np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif codec == 1 and bits == 24:
        raw = np.frombuffer(data[:len(data) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        q = (raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16))
        q = np.where(q >= 1 << 23, q - (1 << 24), q)
        y = (q.astype(np.float64) / float(1 << 23)).astype(np.float32)
    elif codec == 1 and bits == 32:
        y = (np.frombuffer(data[:len(data) // 4 * 4], dtype="<i4").astype(np.float64) / float(1 << 31)).astype(np.float32)
    elif codec == 3 and bits == 32:
        y = np.frombuffer(data[:len(data) // 4 * 4], dtype="<f4").astype(np.float32)
    elif codec == 3 and bits == 64:
        y = np.frombuffer(data[:len(data) // 8 * 8], dtype="<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE encoding (codec={codec}, bits={bits})")
    if channels > 1:
        y = y[:len(y) // channels * channels].reshape(-1, channels).mean(axis=1, dtype=np.float32)      # librosa.to_mono
    return np.ascontiguousarray(y, dtype=np.float32), int(sr)


def read_wav_pcm16(path: str):
    """-> (int16[n], sr): the samples as stored, for mono PCM-16 files (the PCM-16 entry points of the library take them
    as they are); every other encoding raises ValueError -- use ``read_wav``."""
    fmt, data = _chunks(path)
    codec, channels, sr, _, _, bits = fmt
    if codec != 1 or bits != 16 or channels != 1:
        raise ValueError(f"{path}: not mono PCM-16 (codec={codec}, channels={channels}, bits={bits})")
    return np.frombuffer(data[:len(data) & ~1], dtype="<i2").astype(np.int16), int(sr)


def write_wav_pcm16(path: str, pcm: np.ndarray, sr: int = 16000) -> None:
    """Writes int16 samples as the 44-byte-header WAV libsndfile produces for PCM_16 mono."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    n = pcm.size * 2
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + n) + b"WAVEfmt " +
                 struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16) + b"data" + struct.pack("<I", n))
        fh.write(pcm.tobytes())
