"""Minimal RIFF/WAVE PCM-16 mono I/O for the cache artefacts.

The reference reads audio with librosa.load (pipeline1.py:100-106) and writes the cleaned
clip with soundfile as WAV/PCM_16 (pipeline1.py:142).  Mono PCM-16 WAV is what the reference's
clear_audio/ directory holds; MP3 inputs go through mp3io.py, other rates through the GPU resampler.
"""
from __future__ import annotations

import struct

import numpy as np


def read_wav(path: str):
    """-> (float32[n] in [-1, 1), sr).  int16 / 32768 like librosa.load on a PCM-16 file."""
    pcm, sr = read_wav_pcm16(path)
    return (pcm.astype(np.float32) / np.float32(32768.0)), sr


def read_wav_pcm16(path: str):
    """-> (int16[n], sr): the samples as stored (the PCM-16 entry points of the library take them as they are)."""
    with open(path, "rb") as fh:
        blob = fh.read()
    if len(blob) < 12 or blob[0:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    fmt, data, pos = None, None, 12
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        if tag == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", blob, pos + 8)
        elif tag == b"data":
            data = blob[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{path}: fmt or data chunk missing")
    codec, channels, sr, _, _, bits = fmt
    if codec != 1 or bits != 16 or channels != 1:
        raise ValueError(f"{path}: only mono PCM-16 is supported (codec={codec}, channels={channels}, bits={bits})")
    return np.frombuffer(data[:len(data) & ~1], dtype="<i2").astype(np.int16), int(sr)


def write_wav_pcm16(path: str, pcm: np.ndarray, sr: int = 16000) -> None:
    """Writes int16 samples as the 44-byte-header WAV libsndfile produces for PCM_16 mono."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    n = pcm.size * 2
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + n) + b"WAVEfmt " +
                 struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16) + b"data" + struct.pack("<I", n))
        fh.write(pcm.tobytes())
