"""Oracle: dependency-free RIFF/PCM-16 reader + writer.  TEST INFRASTRUCTURE ONLY.

Restates the two I/O conventions on the reference's clean branch:
  * pipeline1.py:142  ``sf.write(out_path, y_clean, sr)``  -> WAV / PCM_16 with
    libsndfile's float->short conversion  clip(lrintf(x * 32768), -32768, 32767);
  * pipeline1.py:102  ``librosa.load(path, sr=16000, mono=True)`` on such a file
    -> int16 / 32768 as float32 (exact).
"""
from __future__ import annotations

import struct

import numpy as np


def quantize_pcm16(y: np.ndarray) -> np.ndarray:
    """float32 -> int16 exactly like libsndfile's f2s_clip_array (normalised floats)."""
    scaled = np.asarray(y, dtype=np.float32) * np.float32(32768.0)
    q = np.rint(scaled)                       # round-half-even == lrintf
    return np.clip(q, -32768.0, 32767.0).astype(np.int16)


def dequantize_pcm16(q: np.ndarray) -> np.ndarray:
    return (np.asarray(q, dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def read_wav_pcm16(path: str):
    """Returns (int16[n], sr).  Mono PCM-16 only (what the reference writes)."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos = 12
    fmt = None
    pcm = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", body, 0)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, nch, sr, _, _, bits = fmt
    if tag != 1 or nch != 1 or bits != 16:
        raise ValueError(f"{path}: expected mono PCM-16, got tag={tag} ch={nch} bits={bits}")
    return np.frombuffer(pcm[:len(pcm) // 2 * 2], dtype="<i2").copy(), sr


def write_wav_pcm16(path: str, q: np.ndarray, sr: int = 16000) -> None:
    q = np.asarray(q, dtype="<i2")
    nbytes = q.size * 2
    hdr = b"RIFF" + struct.pack("<I", 36 + nbytes) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16) + b"data" + struct.pack("<I", nbytes)
    with open(path, "wb") as fh:
        fh.write(hdr)
        fh.write(q.tobytes())
