"""MPEG audio (Layer III) file reader for ``load_audio`` -- host-side I/O, not arithmetic of the hot path.

The reference reads its 905 inputs with ``librosa.load(path, sr=16000, mono=True)`` (pipeline1.py:100-106), which
hands MP3 files to libsndfile >= 1.1 (python-soundfile), i.e. to libmpg123 with gapless decoding on.  This module
reproduces that reader's OBSERVABLE conventions on top of FFmpeg's ``mp3float`` decoder (libavcodec is in the image
inside the opencv wheel; it is loaded with ctypes, nothing is linked or vendored):

  * ID3v2 / ID3v1 tags are skipped, frames are located by their sync words;
  * a leading Xing/Info frame is metadata, not audio;
  * gapless trimming like libmpg123: with an Info frame the first ``delay + 529`` decoded samples are dropped and the
    stream ends at ``total - padding + 529`` (clamped to what was decoded); without one nothing is trimmed.

Evidence that this is the reference's convention: with it ``ceil(n * 16000 / 22050)`` equals the length of the
committed ``clear_audio/<stem>.wav`` for all 888 stems (788 Lavf-tagged files: 529 samples dropped; 100 Lavc-tagged
files: delay 576 + padding 6xx-10xx; 6 untagged files: none) -- tests/test_oracle_golden.py.

The decoder library is optional: without it ``decode_mp3`` raises ``Mp3Unavailable`` and ``load_audio`` logs and
returns ``(None, None)`` like the reference does for an unreadable file.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import struct
import sys

import numpy as np

_BITRATES_V1 = (0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0)
_BITRATES_V2 = (0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0)
_RATES = {3: (44100, 48000, 32000), 2: (22050, 24000, 16000), 0: (11025, 12000, 8000)}
DECODER_DELAY = 529           # libmpg123's GAPLESS_DELAY: 528 samples of filterbank delay + 1


class Mp3Unavailable(RuntimeError):
    pass


class Mp3Frame(tuple):
    """(offset, size, version bits, bitrate, sample rate, channel mode, protection bit)"""


def parse_frames(blob: bytes):
    """Layer III frames of ``blob`` -> list of (offset, size, version, bitrate, sr, mode, no_crc)."""
    pos, end = 0, len(blob)
    if blob[:3] == b"ID3" and len(blob) >= 10:
        pos = 10 + ((blob[6] << 21) | (blob[7] << 14) | (blob[8] << 7) | blob[9])
    if end >= 128 and blob[end - 128:end - 125] == b"TAG":
        end -= 128
    frames = []
    while pos + 4 <= end:
        h = struct.unpack_from(">I", blob, pos)[0]
        ver, layer, bri, sri = (h >> 19) & 3, (h >> 17) & 3, (h >> 12) & 15, (h >> 10) & 3
        if (h >> 21) != 0x7FF or ver == 1 or layer != 1 or bri in (0, 15) or sri == 3:
            pos += 1
            continue
        sr = _RATES[ver][sri]
        br = (_BITRATES_V1 if ver == 3 else _BITRATES_V2)[bri] * 1000
        size = (144 if ver == 3 else 72) * br // sr + ((h >> 9) & 1)
        if pos + size > end:
            break
        frames.append((pos, size, ver, br, sr, (h >> 6) & 3, (h >> 16) & 1))
        pos += size
    return frames


def info_tag(blob: bytes, frame):
    """Xing/Info header of ``frame`` -> dict(delay, padding, frames) or None."""
    pos, size, ver, _, _, mode, no_crc = frame
    side = (17 if mode == 3 else 32) if ver == 3 else (9 if mode == 3 else 17)
    off = pos + 4 + (0 if no_crc else 2) + side
    if blob[off:off + 4] not in (b"Xing", b"Info"):
        return None
    flags = struct.unpack_from(">I", blob, off + 4)[0]
    p = off + 8
    n_frames = None
    if flags & 1:
        n_frames = struct.unpack_from(">I", blob, p)[0]
        p += 4
    if flags & 2:
        p += 4
    if flags & 4:
        p += 100
    if flags & 8:
        p += 4
    delay = padding = 0
    if p + 24 <= pos + size:                       # LAME extension: 21 bytes in, 12 bits delay + 12 bits padding
        d = blob[p + 21:p + 24]
        delay, padding = (d[0] << 4) | (d[1] >> 4), ((d[1] & 15) << 8) | d[2]
    return {"delay": delay, "padding": padding, "frames": n_frames}


# ------------------------------------------------------------------------------------------
# FFmpeg binding (ctypes; AVPacket / AVFrame field offsets of libavcodec >= 58: stable public ABI heads)
# ------------------------------------------------------------------------------------------
_PKT_DATA, _PKT_SIZE = 24, 32               # AVPacket: buf, pts, dts, data, size
_FRM_NB_SAMPLES, _FRM_FORMAT = 112, 116     # AVFrame: data[8], linesize[8], extended_data, width, height, nb_samples, format
_av = None


def _lib_dirs():
    dirs = []
    env = os.environ.get("DYS_FFMPEG_LIBDIR")
    if env:
        dirs.append(env)
    for sp in sys.path:
        dirs.append(os.path.join(sp, "opencv_python_headless.libs"))
        dirs.append(os.path.join(sp, "opencv_python.libs"))
        dirs.append(os.path.join(sp, "av.libs"))
    dirs += ["/usr/lib/x86_64-linux-gnu", "/usr/local/lib", "/usr/lib"]
    return [d for d in dirs if os.path.isdir(d)]


def _load_av():
    global _av
    if _av is not None:
        return _av
    for d in _lib_dirs():
        codec = sorted(glob.glob(os.path.join(d, "libavcodec*.so*")))
        util = sorted(glob.glob(os.path.join(d, "libavutil*.so*")))
        if not codec or not util:
            continue
        try:
            for dep in ("libdrm", "libcrypto", "libssl", "libvpx", "libaom"):        # wheel-private dependencies, if any
                for p in sorted(glob.glob(os.path.join(d, dep + "*.so*"))):
                    try:
                        C.CDLL(p, mode=C.RTLD_GLOBAL)
                    except OSError:
                        pass
            avutil = C.CDLL(util[0], mode=C.RTLD_GLOBAL)
            for p in sorted(glob.glob(os.path.join(d, "libswresample*.so*"))):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
            avcodec = C.CDLL(codec[0], mode=C.RTLD_GLOBAL)
        except OSError:
            continue
        vp = C.c_void_p
        avcodec.avcodec_find_decoder_by_name.restype = vp
        avcodec.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
        avcodec.avcodec_alloc_context3.restype = vp
        avcodec.avcodec_alloc_context3.argtypes = [vp]
        avcodec.avcodec_open2.argtypes = [vp, vp, vp]
        avcodec.av_packet_alloc.restype = vp
        avcodec.av_packet_free.argtypes = [vp]
        avcodec.av_new_packet.argtypes = [vp, C.c_int]
        avcodec.av_packet_unref.argtypes = [vp]
        avcodec.avcodec_send_packet.argtypes = [vp, vp]
        avcodec.avcodec_receive_frame.argtypes = [vp, vp]
        avcodec.avcodec_free_context.argtypes = [vp]
        avutil.av_frame_alloc.restype = vp
        avutil.av_frame_unref.argtypes = [vp]
        avutil.av_frame_free.argtypes = [vp]
        avutil.av_log_set_level(16)                    # errors only
        if not avcodec.avcodec_find_decoder_by_name(b"mp3float"):
            continue
        _av = (avutil, avcodec)
        return _av
    raise Mp3Unavailable("no libavcodec with an mp3float decoder found (set DYS_FFMPEG_LIBDIR)")


def available() -> bool:
    try:
        _load_av()
        return True
    except Mp3Unavailable:
        return False


def _decode_frames(blob: bytes, frames) -> np.ndarray:
    avutil, avcodec = _load_av()
    codec = avcodec.avcodec_find_decoder_by_name(b"mp3float")
    ctx = avcodec.avcodec_alloc_context3(codec)
    if not ctx or avcodec.avcodec_open2(ctx, codec, None) != 0:
        raise Mp3Unavailable("avcodec_open2(mp3float) failed")
    pkt, frm = avcodec.av_packet_alloc(), avutil.av_frame_alloc()
    spf = 1152 if frames and frames[0][2] == 3 else 576
    out = []
    try:
        for (pos, size, *_rest) in frames:
            if avcodec.av_new_packet(pkt, size) != 0:
                raise MemoryError("av_new_packet")
            C.memmove(C.c_void_p.from_address(pkt + _PKT_DATA).value, blob[pos:pos + size], size)
            rc = avcodec.avcodec_send_packet(ctx, pkt)
            avcodec.av_packet_unref(pkt)
            if rc != 0:                                    # undecodable frame: libmpg123 emits silence for it
                out.append(np.zeros(spf, np.float32))
                continue
            while avcodec.avcodec_receive_frame(ctx, frm) == 0:
                ns = C.c_int.from_address(frm + _FRM_NB_SAMPLES).value
                fmt = C.c_int.from_address(frm + _FRM_FORMAT).value
                if fmt not in (3, 8):                      # AV_SAMPLE_FMT_FLT / FLTP (mono: same layout)
                    raise Mp3Unavailable(f"unexpected sample format {fmt}")
                p0 = C.c_void_p.from_address(frm).value
                out.append(np.ctypeslib.as_array(C.cast(p0, C.POINTER(C.c_float)), (ns,)).copy())
                avutil.av_frame_unref(frm)
    finally:
        pp, fp, cp = C.c_void_p(pkt), C.c_void_p(frm), C.c_void_p(ctx)
        avcodec.av_packet_free(C.byref(pp))
        avutil.av_frame_free(C.byref(fp))
        avcodec.avcodec_free_context(C.byref(cp))
    return np.concatenate(out) if out else np.zeros(0, np.float32)


def decode_mp3(blob: bytes):
    """MP3 bytes -> (float32[n] mono samples at the file's own rate, sample rate), trimmed like libmpg123/libsndfile."""
    frames = parse_frames(blob)
    if not frames:
        raise ValueError("no MPEG Layer III frames found")
    if any(f[5] != 3 for f in frames):
        raise ValueError("only mono streams are supported (the reference corpus is mono)")
    tag = info_tag(blob, frames[0])
    y = _decode_frames(blob, frames[1:] if tag else frames)
    if tag:
        n = y.shape[0]
        begin = min(n, tag["delay"] + DECODER_DELAY)
        end = max(begin, min(n, n - tag["padding"] + DECODER_DELAY))
        y = y[begin:end]
    return np.ascontiguousarray(y, dtype=np.float32), frames[0][4]


def read_mp3(path: str):
    with open(path, "rb") as fh:
        return decode_mp3(fh.read())


def decoded_length(blob: bytes) -> int:
    """Number of samples ``decode_mp3`` returns, from the headers alone (no decoder needed)."""
    frames = parse_frames(blob)
    if not frames:
        return 0
    tag = info_tag(blob, frames[0])
    spf = 1152 if frames[0][2] == 3 else 576
    if not tag:
        return len(frames) * spf
    n = (len(frames) - 1) * spf
    begin = min(n, tag["delay"] + DECODER_DELAY)
    return max(begin, min(n, n - tag["padding"] + DECODER_DELAY)) - begin
