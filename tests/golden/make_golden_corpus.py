"""Builds the whole-corpus golden fixtures from the reference's own artefacts (build container only):

    python tests/golden/make_golden_corpus.py

  tests/golden/ref_corpus_wav.npz -- every committed ``clear_audio/<stem>.wav`` (888; the reference's denoise OUTPUT and
      the clean feature function's INPUT) with the committed ``cache_features/<stem>_clean_feats.npy``
  tests/golden/ref_corpus_mp3.npz -- the ``segrigated_samples/**/<stem>.mp3`` file that produced each of them (the first
      one in sorted-path order: clean_audio_and_cache skips existing WAVs, pipeline1.py:134-135) with the committed
      ``cache_features/<stem>_raw_feats.npy`` and the label directory
Only data is copied, never reference source code.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from oracle import wavio  # noqa: E402
import corpus  # noqa: E402

REF = "/root/reference"


def main():
    files = sorted(glob.glob(f"{REF}/segrigated_samples/*/*.mp3"))
    seen, rows = set(), []
    for f in files:
        stem = os.path.basename(f).rsplit(".", 1)[0]
        if stem in seen or not os.path.exists(f"{REF}/clear_audio/{stem}.wav"):
            continue
        seen.add(stem)
        rows.append((stem, f))
    names = np.asarray([r[0] for r in rows])
    pcm, offs = [], [0]
    for stem, _ in rows:
        q, sr = wavio.read_wav_pcm16(f"{REF}/clear_audio/{stem}.wav")
        assert sr == 16000
        pcm.append(q)
        offs.append(offs[-1] + len(q))
    pcm = np.concatenate(pcm)
    packed = corpus.pack_pcm(pcm)
    assert np.array_equal(corpus.unpack_pcm(packed), pcm)
    np.savez(corpus.WAV_NPZ, names=names, pcm_packed=packed, offsets=np.asarray(offs, np.int64),
             clean_feats=np.stack([np.load(f"{REF}/cache_features/{s}_clean_feats.npy") for s, _ in rows]))
    print("wav corpus:", len(rows), "clips,", offs[-1], "samples,", os.path.getsize(corpus.WAV_NPZ) / 1e6, "MB")
    blobs = [open(f, "rb").read() for _, f in rows]
    moffs = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype(np.int64)
    np.savez(corpus.MP3_NPZ, names=names, mp3_bytes=np.frombuffer(b"".join(blobs), dtype=np.uint8), offsets=moffs,
             raw_feats=np.stack([np.load(f"{REF}/cache_features/{s}_raw_feats.npy") for s, _ in rows]),
             labels=np.asarray([os.path.basename(os.path.dirname(f)) for _, f in rows]))
    print("mp3 corpus:", len(rows), "files,", os.path.getsize(corpus.MP3_NPZ) / 1e6, "MB")


if __name__ == "__main__":
    main()
