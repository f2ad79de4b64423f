"""Builds the committed golden fixtures from the reference's own artefacts.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Outputs (committed; /root/reference does not exist on the GPU box):
  tests/golden/ref_clean_pairs.npz  -- a spread of the reference's
      ``clear_audio/<stem>.wav`` (int16 PCM) with the matching committed
      ``cache_features/<stem>_clean_feats.npy`` (float32[149]); the feature function's
      golden input/output pairs.
  tests/golden/ref_scaler_after.npz -- X_after (the 905x149 matrix the reference fed to
      ``StandardScaler().fit`` at pipeline1.py:471, rebuilt from cache_features/ in
      sorted-path order, duplicates included) with mean_/var_/scale_ unpickled from
      output_results/scaler_after.pkl.
  tests/golden/ref_classifier_after.npz -- the class label of each of those 905 rows and the accuracy /
      log-loss the reference published for the cleaned features (output_results/metrics_summary.csv):
      the end-to-end golden of the classifier input loader.
  tests/golden/ref_file_bytes.npz -- the bytes of the smallest committed clear_audio/<stem>.wav and of its
      cache_features/<stem>_clean_feats.npy: goldens of the two on-disk formats.
  tests/golden/ref_qc_after.npz -- snr_after / flat_after / hf_after of output_results/per_file_analysis.csv for
      the stems of ref_clean_pairs.npz: goldens of the three per-file QC scalars.
Only data is copied, never reference source code.
"""
import glob
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import wavio  # noqa: E402

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    feats = sorted(glob.glob(f"{REF}/cache_features/*_clean_feats.npy"))
    stems = [os.path.basename(f)[:-len("_clean_feats.npy")] for f in feats]
    lens = []
    for s in stems:
        q, sr = wavio.read_wav_pcm16(f"{REF}/clear_audio/{s}.wav")
        assert sr == 16000
        lens.append(len(q))
    order = np.argsort(lens)
    # shortest, longest and an even spread in between, total audio kept small (~1.7 MB)
    picks = sorted(set([order[0], order[1], order[-1], order[-2]] +
                       [order[int(round(i))] for i in np.linspace(0, len(order) - 1, 22)]))
    pcm, offs, gold, names = [], [0], [], []
    for i in picks:
        q, _ = wavio.read_wav_pcm16(f"{REF}/clear_audio/{stems[i]}.wav")
        pcm.append(q)
        offs.append(offs[-1] + len(q))
        gold.append(np.load(feats[i]))
        names.append(stems[i])
    np.savez_compressed(os.path.join(HERE, "ref_clean_pairs.npz"), pcm=np.concatenate(pcm),
                        offsets=np.asarray(offs, dtype=np.int64), feats=np.stack(gold),
                        names=np.asarray(names))
    print("pairs:", len(picks), "samples:", offs[-1], "lens:", [lens[i] for i in picks])

    # QC golden: the *_after columns of per_file_analysis.csv (pipeline1.py:394-396, computed on the cleaned WAV) for
    # the picked stems, in the order of ref_clean_pairs.npz
    import csv as _csv
    by_stem = {}
    for r in _csv.DictReader(open(f"{REF}/output_results/per_file_analysis.csv")):
        by_stem.setdefault(os.path.splitext(r["file"])[0], r)
    np.savez_compressed(os.path.join(HERE, "ref_qc_after.npz"), names=np.asarray(names),
                        snr=np.asarray([float(by_stem[s]["snr_after"]) for s in names]),
                        flat=np.asarray([float(by_stem[s]["flat_after"]) for s in names]),
                        hf=np.asarray([float(by_stem[s]["hf_after"]) for s in names]))
    print("qc golden rows:", len(names))

    # scaler golden: rows follow sorted(list_audio_files) (pipeline1.py:91-97), key = basename stem
    files = []
    for r, _, fs in os.walk(f"{REF}/segrigated_samples"):
        for f in fs:
            if f.lower().endswith((".wav", ".mp3", ".flac", ".m4a", ".ogg")):
                files.append(os.path.join(r, f))
    files = sorted(files)
    rows = [np.load(f"{REF}/cache_features/{os.path.basename(p).rsplit('.', 1)[0]}_clean_feats.npy") for p in files]
    X = np.vstack(rows)
    with open(f"{REF}/output_results/scaler_after.pkl", "rb") as fh:
        try:
            sc = pickle.load(fh)
        except Exception:
            import joblib
            sc = joblib.load(f"{REF}/output_results/scaler_after.pkl")
    np.savez_compressed(os.path.join(HERE, "ref_scaler_after.npz"), X=X, mean=sc.mean_, var=sc.var_,
                        scale=sc.scale_, n=np.int64(sc.n_samples_seen_))
    print("scaler rows:", X.shape, "n_samples_seen:", sc.n_samples_seen_)

    # classifier golden: labels of those rows (directory names, pipeline1.py:362-366) and the metrics the
    # reference published for the cleaned features (output_results/metrics_summary.csv, written at
    # pipeline1.py:533-535 after RandomForest(200, random_state=42) etc. on the stratified 80/20 split)
    import csv
    labels = np.asarray([os.path.basename(os.path.dirname(p)) for p in files])
    rows_m = list(csv.DictReader(open(f"{REF}/output_results/metrics_summary.csv")))
    after = {r["model"]: (float(r["accuracy"]), float(r["test_loss"])) for r in rows_m if r["dataset"] == "after"}
    np.savez_compressed(os.path.join(HERE, "ref_classifier_after.npz"), labels=labels,
                        models=np.asarray(list(after)), accuracy=np.asarray([after[m][0] for m in after]),
                        test_loss=np.asarray([after[m][1] for m in after]))
    print("classifier golden:", {m: after[m] for m in after})

    # on-disk format golden: the raw bytes of one committed clear_audio WAV (header + samples) and of one
    # committed cache_features .npy -- what soundfile.write / np.save produced at pipeline1.py:142 / :439
    small = min(feats, key=lambda f: os.path.getsize(f"{REF}/clear_audio/{os.path.basename(f)[:-len('_clean_feats.npy')]}.wav"))
    stem = os.path.basename(small)[:-len("_clean_feats.npy")]
    np.savez_compressed(os.path.join(HERE, "ref_file_bytes.npz"),
                        wav=np.frombuffer(open(f"{REF}/clear_audio/{stem}.wav", "rb").read(), dtype=np.uint8),
                        npy=np.frombuffer(open(small, "rb").read(), dtype=np.uint8), stem=np.asarray(stem))
    print("file-bytes golden:", stem)


if __name__ == "__main__":
    main()
