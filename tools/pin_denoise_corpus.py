#!/usr/bin/env python
"""Pins the oracle's load + denoise chain against the reference's own artefacts (build container only).

For every stem of /root/reference:  segrigated_samples/**/<stem>.mp3  -> mp3io.decode_mp3 (libmpg123 conventions)
-> oracle.resample (soxr_hq restatement) -> (a) oracle.features vs cache_features/<stem>_raw_feats.npy,
(b) oracle.denoise.clean_audio vs clear_audio/<stem>.wav (SNR of the difference, LSB mismatch rate),
(c) oracle.features of that PCM vs cache_features/<stem>_clean_feats.npy.
With --perturb the spectral gate is run with one parameter off its noisereduce default, to show that the comparison
discriminates.  Writes profiles/r02_denoise_pin_corpus.json.

    python tools/pin_denoise_corpus.py [--limit N] [--perturb]
"""
import argparse
import glob
import importlib
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
PERTURBATIONS = {
    "default": (1.0, None),
    "n_grad_freq 16->8": (1.0, {"n_grad_freq": 8}),
    "n_grad_time 3->1": (1.0, {"n_grad_time": 1}),
    "thresh 2->1.5": (1.0, {"thresh": 1.5}),
    "slope 10->5": (1.0, {"slope": 5.0}),
    "time_constant 2->1 s": (1.0, {"time_constant_s": 1.0}),
    "prop_decrease 1.0->0.8": (0.8, None),
}


def corpus_files():
    files = sorted(glob.glob(f"{REF}/segrigated_samples/*/*.mp3"))
    seen, out = set(), []
    for f in files:                                   # the first file in sorted order owns the stem (pipeline1.py:134-135)
        stem = os.path.basename(f).rsplit(".", 1)[0]
        if stem in seen or not os.path.exists(f"{REF}/clear_audio/{stem}.wav"):
            continue
        seen.add(stem)
        out.append((stem, f))
    return out


def pcm_stats(q_ref, q):
    d = q.astype(np.float64) - q_ref.astype(np.float64)
    sig = float((q_ref.astype(np.float64) ** 2).sum())

    def snr(dd, s):
        return float(10 * np.log10(s / max(float((dd ** 2).sum()), 1e-9)))
    body = slice(0, max(1, len(d) - 64))
    return {"snr_db": snr(d, sig), "snr_db_without_last_64": snr(d[body], float((q_ref[body].astype(np.float64) ** 2).sum())),
            "lsb_mismatch": float(np.mean(d != 0)), "max_abs_lsb": int(np.abs(d).max()),
            "mean_abs_lsb": float(np.abs(d).mean())}


def _work(job):
    stem, path, mode = job
    os.environ["OMP_NUM_THREADS"] = "1"
    mp3io = importlib.import_module("recognizing-speech-dysfluencies-in-stuttering_b200.mp3io")
    from oracle import denoise, features, resample, wavio
    y22, sr = mp3io.read_mp3(path)
    y = resample.resample(y22, sr, 16000)
    q_ref, _ = wavio.read_wav_pcm16(f"{REF}/clear_audio/{stem}.wav")
    rec = {"stem": stem, "n22": int(len(y22)), "n16": int(len(y)), "n_ref": int(len(q_ref)), "sr": sr}
    if len(y) != len(q_ref):
        return rec
    if mode == "full":
        g_raw = np.load(f"{REF}/cache_features/{stem}_raw_feats.npy")
        g_clean = np.load(f"{REF}/cache_features/{stem}_clean_feats.npy")
        o = features.extract_features(y)
        e = np.abs(o - g_raw)
        rec["raw_feat_err"] = [float(e[:40].max()), float(e[40:80].max()), float(e[80:120].max()), float(e[120:144].max())]
        rec["raw_feat_pass_1e-3"] = bool(np.all(e[:120] <= 1e-3 + 1e-4 * np.abs(g_raw[:120])))
        q = denoise.clean_audio(y)
        rec.update(pcm_stats(q_ref, q))
        e = np.abs(features.extract_features(wavio.dequantize_pcm16(q)) - g_clean)
        rec["clean_feat_err"] = [float(e[:40].max()), float(e[40:80].max()), float(e[80:120].max()), float(e[120:144].max())]
    else:
        rec["perturb"] = {}
        for name, (prop, pt) in PERTURBATIONS.items():
            rec["perturb"][name] = pcm_stats(q_ref, denoise.clean_audio(y, prop, pt))["snr_db_without_last_64"]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--perturb", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_denoise_pin_corpus.json"))
    args = ap.parse_args()
    files = corpus_files()
    if args.perturb:
        files = files[::max(1, len(files) // 120)]
    if args.limit:
        files = files[:: max(1, len(files) // args.limit)]
    mode = "perturb" if args.perturb else "full"
    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
        recs = pool.map(_work, [(s, f, mode) for s, f in files], chunksize=2)
    summary = {"clips": len(recs), "length_matches": sum(r["n16"] == r["n_ref"] for r in recs)}
    if mode == "full":
        ok = [r for r in recs if "snr_db" in r]
        for key in ("snr_db", "snr_db_without_last_64", "lsb_mismatch", "max_abs_lsb", "mean_abs_lsb"):
            v = np.array([r[key] for r in ok], dtype=np.float64)
            summary[key] = {"min": float(v.min()), "p05": float(np.percentile(v, 5)), "median": float(np.median(v)),
                            "p95": float(np.percentile(v, 95)), "max": float(v.max())}
        for key in ("raw_feat_err", "clean_feat_err"):
            v = np.array([r[key] for r in ok])
            summary[key] = {"groups": ["mfcc", "delta", "delta2", "chroma"], "median": np.median(v, 0).tolist(),
                            "p95": np.percentile(v, 95, axis=0).tolist(), "max": v.max(0).tolist()}
        summary["raw_feat_pass_1e-3"] = sum(r["raw_feat_pass_1e-3"] for r in ok)
    else:
        names = list(PERTURBATIONS)
        tab = np.array([[r["perturb"][k] for k in names] for r in recs if "perturb" in r])
        summary["perturbation_snr_db_median"] = dict(zip(names, np.median(tab, 0).tolist()))
        summary["perturbation_snr_db_p05"] = dict(zip(names, np.percentile(tab, 5, axis=0).tolist()))
        summary["perturbation_snr_db_p95"] = dict(zip(names, np.percentile(tab, 95, axis=0).tolist()))
        summary["clips_where_default_is_best"] = int(np.sum(np.argmax(tab, 1) == 0))
    out = args.out if mode == "full" else args.out.replace(".json", "_perturb.json")
    json.dump({"summary": summary, "clips": recs}, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
