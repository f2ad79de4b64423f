#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end feature extraction feeding the reference's dysfluency classifier.

    python tools/config5_e2e.py [--clips 1000000] [--chunk 50000]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/config5_e2e.py --clips 1000000

Clips are generated ON THE DEVICE by seed (a torch restatement of the synthetic generator's recipe: harmonic voiced
source x raised-cosine syllable gate + white noise, peak 0.5 -- input plumbing, not part of the measured path), each
rank takes a contiguous share, extracts raw + clean vectors chunk by chunk, fits the global scaler over ALL ranks'
rows (the path's one all-reduce), standardises on the device and hands the host matrix to the reference's classifier:
RandomForest(200, random_state=42) trained on the reference's own cached features (tests/golden fixtures), as
main1.py:952-999 does at inference time.  Prints one JSON line with the per-stage times.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_SAMPLES, SR = 48000, 16000


def synth_on_device(first: int, count: int, dev) -> torch.Tensor:
    """float32 [count, 48000] on `dev`; clip i depends only on (1234 + first + i)."""
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + first)
    t = torch.arange(N_SAMPLES, device=dev, dtype=torch.float32) / SR
    f0 = 90.0 + 210.0 * torch.rand((count, 1), generator=g, device=dev)
    y = torch.zeros((count, N_SAMPLES), device=dev)
    for k in range(1, 9):
        ph = 2 * math.pi * torch.rand((count, 1), generator=g, device=dev)
        y += torch.sin(2 * math.pi * k * f0 * t + ph) / k
    rate = 2.0 + 4.0 * torch.rand((count, 1), generator=g, device=dev)
    u = torch.remainder(rate * t + torch.rand((count, 1), generator=g, device=dev), 1.0)
    y *= torch.where(u < 0.6, 0.5 - 0.5 * torch.cos(2 * math.pi * u / 0.6), torch.zeros_like(u))
    y += 0.01 * torch.randn((count, N_SAMPLES), generator=g, device=dev)
    y *= 0.5 / y.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
    return y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=25_000)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch.distributed as dist
    import dysb200 as pkg
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    fe = pkg.frontend
    lo, hi = pkg.sharding.shard_range(args.clips, rank, world)

    # the reference classifier, trained on the reference's cached vectors (golden fixtures)
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.preprocessing import LabelEncoder
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_scaler_after.npz"))
    c = np.load(os.path.join(ROOT, "tests", "golden", "ref_classifier_after.npz"))
    train_sc = pkg.scaler.GlobalScaler().fit(torch.from_numpy(g["X"]).to(dev))
    rf = RandomForestClassifier(n_estimators=200, random_state=42, n_jobs=-1)
    rf.fit(train_sc.transform(torch.from_numpy(g["X"]).to(dev)).cpu().numpy(), LabelEncoder().fit_transform(c["labels"]))

    fe.extract_features_batch(synth_on_device(0, 64, dev), denoise=True)        # warm-up (tables, arenas)
    torch.cuda.synchronize()
    t_gen = t_feat = 0.0
    raws, cleans = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for c0 in range(lo, hi, args.chunk):
        cnt = min(args.chunk, hi - c0)
        w0 = time.perf_counter()
        y = synth_on_device(c0, cnt, dev)
        torch.cuda.synchronize()
        t_gen += time.perf_counter() - w0
        e0.record()
        raw, clean = fe.extract_features_batch(y, denoise=True)
        e1.record()
        torch.cuda.synchronize()
        t_feat += e0.elapsed_time(e1) / 1e3
        raws.append(raw)
        cleans.append(clean)
        del y
    clean = torch.cat(cleans)
    w0 = time.perf_counter()
    Z, scaler = pkg.scaler.classifier_inputs(clean, gather=False)               # global CMVN: one all-reduce per pass
    torch.cuda.synchronize()
    t_cmvn = time.perf_counter() - w0
    w0 = time.perf_counter()
    Zc = train_sc.to_sklearn().transform(clean.cpu().numpy())                   # main1.py:987-989: the training scaler
    pred_par = rf.predict(Zc)                                                   # n_jobs=-1 like the reference (main1.py:862)
    t_rf = time.perf_counter() - w0
    # scikit-learn adds the trees' probabilities in thread-completion order: with n_jobs=-1 near-ties between two classes flip
    # from run to run ON THE SAME MATRIX (that, not the features, is why round 1's 1-GPU and 8-GPU class counts differed by
    # ~60 in 1e6).  The counts reported for the 1-GPU / 8-GPU comparison use the single-thread, fixed-order sum.
    par_counts = [np.bincount(pred_par, minlength=3).tolist(), np.bincount(rf.predict(Zc), minlength=3).tolist()]
    rf.set_params(n_jobs=1)
    pred = rf.predict(Zc)
    counts = np.bincount(pred, minlength=3).astype(np.float64)
    stats = torch.tensor([t_feat, t_gen, t_cmvn, t_rf, float(hi - lo)], dtype=torch.float64, device=dev)
    cls = torch.from_numpy(counts).to(dev)
    # order-independent fingerprints of the BITS of every vector (and of the generated clips): the 1-GPU and the 8-GPU run of
    # one commit must print the same numbers -- a clip's vectors depend on nothing but the clip
    raw = torch.cat(raws)
    sums = torch.stack([raw.view(torch.int32).to(torch.int64).sum(), clean.view(torch.int32).to(torch.int64).sum(),
                        (raw.view(torch.int32).to(torch.int64) * 2654435761 % 1000003).sum(),
                        (clean.view(torch.int32).to(torch.int64) * 2654435761 % 1000003).sum()])
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats[4:], op=dist.ReduceOp.SUM)
        dist.all_reduce(cls, op=dist.ReduceOp.SUM)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        stats[:4] = mx[:4]
    if rank == 0:
        t_feat, t_gen, t_cmvn, t_rf, total = stats.tolist()
        line = {"workload": "configs[4]: end-to-end feature extraction (raw + clean) feeding the reference classifier",
                "n_gpus": world, "clips": int(total), "audio_seconds": total * 3.0,
                "feature_extraction_s": t_feat, "audio_sec_per_sec_features": total * 3.0 / t_feat,
                "on_device_generation_s": t_gen, "global_cmvn_fit_and_apply_s": t_cmvn,
                "classifier_predict_s_cpu": t_rf, "predicted_class_counts": cls.tolist(),
                "rank0_counts_of_two_parallel_predicts_of_the_same_rows": par_counts,
                "scaler_mean_first3": scaler.mean_[:3].tolist(),
                "feature_bits_fingerprint": {"raw_sum": int(sums[0]), "clean_sum": int(sums[1]), "raw_hash": int(sums[2]),
                                             "clean_hash": int(sums[3])},
                "note": "times are max over ranks; the classifier is the reference's scikit-learn RandomForest on the host"}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
