"""First-light GPU check: stage-wise errors of the CUDA path against the oracle (prints, never asserts)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dysb200  # noqa: E402
from oracle import denoise as D  # noqa: E402
from oracle import features as F  # noqa: E402

fe = dysb200.frontend
syn = dysb200.synth


def stage_report(name, y):
    it = F.intermediates(y)
    g = fe.debug_feature_stages(y)
    P, Pg = it["power"], g["power"]
    rel = np.abs(Pg - P).max() / max(P.max(), 1e-30)
    Lu = 10 * np.log10(np.maximum(1e-10, it["mel"]))
    print(f"[{name}] T={g['frames']} status={g['status']} power rel-to-max err {rel:.3e}; "
          f"logmel(unclamped) err {np.abs(g['logmel_unclamped'] - Lu).max():.3e}; "
          f"mfcc err {np.abs(g['mfcc'] - it['mfcc']).max():.3e}; tuning gpu {g['tuning_index']} "
          f"oracle {F.tuning_index(it['tuning'])} peaks {g['peak_count']} (oracle {len(F.piptrack_peaks(P)[0])}); "
          f"chroma err {np.abs(g['chroma'] - it['chroma']).max():.3e}")
    ref = F.extract_features(y)
    d = np.abs(g["features"] - ref)
    print(f"      features: mfcc {d[:40].max():.3e} delta {d[40:80].max():.3e} delta2 {d[80:120].max():.3e} "
          f"chroma {d[120:144].max():.3e}")


def main():
    print(torch.cuda.get_device_name(0))
    for i in range(3):
        stage_report(f"synth{i}", syn.synth_clip(i))
    for name, y in syn.edge_clips():
        if len(y) >= 4096:
            try:
                stage_report(name, y)
            except Exception as e:
                print(name, "stage_report failed:", e)
    # denoise
    for i in range(2):
        y = syn.synth_clip(i)
        ref = D.reduce_noise(y)
        got, peak, flag = fe.debug_denoise(y)
        print(f"[denoise synth{i}] max abs err {np.abs(got - ref).max():.3e} (peak ref {np.abs(ref).max():.6f} gpu {peak:.6f} flag {flag}) "
              f"n_diff_f32 {(got != ref).sum()}")
        qg = np.clip(np.rint(got / np.float32(peak) * np.float32(32768)), -32768, 32767)
        qr = D.clean_audio(y)
        print(f"      pcm16 flips {(qg != qr).sum()} of {len(qr)}")
    # batch raw+clean vs oracle
    X = syn.synth_batch(8)
    raw, clean, status = fe.extract_features_batch(X, denoise=True, return_status=True)
    raw, clean = raw.cpu().numpy(), clean.cpu().numpy()
    for i in range(8):
        r = F.extract_features(X[i]); c = F.extract_features(D.clean_then_load(X[i]))
        print(f"clip {i}: raw err {np.abs(raw[i]-r).max():.3e} clean err {np.abs(clean[i]-c).max():.3e} "
              f"(mfcc {np.abs(clean[i]-c)[:120].max():.3e} chroma {np.abs(clean[i]-c)[120:144].max():.3e})")
    print("status", status.cpu().numpy())
    # timing
    B = 2048
    Xb = torch.from_numpy(np.tile(X, (B // 8, 1))).cuda()
    for denoise in (False, True):
        fe.extract_features_batch(Xb, denoise=denoise)
        torch.cuda.synchronize()
        t0 = time.time()
        fe.extract_features_batch(Xb, denoise=denoise)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"denoise={denoise}: {B} clips in {dt*1e3:.1f} ms -> {B/dt:.0f} clips/s = {B*3/dt:.3e} x realtime")


if __name__ == "__main__":
    main()
