"""Single-clip latency of the reference-named calls (the sidebar inference path, main1.py:952-999)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dysb200 as pkg
fe = pkg.frontend
y = pkg.synth.synth_clip(3)
def wall(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("extract_features(y)                 %.3f ms" % wall(lambda: fe.extract_features(y, 16000)))
print("clean_audio(y) -> PCM               %.3f ms" % wall(lambda: fe.clean_audio(y)))
print("raw + clean vectors of one clip     %.3f ms" % wall(lambda: [t.cpu() for t in fe.extract_features_batch([y], denoise=True)]))
d = torch.from_numpy(y).cuda()[None]
print("  same, clip already on the device  %.3f ms" % wall(lambda: [t.cpu() for t in fe.extract_features_batch(d, denoise=True)]))
for b in (8, 64, 512):
    D = d.repeat(b, 1)
    print("batch of %3d device clips            %.3f ms" % (b, wall(lambda: [t.cpu() for t in fe.extract_features_batch(D, denoise=True)], 20)))
