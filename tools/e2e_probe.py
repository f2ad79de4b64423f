"""Timeline probe of the host streaming path: when does each chunk's copy land and when is its compute done?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dysb200 as pkg
fe = pkg.frontend
N, L = 10000, 48000
base = torch.from_numpy(pkg.synth.synth_batch(100))
host = base.repeat(N // 100, 1).contiguous().pin_memory()
dev = torch.device("cuda", 0)
out_raw = torch.empty((N, 149)).pin_memory(); out_clean = torch.empty((N, 149)).pin_memory()
staging = torch.empty((N * L,), dtype=torch.float32, device=dev)
copy_s, comp_s, comp_t = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
NSTREAMS = 1

def run(sizes, trace=False):
    cur = torch.cuda.current_stream()
    t0 = torch.cuda.Event(enable_timing=True); t0.record(cur)
    copy_s.wait_stream(cur); comp_s.wait_stream(cur)
    landed, done, c0 = [], [], 0
    with torch.cuda.stream(copy_s):
        for cnt in sizes:
            staging[c0 * L:(c0 + cnt) * L].copy_(host[c0:c0 + cnt].reshape(-1), non_blocking=True)
            ev = torch.cuda.Event(enable_timing=True); ev.record(copy_s); landed.append(ev); c0 += cnt
    c0 = 0
    big = max(sizes)
    starts = torch.arange(big, dtype=torch.int64, device=dev) * L
    lens = torch.full((big,), L, dtype=torch.int32, device=dev)
    comp_t.wait_stream(cur)
    for i, (cnt, ev) in enumerate(zip(sizes, landed)):
        st = (comp_s, comp_t)[i % NSTREAMS]
        with torch.cuda.stream(st):
            st.wait_event(ev)
            raw, clean, _, _ = fe._run_device(staging[c0 * L:(c0 + cnt) * L], starts[:cnt], lens[:cnt], L, True, 1.0, False, None, 1, slot=41 + i % NSTREAMS)
            out_raw[c0:c0 + cnt].copy_(raw, non_blocking=True); out_clean[c0:c0 + cnt].copy_(clean, non_blocking=True)
            e = torch.cuda.Event(enable_timing=True); e.record(st); done.append(e); c0 += cnt
    cur.wait_stream(comp_s); cur.wait_stream(comp_t); cur.synchronize()
    if trace:
        for cnt, a, b in zip(sizes, landed, done):
            print("   chunk %5d  copy landed %6.2f ms   compute done %6.2f ms" % (cnt, t0.elapsed_time(a), t0.elapsed_time(b)))

def wall(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

for ns in (1, 2):
    NSTREAMS = ns
    for name, sizes in (("625s", [156, 312] + [625] * 15 + [157]), ("1000s", [250, 500] + [1000] * 9 + [250]),
                        ("1250s", [156, 312, 625] + [1250] * 7 + [157])):
        assert sum(sizes) == N, sum(sizes)
        print("streams", ns, name, "%.2f ms" % wall(lambda: run(sizes)))
run([156, 312] + [625] * 15 + [157], trace=True)
