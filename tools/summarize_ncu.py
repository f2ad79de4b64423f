"""Summaries of ncu captures for profiles/ (run in the build container; needs no GPU).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches_summary.csv
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep [...]  > profiles/rNN_ncu_full_summary.csv
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def kname(s):
    m = re.search(r"\b(k_\w+)", s)
    return m.group(1) if m else s.split("(")[0][-40:]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {n: j for j, n in enumerate(rows[h])}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
        n = kname(r[idx["Kernel Name"]])
        tot[n] += v
        cnt[n] += 1
    s = sum(tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised times: compare SHARES")
    print("kernel,launches,total_ms,avg_ms,share")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k},{cnt[k]},{v:.3f},{v / cnt[k]:.4f},{v / s:.4f}")


SCALE = {"ns": ("ms", 1e-6), "us": ("ms", 1e-3), "ms": ("ms", 1.0), "s": ("ms", 1e3), "second": ("ms", 1e3),
         "byte": ("Mbyte", 1e-6), "Kbyte": ("Mbyte", 1e-3), "Mbyte": ("Mbyte", 1.0), "Gbyte": ("Mbyte", 1e3),
         "Tbyte": ("Mbyte", 1e6)}


def full(paths):
    """One row per profiled launch; times normalised to ms and byte counts to Mbyte (ncu picks a unit per report)."""
    print("# ncu --set full --clock-control none; one row per profiled launch")
    first = True
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        idx = {n: j for j, n in enumerate(rows[0])}
        cols = [m for m in KEY_METRICS if m in idx]
        units = {c: rows[1][idx[c]] for c in cols}
        if first:
            print("kernel," + ",".join(f"{c} [{SCALE.get(units[c], (units[c], 1.0))[0]}]" for c in cols))
            first = False
        for r in rows[2:]:
            vals = []
            for c in cols:
                v = r[idx[c]].replace(",", "")
                if units[c] in SCALE and c not in ("launch__shared_mem_per_block_dynamic",):
                    try:
                        v = f"{float(v) * SCALE[units[c]][1]:.6g}"
                    except ValueError:
                        pass
                vals.append(v)
            print(kname(r[idx["Kernel Name"]]) + "," + ",".join(vals))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
