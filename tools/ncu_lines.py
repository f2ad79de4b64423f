"""Per-source-line stall samples of one kernel from an ncu report captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_frame_spectra [top_n]

Prints the lines with the most warp-stall samples, the executed warp-instructions attributed to
them and the dominant stall reasons (needs -lineinfo at compile time; runs without a GPU).
"""
import collections
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k",
                          f"regex:{kernel}", "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    names = rows[hdr]
    col = {n: i for i, n in enumerate(names)}
    stall_cols = [(n, i) for i, n in enumerate(names) if n.startswith("stall_") and "Not Issued" not in n]
    samp = col["# Samples"]
    inst = col["Instructions Executed"]
    per_line = collections.OrderedDict()
    cur = None
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) < len(names) or r[0] == "Line No":
            continue
        if r[0] != "":                       # a CUDA source line header: its numbers are the sum of its SASS rows
            cur = (f"{fname}:{r[0]}", r[1].strip()[:90])
            d = per_line.setdefault(cur, {"samples": 0, "inst": 0, "stalls": collections.Counter()})
            d["samples"] += num(r[samp])
            d["inst"] += num(r[inst])
            for n, i in stall_cols:
                d["stalls"][n] += num(r[i])
    tot_s = sum(d["samples"] for d in per_line.values()) or 1
    tot_i = sum(d["inst"] for d in per_line.values()) or 1
    allst = collections.Counter()
    for d in per_line.values():
        allst.update(d["stalls"])
    print(f"# {kernel}: {tot_s} stall samples, {tot_i} warp-instructions")
    print("# stall mix:", ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in allst.most_common(8)))
    print("line,samples%,inst%,top stalls,source")
    for (ln, src), d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = " ".join(f"{k[6:]}:{v}" for k, v in d["stalls"].most_common(3))
        print(f"{ln},{100 * d['samples'] / tot_s:.1f},{100 * d['inst'] / tot_i:.1f},{st},{src}")


if __name__ == "__main__":
    main()
