#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped libdysb200.so (sm_100a cubins): what the B200 actually executes.

    python tools/sass_summary.py [lib] > profiles/r02_sass_summary.txt

Columns: registers / static shared memory / local (stack) bytes from `cuobjdump -res-usage`, then counts of the
opcodes that matter for this path -- packed fp32 (FADD2/FMUL2/FFMA2, Blackwell), fp64 FMA, shared / global / local
memory traffic, bulk-copy engine (UBLKCP = cp.async.bulk, UTMALDG/UTMASTG = cp.async.bulk.tensor), mbarrier (SYNCS),
shuffles, MUFU, and tensor-core opcodes (UTC*MMA / HMMA / DMMA: expected 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "recognizing-speech-dysfluencies-in-stuttering_b200", "libdysb200.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
GROUPS = [("FADD2", r"^FADD2"), ("FMUL2", r"^FMUL2"), ("FFMA2", r"^FFMA2"), ("FFMA", r"^FFMA$|^FFMA\."), ("FADD", r"^FADD$|^FADD\."),
          ("FMUL", r"^FMUL$|^FMUL\."), ("DFMA", r"^DFMA"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("MUFU", r"^MUFU"),
          ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDL", r"^LDL"), ("STL", r"^STL"),
          ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("UBLKCP", r"^UBLKCP"), ("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"),
          ("SYNCS", r"^SYNCS"), ("LDGSTS", r"^LDGSTS"), ("CCTL/PREFETCH", r"^CCTL"), ("TENSOR", r"^UTC.*MMA|^HMMA|^DMMA|^IMMA|^QGMMA|^LDTM|^STTM")]


_names = {}


def demangle(name: str) -> str:
    if name not in _names:
        try:
            full = subprocess.run(["/usr/local/cuda/bin/cu++filt", name], capture_output=True, text=True).stdout.strip()
        except OSError:
            full = name
        m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?\(", full)
        _names[name] = (m.group(1) + (m.group(2) or "")) if m else name
    return _names[name]


def main():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", ln))
            cur = None
    counts, total = collections.OrderedDict(), {}
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if cur and m:
            op = m.group(1)
            total[cur] += 1
            for g, pat in GROUPS:
                if re.match(pat, op):
                    counts[cur][g] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}  (cuobjdump -sass / -res-usage; static instruction counts, not executed counts)")
    hdr = ["kernel", "instr", "REG", "SHARED", "STACK"] + [g for g, _ in GROUPS]
    print("\t".join(hdr))
    for fn in sorted(counts, key=demangle):
        u = usage.get(fn, {})
        row = [demangle(fn), str(total[fn]), u.get("REG", "?"), u.get("SHARED", "?"), u.get("STACK", "?")]
        row += [str(counts[fn][g]) for g, _ in GROUPS]
        print("\t".join(row))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\t".join(["ALL", str(sum(total.values())), "", "", ""] + [str(tot[g]) for g, _ in GROUPS]))


if __name__ == "__main__":
    main()
