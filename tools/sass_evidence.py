#!/usr/bin/env python3
"""Count the SASS mnemonics that prove what each kernel of the embedded sm_100a cubin uses (B200_PROFILING.md:
UTCHMMA / UTCQMMA / UTCIMMA = tcgen05.mma, UTMALDG = TMA tensor load, UTCCP = tcgen05.cp, LDTM = tcgen05.ld, SYNCS =
mbarrier, LDS/STS = shared memory) -- runs on the CPU box (cuobjdump), writes profiles/r01_sass_evidence.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUBIN = os.path.join(ROOT, "coast_b200", "csrc", "coast_kernels.cubin")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTCCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDS", "STS", "LDG",
        "STG", "LDL", "STL", "SHFL", "VOTE", "REDUX", "PRMT", "LOP3", "SHF", "IADD3", "IMAD", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", CUBIN], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    want = sys.argv[1:] or None
    lines = ["# SASS evidence (r01) -- `python tools/sass_evidence.py`, from coast_b200/csrc/coast_kernels.cubin (sm_100a)", "",
             "Static instruction counts per kernel (not dynamic).  tcgen05.mma = `UTC*MMA`, tcgen05.cp = `UTCCP`, tcgen05.ld = `LDTM`,",
             "TMA tensor load = `UTMALDG`, mbarrier = `SYNCS`.  Only non-zero columns of interest are listed.", "",
             "| kernel | total | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for k, c in per.items():
        if want and not any(w in k for w in want):
            continue
        if k.endswith("_inj1") and not want:
            continue                                          # the injector variants differ only by the fault hooks
        lines.append(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
    text = "\n".join(lines) + "\n"
    dst = os.path.join(ROOT, "profiles", "r01_sass_evidence.md")
    with open(dst, "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    main()
