#!/usr/bin/env python
"""Blackwell-native evidence from the built library: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM /
TMA (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG) plus an
excerpt of the MMA issue loop of the main GEMM kernel.  Runs here (no GPU):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smirk_b200", "libsmirk_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "FFMA2", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    excerpt = []
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"smk::\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"\b%s\b" % p, ln) or (p in ("UTMALDG", "UTCHMMA", "LDTM", "STTM") and p in ln):
                per[cur][p] += 1
                if ".IM2COL" in ln:
                    per[cur]["UTMALDG.IM2COL"] += 1
                if p == "UTCHMMA" and re.search(r"UTCHMMA\s+tmem\[", ln):
                    per[cur]["UTCHMMA(A in TMEM)"] += 1
                if p == "UTCHMMA" and "gemm_tc_kernel<256" in cur and len(excerpt) < 14:
                    excerpt.append(ln.strip())
    tot = collections.Counter()
    print("# SASS mnemonic counts per kernel — %s (cuobjdump -sass)" % os.path.relpath(LIB, ROOT))
    for k, c in per.items():
        if c:
            tot.update(c)
            print("%-110s %s" % (k[:110], "  ".join("%s=%d" % (a, b) for a, b in sorted(c.items()))))
    print("\n# totals: " + "  ".join("%s=%d" % (a, b) for a, b in sorted(tot.items())))
    print("\n# excerpt: tcgen05.mma issue in gemm_tc_kernel<256,...> (128 x 256 x 8 TF32 per instruction)")
    for ln in excerpt:
        print("    " + ln)


if __name__ == "__main__":
    main()
