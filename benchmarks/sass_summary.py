#!/usr/bin/env python
"""profiles/<tag>_sass_summary.md: per-kernel counts of the SASS mnemonics that tell how a kernel moves data and computes
(cuobjdump -sass of the built library; no GPU needed).

    python benchmarks/sass_summary.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smow_net_b200", "libsmow_b200.so")
# mnemonic prefix -> what it proves (B200_PROFILING.md "What proves a Blackwell-native kernel")
COLS = [("UTCHMMA", "tcgen05.mma (5th-gen tensor cores)"), ("LDTM", "tcgen05.ld (TMEM -> registers)"),
        ("UTMALDG", "TMA tensor-map load"), ("UTMASTG", "TMA tensor-map store"), ("UBLKCP", "1-D bulk copy (cp.async.bulk)"),
        ("UBLKPF", "bulk L2 prefetch"), ("LDGSTS", "cp.async (16-byte async copy)"), ("SYNCS", "mbarrier ops"),
        ("FFMA2", "packed fp32x2 FMA"), ("FFMA", "fp32 FMA (incl. FFMA2)"), ("HMMA", "warp-level mma.sync (the 8-token-wide tokenizer GEMMs, 3xTF32)"),
        ("REDG", "global reductions (red.global.add)"), ("ATOMG", "global atomics"), ("LDG", "global loads"), ("STG", "global stores"),
        ("LDS", "shared loads"), ("STS", "shared stores"), ("SHFL", "warp shuffles")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("smow::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for pre, _ in COLS:
                if op.startswith(pre):
                    cur[pre] += 1
    out = os.path.join(ROOT, "profiles", tag + "_sass_summary.md")
    with open(out, "w") as f:
        f.write("# %s — SASS evidence per kernel (`cuobjdump -sass smow_net_b200/libsmow_b200.so`, sm_100a)\n\n" % tag)
        f.write("Counts of instruction mnemonics (static, per kernel instantiation). " + "; ".join("`%s` = %s" % c for c in COLS) + ".\n\n")
        f.write("| kernel | instr | " + " | ".join(c for c, _ in COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
        tot = collections.Counter()
        for name, c in kernels.items():
            f.write("| %s | %d | %s |\n" % (name[:70], c["_total"], " | ".join(str(c[p]) if c[p] else "" for p, _ in COLS)))
            tot.update(c)
        f.write("| **all %d kernels** | %d | %s |\n" % (len(kernels), tot["_total"], " | ".join(str(tot[p]) for p, _ in COLS)))
        tc = [n for n, c in kernels.items() if c["UTCHMMA"]]
        f.write("\nKernels with tcgen05.mma (`UTCHMMA`) + TMEM loads (`LDTM`) + TMA (`UTMALDG` / `UTMASTG`): %s.\n" % ", ".join("`%s`" % n for n in tc))
        hm = [n for n, c in kernels.items() if c["HMMA"]]
        f.write("\nKernels with warp-level TF32 MMAs (`HMMA.1688.F32.TF32`; N = 8 tokens is below a tcgen05 tile): %s.\n" % ", ".join("`%s`" % n for n in hm))
    print(out)


if __name__ == "__main__":
    main()
