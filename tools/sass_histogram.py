#!/usr/bin/env python
"""tools/sass_histogram.py [OUT.md] -- SASS opcode histogram of every kernel in spectral_b200/lib/libspectral.so
(`cuobjdump -sass`, runs on the build box: no GPU needed).  Lists the mnemonics that prove what each kernel is built on:
UBLKCP / SYNCS (TMA 1-D bulk copy + mbarrier), VOTE / POPC (warp votes), SHFL, REDUX / CREDUX, DFMA / DADD / DMUL (FP64 pipe),
DMMA (FP64 tensor pipe), LDS / STS, LDL / STL (local memory), BAR, ATOM / RED."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectral_b200", "lib", "libspectral.so")
KEYS = ["UBLKCP", "SYNCS", "VOTE", "POPC", "SHFL", "REDUX", "CREDUX", "DFMA", "DADD", "DMUL", "DMMA", "DSETP", "LDS", "STS", "LDG", "STG",
        "LDL", "STL", "BAR", "ATOM", "ATOMG", "RED", "MUFU"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.split("\n")
    out = ["# SASS opcode histogram per kernel of libspectral.so (sm_100a), `python tools/sass_histogram.py`", "",
           "Static instruction counts (`cuobjdump -sass`).  cub:: kernels are the radix sort / scans of the shared-KKT tile builder.", "",
           "| kernel | total | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm.replace("(anonymous namespace)::", ""))[:60]
        out.append("| `%s` | %d | " % (short, sum(c.values())) + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
