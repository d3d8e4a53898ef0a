"""SASS evidence per kernel of libpnb200.so (runs in the build container, no GPU needed):
counts of the Blackwell tensor-core / tensor-memory / bulk-copy mnemonics in every kernel's sm_100a SASS.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit (mbarrier arrive),
UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy), USETMAXREG = setmaxnreg (warpgroup register re-allocation), UTMALDG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier ops,
HMMA = legacy mma.sync (must be 0 everywhere)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pointnerf_b200", "csrc", "libpnb200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "USETMAXREG", "SYNCS", "HMMA", "FFMA", "MUFU", "ATOM", "RED", "LDG", "STG", "SHFL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_instr"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_instr"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (mn in ("ATOM", "RED") and op.startswith(mn)):
                    kernels[cur][mn] += 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("SASS mnemonic counts per kernel of %s (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-44s %7s " % ("kernel", "instr") + " ".join("%7s" % m for m in MNEMONICS))
    for (name, c), d in zip(kernels.items(), dem):
        short = re.sub(r"\(.*", "", d).replace("pnb::", "").replace("bw::", "")
        print("%-44s %7d " % (short[:44], c["_instr"]) + " ".join("%7d" % c[m] for m in MNEMONICS))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("%-44s %7d " % ("TOTAL", tot["_instr"]) + " ".join("%7d" % tot[m] for m in MNEMONICS))


if __name__ == "__main__":
    main()
