"""Instruction histogram of the built library's SASS, per kernel: which functions carry tcgen05 (UTCHMMA), TMEM loads
(LDTM), TMA bulk / tensor copies (UBLKCP / UTMALDG / UTMASTG), cp.async (LDGSTS), packed FP32 FMA (FFMA2) ...

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt

Runs in the build container (cuobjdump disassembles the sm_100a cubins without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zaru_b200", "libzaru_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "HMMA", "LDG", "STG", "LDS", "STS",
       "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
                    break
    names = list(per)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static SASS, sm_100a)")
    print("# " + " ".join(f"{o:>8}" for o in ["total"] + OPS) + "  kernel")
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = per[n]
        tot.update(c)
        short = d.replace("zb::(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\((?:zb::|float|int|unsigned|long|char|const).*\)$", "", short)
        print("  " + " ".join(f"{c[o]:>8}" for o in ["_total"] + OPS) + "  " + short)
    print("# " + " ".join(f"{tot[o]:>8}" for o in ["_total"] + OPS) + "  ALL KERNELS")
    print("# tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, tcgen05.commit/mbarrier = UTCBAR/SYNCS, cp.async.bulk = UBLKCP,")
    print("# cp.async.bulk.tensor (tensor-map TMA) = UTMALDG/UTMASTG, cp.async = LDGSTS, packed FP32 FMA = FFMA2")
    return 0


if __name__ == "__main__":
    sys.exit(main())
