"""Per-kernel SASS opcode counts of the built objects (cuobjdump -sass): the mnemonics that prove the Blackwell-native
paths - UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor loads / stores), UTCBAR (tcgen05.commit),
STSM / LDSM (stmatrix / ldmatrix), HMMA (mma.sync), SYNCS (mbarrier).  Usage: python tools/sass_opcounts.py > profiles/rNN_sass_opcounts.txt"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "STSM", "LDSM", "HMMA", "SYNCS", "REDG",
       "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    objs = sorted((ROOT / "d3roma_b200" / "csrc" / "build").glob("*.o"))
    head = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass opcode counts per kernel (sm_100a objects built from the tree at/after commit {head}; "
          f"nvcc {subprocess.run(['nvcc', '--version'], capture_output=True, text=True).stdout.split('release ')[-1].split(',')[0]})")
    print("# columns: " + " ".join(OPS) + " | total instructions")
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", str(o)], capture_output=True, text=True).stdout
        per = defaultdict(Counter)
        cur = None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and cur:
                op = m.group(1)
                per[cur]["__total"] += 1
                base = op.split(".")[0]
                per[cur][base] += 1
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    per[cur]["UTCHMMA.2CTA"] += 1
        if not per:
            continue
        names = demangle(list(per))
        print(f"\n== {o.name}")
        for fn, c in per.items():
            short = re.sub(r"\(anonymous namespace\)::", "", names.get(fn, fn))
            short = re.sub(r"\(.*", "", short).replace("void d3::", "")
            print(f"{short:<64s} " + " ".join(f"{c.get(k, 0):>4d}" for k in OPS) + f" | {c['__total']}")


if __name__ == "__main__":
    main()
