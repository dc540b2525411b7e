#!/usr/bin/env python
"""SASS opcode histogram per kernel of libwsage.so (cuobjdump -sass; runs on the CPU box)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

SO = Path(__file__).resolve().parent.parent / "scdeepsort_b200" / "csrc" / "libwsage.so"
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "FFMA", "LDS", "STS", "LDG", "STG", "RED", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    print(f"SASS opcode histogram of {SO.relative_to(SO.parents[2])} (cuobjdump -sass, sm_100a), final build of round 2.")
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / "
          "reduce-add, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops.\n")
    print(f"{'kernel':<70}" + "".join(f"{o:>9}" for o in OPS) + f"{'total':>9}")
    blocks = re.split(r"Function : \S+", sass)[1:]
    for name, body in zip(names, blocks):
        cnt = collections.Counter()
        total = 0
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", body, re.M):
            op = m.group(1)
            total += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    cnt[o] += 1
        short = re.sub(r"\(.*", "", name)[:69]
        print(f"{short:<70}" + "".join(f"{cnt[o]:>9}" for o in OPS) + f"{total:>9}")


if __name__ == "__main__":
    main()
