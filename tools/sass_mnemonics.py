#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 / TMEM / TMA) in the built library.

    python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt        # needs cuobjdump; no GPU

UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load
(cp.async.bulk.tensor), UTMASTG = TMA tensor store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier operations,
FFMA2 = packed fp32x2 FMA (sm_100), DMMA = fp64 mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "anncur_b200", "libanncur_b200.so")
KEEP_FULL = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "DMMA", "UTCBAR")
KEEP_BASE = ("UTMAPF", "UTCCP", "FFMA2", "HMMA", "UTCATOMSWS", "SYNCS", "UBLKCP", "UBLKRED")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in KEEP_FULL:
            counts[cur][op] += 1
        elif base in KEEP_BASE:
            counts[cur][base] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of anncur_b200/libanncur_b200.so (cuobjdump -sass, sm_100a), tools/sass_mnemonics.py")
    print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTMASTG = TMA tensor")
    print("# store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA, DMMA = fp64 mma.sync\n")
    total = collections.Counter()
    for name, d in zip(names, dem):
        c = counts[name]
        if not c:
            continue
        print(re.sub(r"\(.*", "", d)[:120])
        print("    " + ", ".join(f"{op} x{n}" for op, n in sorted(c.items())))
        total.update(c)
    print("\nTOTAL  " + ", ".join(f"{op} x{n}" for op, n in sorted(total.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main())
