#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libtheoremsearch's objects (cuobjdump -sass build/csrc/*.o).

  python profiles/sass_histogram.py > profiles/sass_r2.txt

For every kernel: instruction count and the counts of the opcodes that prove which hardware path it uses —
UTCHMMA/UTCQMMA/UTCIMMA... (tcgen05.mma), UTMALDG (TMA tensor loads), UBLKCP (1-D bulk TMA), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync), ACQBULK/griddepcontrol (PDL) — plus the ten
most frequent opcodes."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "UTCMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS",
       "HMMA", "IMMA", "QMMA", "LDS", "LDG", "LDGSTS", "RED", "ATOM", "ATOMG", "F2FP", "HFMA2", "FFMA", "DFMA", "ACQBULK",
       "PREEXIT", "CCTL", "MEMBAR", "ERRBAR", "SHFL", "NANOSLEEP")


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "build", "csrc", "*.o")))
    if not objs:
        sys.exit("build first: python -c 'import __graft_entry__ as g; g.build()'")
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        kernels, name = collections.OrderedDict(), None
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = m.group(1)
                kernels[name] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and name:
                kernels[name][m.group(1)] += 1
        print(f"== {os.path.basename(obj)}")
        for kn, c in kernels.items():
            dem = subprocess.run(["cu++filt", kn], capture_output=True, text=True).stdout.strip() or kn
            dem = re.sub(r"\(.*", "", dem)
            total = sum(c.values())
            keys = " ".join(f"{k}={c[k]}" for k in KEY if c.get(k))
            top = " ".join(f"{k}:{v}" for k, v in c.most_common(10))
            print(f"  {dem}  [{total} instr]\n      key: {keys}\n      top: {top}")


if __name__ == "__main__":
    main()
