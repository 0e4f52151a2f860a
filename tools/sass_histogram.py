#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass; runs without a GPU):
    python tools/sass_histogram.py > profiles/sass_r2.txt
Shows which Blackwell-specific instructions each hand-written kernel is made of: UBLKCP (bulk-copy TMA), SYNCS (mbarrier),
IDP.4A (integer dot product), REDUX (warp reduction), POPC, and that no tensor-core or library code is in the path."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "stm32f4_sdr_gps_b200" / "lib" / "libgpsb_cuda.so"
KEY = ["UBLKCP", "SYNCS", "IDP", "REDUX", "POPC", "ATOMS", "BAR", "LDG", "LDS", "STS", "STG", "SHF", "LOP3", "PRMT", "IMAD",
       "SHFL", "MUFU", "F2I", "I2F", "DFMA", "HMMA", "UTC", "CALL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    arch = re.search(r"arch = (sm_\w+)", out)
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("gpsb::", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            op = m.group(1)
            kernels[name][op.split(".")[0]] += 1
            if op.startswith(("IDP.4A", "SYNCS.", "UBLKCP", "REDUX", "ATOMS")):
                kernels[name][".".join(op.split(".")[:2])] += 0     # keep the qualified form in view
                kernels[name]["~" + ".".join(op.split(".")[:3])] += 1
    print("SASS opcode histogram of %s (%s), cuobjdump -sass, per kernel: instruction count, then selected opcodes" %
          (LIB.name, arch.group(1) if arch else "?"))
    print("(UBLKCP = cp.async.bulk TMA copy, SYNCS = mbarrier, IDP.4A = dp4a, REDUX = warp reduce; HMMA / UTC* = tensor cores: none)\n")
    for k, c in kernels.items():
        total = sum(v for op, v in c.items() if not op.startswith("~"))
        picks = ["%s %d" % (op, c[op]) for op in KEY if c.get(op)]
        detail = sorted(op[1:] + " %d" % v for op, v in c.items() if op.startswith("~"))
        print("%-58s %6d  %s" % (k[:58], total, "  ".join(picks)))
        if detail:
            print("%-58s         %s" % ("", "  ".join(detail)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
