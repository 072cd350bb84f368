"""Writes profiles/<tag>_sass_listing.txt: per kernel of libhyphy_b200.so the counts of the opcodes that prove which
hardware paths the code uses (tcgen05: UTCHMMA / LDTM / STTM / UTCBAR; bulk copies: UBLKCP / UBLKPF; FP64 tensor pipe:
DMMA; waterfall loops around uniform-operand instructions: ELECT + BRA.U.ANY) and the first occurrence of each.
    python tools/sass_listing.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "hyphy_b200", "libhyphy_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOM", "UBLKCP", "UBLKPF", "UTMALDG", "DMMA", "HMMA", "ELECT", "BRA.U.ANY", "R2UR", "SYNCS", "DFMA", "FFMA"]
out = [f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a); opcode counts per kernel and first occurrence", ""]
cur, counts, first = None, None, None


def flush():
    if cur is None:
        return
    short = re.sub(r"^_ZN3hb2\d+", "", cur)
    out.append(f"## {cur}")
    out.append("   " + "  ".join(f"{k}:{v}" for k, v in counts.items() if v))
    for k, v in first.items():
        out.append(f"      first {k:10s} {v}")
    out.append("")


for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, counts, first = m.group(1), collections.OrderedDict((o, 0) for o in OPS), collections.OrderedDict()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1).strip()
    for o in OPS:
        if re.search(r"(^|\s|@!?U?P\d\s+)" + re.escape(o) + r"(\.|\s|$)", ins):
            counts[o] += 1
            if o in ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "DMMA") and o not in first:
                first[o] = ins[:150]
flush()
path = os.path.join(ROOT, "profiles", f"{tag}_sass_listing.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out), "lines")
