"""profiles/sass_opcodes.txt: opcode counts per kernel of libsvr_b200.so from `cuobjdump -sass` (evidence that the
tcgen05 / TMA / mbarrier paths are what the library executes)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "single-view-3d-reconstruction_b200" / "libsvr_b200.so"
OPS = ["SYNCS", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "FFMA2", "REDG", "RED", "LDTM", "STTM", "UTCHMMA", "UTCBAR"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
kern, n, cnt = None, Counter(), OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cnt[kern] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        n[kern] += 1
        op = m.group(1).split(".")[0]
        if op in OPS:
            cnt[kern][op] += 1
dst = ROOT / "profiles" / "sass_opcodes.txt"
with open(dst, "w") as f:
    f.write("# SASS opcode evidence: `cuobjdump -sass single-view-3d-reconstruction_b200/libsvr_b200.so` (sm_100a), opcode counts per kernel (tools/sass_opcodes.py).\n"
            "# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store,\n"
            "# UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier, FFMA2 = packed fp32 FMA, REDG/RED = vector reductions.\n"
            "# kernel | SASS instructions | opcode counts\n")
    for k, c in cnt.items():
        if c:
            f.write(f"{k:72s} {n[k]:5d}  " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]) + "\n")
print(dst)
