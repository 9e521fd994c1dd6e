#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS opcodes in the built objects (evidence that the tensor-core kernels are
tcgen05 / TMEM / bulk-copy code): python profiles/make_sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTCCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG"]
print("# opcode counts per kernel (cuobjdump -sass of torchdyn_b200/csrc/*.o, sm_100a); UTCHMMA = tcgen05.mma, LDTM/STTM = "
      "tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
for obj in sorted(glob.glob(os.path.join(ROOT, "torchdyn_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                for o in OPS:
                    if op == o or op.startswith(o + "."):
                        counts[cur][o] += 1
    print(f"\n## {os.path.basename(obj)}")
    for k, c in counts.items():
        print(f"{k}: " + ", ".join(f"{o}={c[o]}" for o in OPS if c[o]))
