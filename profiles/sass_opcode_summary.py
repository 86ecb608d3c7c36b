"""Static SASS opcode counts per kernel of the tcgen05 kernels (cuobjdump -sass on the in-tree object file).
usage: python profiles/sass_opcode_summary.py [object file] > profiles/rNN_sass_opcode_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "d2d-ppo_b200", "csrc", "build", "learner_api.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "SYNCS", "MUFU.EX2", "MUFU.RCP", "F2FP", "HMMA", "RED.E", "ATOMG",
        "STS", "LDS", "BAR.SYNC"]
counts, cur, order = collections.defaultdict(collections.Counter), None, []
it = iter(names)
for line in sass.splitlines():
    if "Function :" in line:
        cur = next(it).replace("d2d::", "").replace("void ", "")
        order.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + ".") or op.startswith(c):
                counts[cur][c] += 1
print("SASS opcode counts (static, per kernel) of the tcgen05 kernels in d2d-ppo_b200/csrc/learner_api.cu, built with")
print("nvcc 12.9 -gencode arch=compute_100a,code=sm_100a; command: python profiles/sass_opcode_summary.py")
print("UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops.")
print("No UTMALDG (TMA): operands are staged by the threads because the fp32 -> fp16-plane split happens in registers.\n")
print(f"{'kernel':78s}" + "".join(f"{c:>9s}" for c in COLS))
for k in order:
    if counts[k]["UTCHMMA"]:
        print(f"{k[:78]:78s}" + "".join(f"{counts[k][c]:9d}" for c in COLS))
