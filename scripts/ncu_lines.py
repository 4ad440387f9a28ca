"""Aggregates ncu per-SASS-instruction counters by CUDA source line (needs -lineinfo builds).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep k_resize eot_fwd [--metric "Instructions Executed"] [--top 30]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern, tu = sys.argv[1], sys.argv[2], sys.argv[3]
metric = "Instructions Executed"
top = 30
if "--metric" in sys.argv:
    metric = sys.argv[sys.argv.index("--metric") + 1]
if "--top" in sys.argv:
    top = int(sys.argv[sys.argv.index("--top") + 1])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "mladversarialobjectdetection_b200", "libeotpatch.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, f"{tu}.sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
per_instr, cur, on = [], None, False
for l in dis:
    if l.startswith("\t.section\t.text."):
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and ".dword" not in l and ".word" not in l:
        per_instr.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(r for r in rows if "Source" in r and "Address" in r)
body = []
for r in rows[rows.index(hdr) + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
ie = hdr.index(metric)
n = min(len(per_instr), len(body))
agg = collections.Counter()
for i in range(n):
    try:
        agg[per_instr[i]] += float(body[i][ie])
    except ValueError:
        pass
tot = sum(agg.values()) or 1.0
src = {}
print(f"{kern}: {metric} total {tot:.0f} over {n} SASS instructions ({len(per_instr)} disassembled, {len(body)} profiled)")
for k, v in agg.most_common(top):
    if k is None:
        print(f"{v / tot:6.1%} <no line>")
        continue
    fn, ln = k
    if fn not in src:
        p = os.path.join(root, "mladversarialobjectdetection_b200", "csrc", fn)
        src[fn] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[fn][ln - 1].strip()[:105] if 0 < ln <= len(src[fn]) else ""
    print(f"{v / tot:6.1%} {fn}:{ln:4d}  {text}")
