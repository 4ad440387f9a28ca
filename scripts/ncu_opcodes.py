"""Executed warp-instructions by SASS opcode for one kernel of an ncu report (source page).
    python scripts/ncu_opcodes.py REP KERNEL_REGEX [--top N]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(r for r in rows if "Source" in r and "Address" in r)
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
agg = collections.Counter(); tot = 0.0
for r in rows[rows.index(hdr) + 1:]:
    if r and r[0] == "Kernel Name": break
    if len(r) != len(hdr): continue
    try: n = float(r[iex])
    except ValueError: continue
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    agg[op] += n; tot += n
print(f"{kern}: {tot/1e6:.2f} M warp-instructions")
for k, v in agg.most_common(top): print(f"{v/tot:6.1%} {v/1e6:8.2f}M  {k}")
