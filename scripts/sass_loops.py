"""Static view of a kernel's SASS: every backward branch = a loop; prints [start,end] line numbers, length and the
memory ops inside.  python scripts/sass_loops.py OBJ MANGLED_SUBSTRING"""
import re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur=None; funcs={}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur]=[]; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1),16), m.group(2).strip()))
for name, ins in funcs.items():
    if key not in name: continue
    print(name, len(ins), "instructions")
    addr2idx = {a:i for i,(a,_) in enumerate(ins)}
    for i,(a,t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1),16)
            if tgt <= a and tgt in addr2idx:
                j = addr2idx[tgt]
                body = [x for _,x in ins[j:i+1]]
                mem = [re.split(r"\s+", x.lstrip("@!UP0123456789 "))[0] for x in body if re.search(r"\b(LDG|STG|LDS|STS|ATOM|RED|SHFL|LD\.|ST\.)", x)]
                print(f"  loop [{j},{i}] len={i-j+1}  mem: {' '.join(mem[:24])}")
