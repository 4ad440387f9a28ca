#!/bin/bash
# per-kernel device times + executed warp instructions of one forward / backward call (ncu launch list; caches left as they are,
# serialised: compare shares).   bash scripts/fwd_times.sh [fwd|bwd|fwd,bwd] [extra kernel_loop.py args]
what=${1:-fwd}; shift
python scripts/kernel_loop.py --iters 1 --warmup 1 --what $what "$@" > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches.csv python scripts/kernel_loop.py --iters 1 --warmup 1 --what $what "$@" > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); im=hdr.index('Metric Name'); iid=hdr.index('ID'); iu=hdr.index('Metric Unit')
per=collections.OrderedDict()
for r in rows[1:]:
    d=per.setdefault(r[iid], {'k': r[ik].split('(')[0]})
    v=float(r[iv].replace(',',''))
    u=r[iu]
    if 'time' in r[im]: v*= {'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
    if 'bytes' in r[im]: v*= {'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u,1)
    d[r[im]]=v
out=list(per.values()); n=len(out)//2
tot=0
for d in out[-n:]:
    t=d['gpu__time_duration.sum']; tot+=t
    print(f"{t:8.1f} us {d['smsp__inst_executed.sum']/1e6:7.2f} Minst  rd {d['dram__bytes_read.sum']:7.1f} MB wr {d['dram__bytes_write.sum']:7.1f} MB  {d['k']}")
print("sum", round(tot,1))
PY
