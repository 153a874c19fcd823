#!/bin/bash
# per-kernel device time list (ncu, 1 replay) for one step at 2M reads incl. the fused e2e path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --reads ${READS:-2000000} --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]
iK=hdr.index('Kernel Name'); iV=hdr.index('Metric Value'); iU=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    nm=r[iK].split('(')[0][-48:]
    v=float(r[iV].replace(',',''));
    if r[iU]=='ns': v/=1e6
    elif r[iU]=='us': v/=1e3
    elif r[iU]=='s': v*=1e3
    e=agg.setdefault(nm,[0,0.0]); e[0]+=1; e[1]+=v
tot=sum(v[1] for v in agg.values())
for nm,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f'{nm:50s} n={c:4d} total_ms={t:9.3f} ({100*t/tot:4.1f}%)')
PY
