mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); ns=v*{'ns':1,'us':1e3,'ms':1e6,'s':1e9}.get(r[ui],1)
    agg[r[ki][:64]][0]+=1; agg[r[ki][:64]][1]+=ns
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{v[1]/1e6:10.3f} ms {v[0]:4d}x {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]/1e3:9.1f} us  {k}")
PY
