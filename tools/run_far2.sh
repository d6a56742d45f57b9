mkdir -p gpurun_out
for f in 0 1; do
SR_K1_FAR=$f ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_far$f.csv python tools/prof_run.py k1b 16 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2u_far$f.csv")))
hdr=None; tot=collections.Counter(); cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            name=d['Kernel Name'].split('(')[0].replace('void <unnamed>::','').replace('<unnamed>::','').split('<')[0]
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            ms=v*{'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3,'nsecond':1e-6,'usecond':1e-3,'msecond':1.0,'second':1e3}.get(u,1e-6)
            tot[name]+=ms; cnt[name]+=1
print("FAR=$f")
for k,v in tot.most_common(6): print("  %-24s %4d launches %8.3f ms total %8.3f ms each"%(k,cnt[k],v,v/cnt[k]))
PY
done
