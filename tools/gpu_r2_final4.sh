set -u
out=gpurun_out/r2final4; mkdir -p $out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/latency_launches.csv python tools/latency_kernels.py > $out/ncu.log 2>&1
python - $out <<'PY'
import csv, sys, collections
rows=[r for r in csv.reader(open(sys.argv[1]+"/latency_launches.csv")) if len(r)>14 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[(r[4][:28], r[8])].append(float(r[14]))
for k,v in agg.items():
    v=sorted(v); print(k, len(v), "median ns", v[len(v)//2], "min", v[0])
PY
