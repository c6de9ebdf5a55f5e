set -u
out=gpurun_out/r2final6; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" | tee $out/summary.txt; tail -2 $out/smoke.log >> $out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_default.json 2> $out/bench_default.err; echo "bench exit $?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference exit $?" >> $out/summary.txt
python - $out <<'PY' >> $out/summary.txt
import json, sys
for tag in ("default", "reference"):
    d = json.loads([l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1])
    print(tag, d["value"], "e2e", d["e2e"]["value"], "frac", d.get("roofline", {}).get("frac"), "c5", (d.get("c5") or {}).get("value"),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "sustained", (d.get("sustained") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
cat $out/summary.txt
