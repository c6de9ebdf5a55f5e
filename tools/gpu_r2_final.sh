#!/bin/bash
# Round 2, final GPU validation (one GPU): the suite as the driver runs it (timed), the suite again with the full
# parity record, smoke, default bench + reference arm.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2_final.sh'
set -u
out=gpurun_out/r2final
mkdir -p "$out"
/usr/bin/time -v timeout 1200 python -m pytest tests -m gpu -q -x > "$out/pytest_gpu.log" 2> "$out/pytest_gpu.time"
echo "pytest -m gpu (driver style): exit $?" | tee "$out/summary.txt"
tail -4 "$out/pytest_gpu.log" >> "$out/summary.txt"
grep "Elapsed" "$out/pytest_gpu.time" >> "$out/summary.txt"
LCU_PARITY_FULL_RECORD=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > "$out/pytest_full_record.log" 2>&1
echo "pytest full record: exit $?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_full_record.log" >> "$out/summary.txt"
cp gpurun_out/parity_report.json "$out/parity_report.json" 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> "$out/summary.txt" 2>&1
timeout 600 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"
timeout 300 python tools/latency.py > "$out/latency.log" 2>&1
cp gpurun_out/latency.json "$out/latency.json" 2>/dev/null
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys
for tag in ("default", "reference"):
    try:
        d = json.loads([l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1])
        print(tag, d["value"], d["unit"], "e2e", d["e2e"]["value"], "stages", d.get("stage_ms_per_step"), "frac", d.get("roofline", {}).get("frac"),
              "c5", (d.get("c5") or {}).get("value"), (d.get("c5") or {}).get("roofline", {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
              "sustained", (d.get("sustained") or {}).get("value"))
    except Exception as e:
        print(tag, "no bench line:", e)
PY
grep -h "us_per_eval" "$out/latency.log" | cut -c1-400 >> "$out/summary.txt"
tail -30 "$out/summary.txt"
