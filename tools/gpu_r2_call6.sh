#!/bin/bash
# Round 2, GPU call 6 (one GPU): final validation -- GPU suite (k = 2 floors, async pair, folded set_params opt-in),
# single-point latency (graph / folded / plain / two in flight; C host), default bench + reference arm.
#   gpurun --timeout 1800 -- 'bash tools/gpu_r2_call6.sh'
set -u
out=gpurun_out/r2c6
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/summary.txt"
tail -25 "$out/pytest_gpu.log" >> "$out/summary.txt"
cp gpurun_out/parity_report.json "$out/parity_report.json" 2>/dev/null
timeout 300 python tools/latency.py > "$out/latency.log" 2>&1
cp gpurun_out/latency.json "$out/latency.json" 2>/dev/null
cat "$out/latency.log" >> "$out/summary.txt"
gcc -std=c99 -O1 -I include tests/c/host_check.c -L lensed_b200 -llensed_cuda -Wl,-rpath,$PWD/lensed_b200 -lm -o /tmp/host_check
for i in 1 2 3; do LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
for i in 1 2; do LCU_FOLD_SETTER=1 LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
timeout 600 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> "$out/summary.txt" 2>&1
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys
for tag in ("default", "reference"):
    try:
        d = json.loads([l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1])
        print(tag, d["value"], d["unit"], "e2e", d["e2e"]["value"], "stages", d.get("stage_ms_per_step"), "frac", d.get("roofline", {}).get("frac"),
              "c5", (d.get("c5") or {}).get("value"), (d.get("c5") or {}).get("roofline", {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(tag, "no bench line:", e)
PY
tail -40 "$out/summary.txt"
