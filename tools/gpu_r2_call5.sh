#!/bin/bash
# Round 2, GPU call 5 (one GPU): set_params folded into the small-launch render kernels, ensemble noise floors.
#   gpurun --timeout 1800 -- 'bash tools/gpu_r2_call5.sh'
set -u
out=gpurun_out/r2c5
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/summary.txt"
tail -25 "$out/pytest_gpu.log" >> "$out/summary.txt"
cp gpurun_out/parity_report.json "$out/parity_report.json" 2>/dev/null
timeout 300 python tools/latency.py > "$out/latency.log" 2>&1
cp gpurun_out/latency.json "$out/latency.json" 2>/dev/null
cat "$out/latency.log" >> "$out/summary.txt"
# the C host, with and without the folded set_params
gcc -std=c99 -O1 -I include tests/c/host_check.c -L lensed_b200 -llensed_cuda -Wl,-rpath,$PWD/lensed_b200 -lm -o /tmp/host_check
for i in 1 2 3; do LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
for i in 1 2 3; do LCU_NO_FOLD_SETTER=1 LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench_c5_strong_n1.json" 2> "$out/bench_c5_strong_n1.err"
timeout 300 python bench.py --workload c5 --parallelism rows --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench_c5_rows_n1.json" 2> "$out/bench_c5_rows_n1.err"
timeout 600 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys
for tag in ("c5_strong_n1", "c5_rows_n1", "default", "reference"):
    try:
        d = json.loads([l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1])
        print(tag, d["value"], d["unit"], "e2e", d["e2e"]["value"], "stages", d.get("stage_ms_per_step"), "frac", d.get("roofline", {}).get("frac"),
              "c5", (d.get("c5") or {}).get("value"), "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(tag, "no bench line:", e)
PY
# evidence: full ncu capture of the C5 render kernel as built now (4 resident blocks) and of the FFMA micro-benchmark
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lcu_render_pair -s 3 -c 1 \
    -o "$out/render_pair_c5_mb4" python bench.py --workload c5 --scaling weak --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > "$out/ncu_c5.log" 2>&1
python tools/ncu_summary.py "$out/render_pair_c5_mb4.ncu-rep" "$out/render_pair_c5_mb4.txt" >> "$out/summary.txt" 2>&1
timeout 400 ncu --set full --clock-control none -k regex:ffma_kernel -c 1 \
    -o "$out/ffma_peak" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 --sustain 0 > "$out/ncu_ffma.log" 2>&1
python tools/ncu_summary.py "$out/ffma_peak.ncu-rep" "$out/ffma_peak.txt" >> "$out/summary.txt" 2>&1
tail -60 "$out/summary.txt"
