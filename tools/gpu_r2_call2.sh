#!/bin/bash
# Round 2, GPU call 2 (one GPU):
#   gpurun --timeout 1800 -- 'bash tools/gpu_r2_call2.sh'
# 1. the whole GPU suite on the reference's own object files, floor-aware parity recorded (parity_report.json);
# 2. the default bench line (C4 + C5 sub-record + sustained + cpu baseline + parity) and the reference arm;
# 3. C5 occupancy A/B (LCU_PAIR_MINBLOCKS 3 / 4 / 5);
# 4. launch list of the default step.
set -u
out=gpurun_out/r2c2
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/summary.txt"
tail -30 "$out/pytest_gpu.log" >> "$out/summary.txt"
cp gpurun_out/parity_report.json "$out/parity_report.json" 2>/dev/null

timeout 600 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"
for mb in 3 4 5; do
  LCU_NVRTC_FLAGS=-DLCU_PAIR_MINBLOCKS=$mb timeout 300 python bench.py --workload c5 --scaling weak --batch 8 --steps 5 --warmup 3 --no-cpu-baseline \
      > "$out/bench_c5_mb$mb.json" 2> "$out/bench_c5_mb$mb.err"
done
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys
for tag in ("default", "reference", "c5_mb3", "c5_mb4", "c5_mb5"):
    try:
        line = [l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(tag, d["value"], d["unit"], "e2e", d["e2e"]["value"], "render ms", d.get("stage_ms_per_step", {}).get("render"),
              "frac", d.get("roofline", {}).get("frac"), "sustained", (d.get("sustained") or {}).get("value"),
              "c5", (d.get("c5") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if "parity" in d: print(" parity", json.dumps(d["parity"])[:1500])
    except Exception as e:
        print(tag, "no bench line:", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file "$out/launches_default.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --sustain 0 \
    > "$out/ncu_launches.log" 2>&1
cat "$out/summary.txt"
