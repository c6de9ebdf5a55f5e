#!/bin/bash
# Round 2, GPU call 1: LCU_PF_LIBM_PAIR=1 is now the default.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
# 1. the whole GPU suite with the new default;
# 2. C5 throughput with the switch off and on (same box, back to back), C4 default;
# 3. ncu: launch list + one full capture of lcu_render_pair on C5 (default build).
set -u
out=gpurun_out/r2c1
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$out/gpu.txt" 2>&1
nproc >> "$out/gpu.txt"; lscpu | grep -E "Model name|Flags" | cut -c1-400 >> "$out/gpu.txt"

timeout 900 python -m pytest tests -m gpu -q -x > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/summary.txt"
tail -5 "$out/pytest_gpu.log" >> "$out/summary.txt"

LCU_NVRTC_FLAGS=-DLCU_PF_LIBM_PAIR=0 timeout 300 python bench.py --workload c5 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline \
    > "$out/bench_c5_off.json" 2> "$out/bench_c5_off.err"
timeout 300 python bench.py --workload c5 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline \
    > "$out/bench_c5_on.json" 2> "$out/bench_c5_on.err"
timeout 300 python bench.py --workload c5 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --math strict \
    > "$out/bench_c5_on_strict.json" 2> "$out/bench_c5_on_strict.err"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$out/bench_c4.json" 2> "$out/bench_c4.err"
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys
for tag in ("c5_off", "c5_on", "c5_on_strict", "c4"):
    try:
        line = [l for l in open(f"{sys.argv[1]}/bench_{tag}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(tag, d["value"], d["unit"], "render ms", d.get("stage_ms_per_step", {}).get("render"), "frac", d["roofline"]["frac"], d.get("clocks"))
    except Exception as e:
        print(tag, "no bench line:", e)
PY

timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file "$out/launches_c5.csv" python bench.py --workload c5 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline \
    > "$out/ncu_launches.log" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lcu_render_pair -s 3 -c 1 \
    -o "$out/render_pair_c5" python bench.py --workload c5 --batch 2 --steps 1 --warmup 3 --no-cpu-baseline \
    > "$out/ncu_full.log" 2>&1
python tools/ncu_summary.py "$out/render_pair_c5.ncu-rep" "$out/render_pair_c5.txt" >> "$out/summary.txt" 2>&1
cat "$out/summary.txt"
