#!/bin/bash
# First GPU step for -DLCU_PF_LIBM_PAIR=1 (atan2 / sincos / pow / powr of the
# two-rays kernel written out with packed arithmetic, shim.cuh).
# The switch has only CPU evidence so far (tests/test_pair_math.py,
# tests/test_pair_rays.py).  One gpurun call, about 6 minutes of box time:
#
#   gpurun --timeout 900 -- 'bash tools/gpu_check_libm_pair.sh'
#
# 1. the bit-parity tests of the two-rays kernel against the one-ray kernel and
#    the random-model parity tests, with the switch on (they cover every EPL
#    configuration, C5 and both math modes);
# 2. C5 throughput without and with it (same box, back to back);
# 3. ncu: launch list and one full capture of lcu_render_pair on C5 with it.
# Everything lands in gpurun_out/libm_pair/.
set -u
out=gpurun_out/libm_pair
mkdir -p "$out"
FLAG=-DLCU_PF_LIBM_PAIR=1

LCU_NVRTC_FLAGS=$FLAG timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "two_rays or random_models or c5" > "$out/parity_on.log" 2>&1
echo "parity with $FLAG: exit $?" | tee "$out/summary.txt"

timeout 300 python bench.py --workload c5 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench_c5_off.json" 2> "$out/bench_c5_off.err"
LCU_NVRTC_FLAGS=$FLAG timeout 300 python bench.py --workload c5 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline \
    > "$out/bench_c5_on.json" 2> "$out/bench_c5_on.err"
python - "$out" <<'EOF' | tee -a "$out/summary.txt"
import json, sys
for tag in ("off", "on"):
    try:
        line = [l for l in open(f"{sys.argv[1]}/bench_c5_{tag}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(tag, d["value"], d["unit"], "render ms", d.get("stage_ms_per_step", {}).get("render"), "lnew rel", d.get("parity_lnew_rel"))
    except Exception as e:
        print(tag, "no bench line:", e)
EOF

LCU_NVRTC_FLAGS=$FLAG timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file "$out/launches_c5_on.csv" python bench.py --workload c5 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline \
    > "$out/ncu_launches.log" 2>&1
LCU_NVRTC_FLAGS=$FLAG timeout 400 ncu --set full --clock-control none --import-source on -k regex:lcu_render_pair -s 3 -c 1 \
    -o "$out/render_pair_c5_on" python bench.py --workload c5 --batch 2 --steps 1 --warmup 3 --no-cpu-baseline \
    > "$out/ncu_full.log" 2>&1
python tools/ncu_summary.py "$out/render_pair_c5_on.ncu-rep" "$out/render_pair_c5_on.txt" >> "$out/summary.txt" 2>&1
cat "$out/summary.txt"
