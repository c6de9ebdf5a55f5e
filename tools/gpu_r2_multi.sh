#!/bin/bash
# Round 2, multi-GPU call: N = number of visible GPUs.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_multi.sh 2'
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_r2_multi.sh 8 4'
# For every N given: the NCCL tests (first N only), then bench.py under torchrun:
# C5 strong scaling (64 points per step in total), C5 row strips (4 points), C4 weak.
set -u
out=gpurun_out/r2multi
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt"
first=$1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > "$out/pytest_multi_gpu_n$first.log" 2>&1
echo "pytest test_multi_gpu (N=$first): exit $?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_multi_gpu_n$first.log" >> "$out/summary.txt"
port=29500
for n in "$@"; do
  for cfg in "c5_strong --workload c5" "c5_rows --workload c5 --parallelism rows" "c4_weak --workload c4 --sustain 0"; do
    set -- $cfg; tag=$1; shift
    port=$((port+1))
    NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps 5 --warmup 3 "$@" > "$out/bench_${tag}_n$n.json" 2> "$out/bench_${tag}_n$n.err"
    echo "$tag N=$n exit $?" >> "$out/summary.txt"
    grep -c "Init COMPLETE" "$out/bench_${tag}_n$n.err" >> "$out/summary.txt"
  done
done
python - "$out" <<'PY' | tee -a "$out/summary.txt"
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(os.path.basename(f), d["n_gpus"], d["scaling"], d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"],
              "stages", d["stage_ms_per_step"], d.get("rows"))
    except Exception as e:
        print(f, "no bench line:", e)
PY
