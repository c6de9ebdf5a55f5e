#!/bin/bash
# Round 2, GPU call 10: all stages of one point in one kernel (lcu_point_s*): bits, give-up path, latency.
set -u
out=gpurun_out/r2c10
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_host.py -m gpu -q -x -k "one_kernel_point or single_point or async_pair or batch_and_split or latency or loglike_from_c or examples or sampler" > "$out/pytest_subset.log" 2>&1
echo "pytest subset: exit $?" | tee "$out/summary.txt"
tail -25 "$out/pytest_subset.log" >> "$out/summary.txt"
timeout 300 python tools/latency.py > "$out/latency.log" 2>&1
cp gpurun_out/latency.json "$out/latency.json" 2>/dev/null
cat "$out/latency.log" >> "$out/summary.txt"
gcc -std=c99 -O1 -I include tests/c/host_check.c -L lensed_b200 -llensed_cuda -Wl,-rpath,$PWD/lensed_b200 -lm -o /tmp/host_check
for i in 1 2 3; do LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
for i in 1 2; do LCU_NO_FUSED_POINT=1 LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
tail -40 "$out/summary.txt"
