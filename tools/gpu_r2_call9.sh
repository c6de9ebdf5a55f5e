#!/bin/bash
# Round 2, GPU call 9: split kernels with two quadrature points per pass (lcu_render_q_s*): bits, latency, small batches.
set -u
out=gpurun_out/r2c9
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_host.py -m gpu -q -x -k "batch_and_split or single_point or async_pair or examples or latency or loglike_from_c or convolution_kernels_same_bits or sampler" > "$out/pytest_subset.log" 2>&1
echo "pytest subset: exit $?" | tee "$out/summary.txt"
tail -5 "$out/pytest_subset.log" >> "$out/summary.txt"
timeout 300 python tools/latency.py > "$out/latency.log" 2>&1
cp gpurun_out/latency.json "$out/latency.json" 2>/dev/null
cat "$out/latency.log" >> "$out/summary.txt"
gcc -std=c99 -O1 -I include tests/c/host_check.c -L lensed_b200 -llensed_cuda -Wl,-rpath,$PWD/lensed_b200 -lm -o /tmp/host_check
for i in 1 2 3; do LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
for i in 1 2; do LCU_NO_SPLIT_PAIR=1 LENSED_PATH=$PWD/tests/golden /tmp/host_check latency 0 100 5000; done >> "$out/summary.txt" 2>&1
timeout 300 python tools/small_batch_profile.py > "$out/small_batch.log" 2>&1
cp gpurun_out/small_batch.json "$out/small_batch.json" 2>/dev/null
LCU_NO_SPLIT_PAIR=1 timeout 300 python tools/small_batch_profile.py > "$out/small_batch_nopairq.log" 2>&1
grep -h " 16 \| 64 " "$out/small_batch.log" "$out/small_batch_nopairq.log" | cut -c1-200 >> "$out/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> "$out/summary.txt" 2>&1
tail -40 "$out/summary.txt"
