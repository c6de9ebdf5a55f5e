#!/bin/bash
# Round 2, GPU call 7 (one GPU): compute-sanitizer over this round's new paths, fresh ncu captures of the C4 kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call7.sh'
set -u
out=gpurun_out/r2c7
mkdir -p "$out"
K="async_pair or single_point_graph or row_strips or empty_and_ragged or dumper_layers or convolution_kernels_same_bits"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > "$out/sanitizer_memcheck.log" 2>&1
echo "memcheck exit $?" | tee "$out/summary.txt"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "async_pair or single_point_graph or row_strips" > "$out/sanitizer_racecheck.log" 2>&1
echo "racecheck exit $?" | tee -a "$out/summary.txt"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" "$out"/sanitizer_*.log >> "$out/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lcu_render_pair -s 3 -c 1 \
    -o "$out/render_pair_c4" python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-c5 --sustain 0 > "$out/ncu_c4.log" 2>&1
python tools/ncu_summary.py "$out/render_pair_c4.ncu-rep" "$out/render_pair_c4.txt" >> "$out/summary.txt" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lcu_convolve -s 3 -c 1 \
    -o "$out/convolve_c4" python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-c5 --sustain 0 > "$out/ncu_conv.log" 2>&1
python tools/ncu_summary.py "$out/convolve_c4.ncu-rep" "$out/convolve_c4.txt" >> "$out/summary.txt" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file "$out/launches_c4_B32_dram.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --sustain 0 > "$out/ncu_launches.log" 2>&1
tail -80 "$out/summary.txt"
