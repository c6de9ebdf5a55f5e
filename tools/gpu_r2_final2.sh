set -u
out=gpurun_out/r2final2; mkdir -p $out
s=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1
echo "pytest -m gpu (driver style): exit $? in $(( $(date +%s) - s )) s" | tee $out/summary.txt
tail -4 $out/pytest_gpu.log >> $out/summary.txt
tools/micro/f32x2_bench > $out/f32x2_bench.log 2>&1
cat $out/f32x2_bench.log >> $out/summary.txt
cat $out/summary.txt
