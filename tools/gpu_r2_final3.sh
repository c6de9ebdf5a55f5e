set -u
out=gpurun_out/r2final3; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tiny_and_ragged or async_pair or empty_and_ragged" > $out/pytest_tiny.log 2>&1
echo "tiny images: exit $?" | tee $out/summary.txt
tail -15 $out/pytest_tiny.log >> $out/summary.txt
cat $out/summary.txt
