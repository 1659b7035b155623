set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2_tests.log
for b in 67108864 268435456 1073741824 4294967296 17179869184; do
  bash tools/quickbench.sh "size$b" --bytes-per-gpu $b --steps 20 >> gpurun_out/s2_sweep.log 2>&1
done
python bench.py > gpurun_out/s2_bench.log 2>&1
cat gpurun_out/s2_tests.log gpurun_out/s2_sweep.log
