timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_jit.py -m gpu -x -q 2>&1 | tail -3
bash tools/quickbench.sh F4smooth
bash tools/quickbench.sh F4mixed --kind mixed_hashed
bash tools/quickbench.sh Z1mixed --config Z1 --kind mixed_hashed
bash tools/quickbench.sh E5mixed --config E5 --kind mixed_hashed
bash tools/quickbench.sh F4smooth4G --bytes-per-gpu 4294967296
