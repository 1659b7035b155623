timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash tools/quickbench.sh F4smooth
bash tools/quickbench.sh F4mixed --kind mixed_hashed
bash tools/quickbench.sh Z1mixed --config Z1 --kind mixed_hashed
bash tools/quickbench.sh Z1smooth --config Z1
bash tools/quickbench.sh E5mixed --config E5 --kind mixed_hashed
bash tools/quickbench.sh P6mixed --config P6 --kind mixed_hashed
bash tools/quickbench.sh P6smooth --config P6
bash tools/quickbench.sh F4zero --kind zero
bash tools/quickbench.sh F4smooth4G --bytes-per-gpu 4294967296
