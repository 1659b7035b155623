timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_jit.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
MPC_B200_LIB=$PWD/cal_22-mpc_b200/libmpc_b200_prev.so bash tools/quickbench.sh F4smooth_prev
MPC_SCHED_STATIC_EIGHTHS=8 bash tools/quickbench.sh F4smooth_static8
MPC_SCHED_STATIC_EIGHTHS=7 bash tools/quickbench.sh F4smooth_static7
MPC_SCHED_STATIC_EIGHTHS=6 bash tools/quickbench.sh F4smooth_static6
done
MPC_B200_LIB=$PWD/cal_22-mpc_b200/libmpc_b200_prev.so bash tools/quickbench.sh F4smooth4G_prev --bytes-per-gpu 4294967296
MPC_SCHED_STATIC_EIGHTHS=7 bash tools/quickbench.sh F4smooth4G_static7 --bytes-per-gpu 4294967296
