python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_v6.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mpc_spec -s 4 -c 1 -f -o gpurun_out/prof_spec_F4_1g_v6 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v6.log 2>&1
tail -1 gpurun_out/plain_v6.log | cut -c1-200
