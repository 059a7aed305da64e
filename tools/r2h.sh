mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2h_pytest.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop > gpurun_out/r2h_ops_coop.json 2> gpurun_out/r2h_ops_coop.err
MT_DEBUG_FLAGS=64 MT_BWD_RT=4 timeout 300 python tools/time_ops.py --tag coop_rt4 > gpurun_out/r2h_ops_coop_rt4.json 2> gpurun_out/r2h_ops_coop_rt4.err
BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2h_launches_warm.csv python tools/profile_step.py > /dev/null 2> gpurun_out/r2h_ncu.err
tail -2 gpurun_out/r2h_pytest.log
