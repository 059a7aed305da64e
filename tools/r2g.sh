mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2g_pytest.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop > gpurun_out/r2g_ops_coop.json 2> gpurun_out/r2g_ops_coop.err
MT_DEBUG_FLAGS=64 MT_FORCE_RT=2 timeout 300 python tools/time_ops.py --tag coop_rt2 > gpurun_out/r2g_ops_coop_rt2.json 2> gpurun_out/r2g_ops_coop_rt2.err
MT_DEBUG_FLAGS=192 timeout 300 python tools/time_ops.py --tag coop_scalar_diag > gpurun_out/r2g_ops_coop_sd.json 2> gpurun_out/r2g_ops_coop_sd.err
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --px 512 --reps 5 --tag coop512 > gpurun_out/r2g_ops512_coop.json 2> gpurun_out/r2g_ops512_coop.err
BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_pytest.log; tail -2 gpurun_out/r2g_bench.err
