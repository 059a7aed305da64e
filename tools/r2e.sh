mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2e_pytest.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop > gpurun_out/r2e_ops_coop.json 2> gpurun_out/r2e_ops_coop.err
MT_DEBUG_FLAGS=32 timeout 300 python tools/time_ops.py --tag chain > gpurun_out/r2e_ops_chain.json 2> gpurun_out/r2e_ops_chain.err
BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
MT_DEBUG_FLAGS=32 BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 --single > gpurun_out/r2e_bench_chain.json 2> gpurun_out/r2e_bench_chain.err
tail -3 gpurun_out/r2e_pytest.log; tail -2 gpurun_out/r2e_bench.err; tail -2 gpurun_out/r2e_bench_chain.err
