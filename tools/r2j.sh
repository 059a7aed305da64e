mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2j_pytest.log
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
BENCH_VERBOSE=1 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err
tail -3 gpurun_out/r2j_pytest.log; tail -c 600 gpurun_out/r2j_bench.json; tail -c 400 gpurun_out/r2j_bench_ref.json
