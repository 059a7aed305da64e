mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2s_tests.log
timeout 600 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
cat gpurun_out/r2s_bench.json | cut -c1-600
