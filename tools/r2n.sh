mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_shell.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2n_pytest_shell.log
timeout 300 python tools/time_convs.py > gpurun_out/r2n_convs.json 2> gpurun_out/r2n_convs.err
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
(timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_train_oracle.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2n_pytest_models.log
tail -3 gpurun_out/r2n_pytest_shell.log; tail -3 gpurun_out/r2n_pytest_models.log; cat gpurun_out/r2n_convs.json; tail -3 gpurun_out/r2n_convs.err; python - <<'P'
import json
d=json.load(open("gpurun_out/r2n_bench.json")); print(d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d["also"].items()})
P
