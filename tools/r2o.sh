mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_shell.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2o_pytest_shell.log
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
(timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2o_pytest_models.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2o_launches_warm.csv python tools/profile_step.py > /dev/null 2> gpurun_out/r2o_ncu.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2o_launches_medt_warm.csv python tools/profile_step.py MedT > /dev/null 2>> gpurun_out/r2o_ncu.err
tail -3 gpurun_out/r2o_pytest_shell.log; tail -3 gpurun_out/r2o_pytest_models.log; python - <<'P'
import json
d=json.load(open("gpurun_out/r2o_bench.json")); print(d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d["also"].items()})
P
