mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_shell.py -m gpu -q > gpurun_out/r3b_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3b_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; echo "bench rc=$?"
MEDT_NO_WGRAD_STREAM=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3b_bench_nows.json 2> gpurun_out/r3b_bench_nows.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3b_bench.json","gpurun_out/r3b_bench_nows.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("gpu_launches_per_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
