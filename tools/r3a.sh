mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r3a_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
MEDT_NO_WGRAD_STREAM=1 timeout 600 python bench.py --no-cpu-baseline --configs "" > gpurun_out/r3a_bench_nows.json 2> gpurun_out/r3a_bench_nows.err; echo "bench rc=$?"
MEDT_OWN_WGRAD=1 timeout 600 python bench.py --no-cpu-baseline --configs "" > gpurun_out/r3a_bench_ownw.json 2> gpurun_out/r3a_bench_ownw.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3a_bench.json","gpurun_out/r3a_bench_nows.json","gpurun_out/r3a_bench_ownw.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("gpu_launches_per_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
