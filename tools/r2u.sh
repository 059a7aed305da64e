mkdir -p gpurun_out
timeout 600 python tools/time_convs.py > gpurun_out/r2u_convs.json 2> gpurun_out/r2u_convs.err; echo "convs rc=$?"; cat gpurun_out/r2u_convs.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2u_tests.log
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2u_bench.json
MEDT_OWN_CONV=all timeout 600 python bench.py > gpurun_out/r2u_bench_all.json 2> gpurun_out/r2u_bench_all.err; echo "bench all rc=$?"
cut -c1-400 gpurun_out/r2u_bench_all.json
python - <<'PY'
import json
for f in ("gpurun_out/r2u_bench.json","gpurun_out/r2u_bench_all.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
