mkdir -p gpurun_out
timeout 300 python tools/tc_bringup.py 1 > gpurun_out/r2z_acc.json 2> gpurun_out/r2z_acc.err; echo "bringup rc=$?"; cat gpurun_out/r2z_acc.json | cut -c1-1200; tail -2 gpurun_out/r2z_acc.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_tc_launches.csv python tools/tc_shapes.py > gpurun_out/r2z.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2z.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2z_tests.log
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2z_bench.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("gpu_launches_per_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
