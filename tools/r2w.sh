mkdir -p gpurun_out
timeout 300 python tools/tc_bringup.py 1 > gpurun_out/r2w_bn16.json 2> gpurun_out/r2w_bn16.err; echo "bn16 rc=$?"
MT_TC_MIN_BN=32 timeout 300 python tools/tc_bringup.py 1 > gpurun_out/r2w_bn32.json 2> gpurun_out/r2w_bn32.err; echo "bn32 rc=$?"
python - <<'PY'
import json
a=json.load(open("gpurun_out/r2w_bn16.json"))["res"]; b=json.load(open("gpurun_out/r2w_bn32.json"))["res"]
for k in a: print(k, {m:(round(a[k][m]*1e6,2), round(b[k][m]*1e6,2)) for m in a[k]})
PY
timeout 900 python tools/time_convs.py > gpurun_out/r2w_convs.json 2> gpurun_out/r2w_convs.err; echo "convs rc=$?"; tail -3 gpurun_out/r2w_convs.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2w_tests.log
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
MEDT_OWN_CONV=strided MT_BN_CLUSTER=0 timeout 600 python bench.py > gpurun_out/r2w_bench_old.json 2> gpurun_out/r2w_bench_old.err; echo "bench old rc=$?"
MEDT_OWN_CONV=strided timeout 600 python bench.py --configs "" --no-cpu-baseline > gpurun_out/r2w_bench_bnonly.json 2> gpurun_out/r2w_bench_bnonly.err; echo "bench bnonly rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2w_bench.json","gpurun_out/r2w_bench_old.json","gpurun_out/r2w_bench_bnonly.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("gpu_launches_per_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
