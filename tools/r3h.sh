mkdir -p gpurun_out
timeout 600 python bench.py --configs "" --no-cpu-baseline --steps 30 > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3h_bench.err | cut -c1-300
MEDT_LIB_PATH=/root/repo/medical-transformer_b200/medt_b200/libmedt_prev.so timeout 600 python bench.py --configs "" --no-cpu-baseline --steps 30 > gpurun_out/r3h_bench_prev.json 2> gpurun_out/r3h_bench_prev.err; echo "bench prev rc=$?"; tail -2 gpurun_out/r3h_bench_prev.err | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3h_tests.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag noinline > gpurun_out/r3h_ops.json 2> gpurun_out/r3h_ops.err; echo "ops rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3h_bench.json","gpurun_out/r3h_bench_prev.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("gpurun_out/r3h_ops.json"))
    print({k:(v.get("fwd_us_graph"),v.get("bwd_us_graph")) for k,v in d["ops"].items()})
    print(d["sum"])
except Exception as e: print("ops ERR", e)
PY
