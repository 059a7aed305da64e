mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_shell.py tests/test_gpu_models.py tests/test_gpu_train_oracle.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2q_pytest.log
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
MEDT_NO_FUSED_HEAD=1 BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline --configs '' > gpurun_out/r2q_bench_nohead.json 2> gpurun_out/r2q_bench_nohead.err
MT_DEBUG_FLAGS=64 MT_BWD_FUSED_ALL=1 timeout 300 python tools/time_ops.py --tag fused_all > gpurun_out/r2q_ops_all.json 2> gpurun_out/r2q_ops_all.err
tail -3 gpurun_out/r2q_pytest.log; python - <<'P'
import json
for f in ("r2q_bench","r2q_bench_nohead"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d.get("also",{}).items()})
d=json.load(open("gpurun_out/r2q_ops_all.json"))
print({k:(v.get("fwd_us_graph"),v.get("bwd_us_graph")) for k,v in d["ops"].items()})
P
