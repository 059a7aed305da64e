mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2l_pytest_axial.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag fused > gpurun_out/r2l_ops_fused.json 2> gpurun_out/r2l_ops_fused.err
MT_DEBUG_FLAGS=64 MT_BWD_NOFUSE=1 timeout 300 python tools/time_ops.py --tag nofuse > gpurun_out/r2l_ops_nofuse.json 2> gpurun_out/r2l_ops_nofuse.err
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag fused512 --px 512 > gpurun_out/r2l_ops512_fused.json 2> gpurun_out/r2l_ops512_fused.err
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
(timeout 1500 python -m pytest tests/test_gpu_reference_script.py tests/test_gpu_train_oracle.py -m gpu -q --tb=short 2>&1 | tail -150) > gpurun_out/r2l_pytest_new.log
tail -3 gpurun_out/r2l_pytest_axial.log; tail -3 gpurun_out/r2l_pytest_new.log; python - <<'P'
import json
for f in ("r2l_ops_fused","r2l_ops_nofuse","r2l_ops512_fused"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k:(v.get("fwd_us_graph"),v.get("bwd_us_graph")) for k,v in d["ops"].items()})
    except Exception as e: print(f, e)
d=json.load(open("gpurun_out/r2l_bench.json")); print(d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d["also"].items()})
P
