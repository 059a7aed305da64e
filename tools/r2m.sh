mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2m_pytest_axial.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag fused > gpurun_out/r2m_ops_fused.json 2> gpurun_out/r2m_ops_fused.err
MT_DEBUG_FLAGS=64 MT_BWD_FUSED_RT=4 timeout 300 python tools/time_ops.py --tag fused_rt4 > gpurun_out/r2m_ops_rt4.json 2> gpurun_out/r2m_ops_rt4.err
MT_DEBUG_FLAGS=64 MT_BWD_FUSED_RT=1 timeout 300 python tools/time_ops.py --tag fused_rt1 > gpurun_out/r2m_ops_rt1.json 2> gpurun_out/r2m_ops_rt1.err
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag fused512 --px 512 > gpurun_out/r2m_ops512_fused.json 2> gpurun_out/r2m_ops512_fused.err
MT_DEBUG_FLAGS=64 MT_BWD_FUSED_RT=2 timeout 300 python tools/time_ops.py --tag fused512rt2 --px 512 > gpurun_out/r2m_ops512_rt2.json 2> gpurun_out/r2m_ops512_rt2.err
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_bwd -s 2 -c 1 -o gpurun_out/r2m_bwd_c16 python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2m_ncu.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_bwd -s 10 -c 1 -o gpurun_out/r2m_bwd_c32 python tools/time_ops.py --reps 1 > /dev/null 2>> gpurun_out/r2m_ncu.err
(timeout 1500 python -m pytest tests/test_gpu_reference_script.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2m_pytest_ref.log
tail -3 gpurun_out/r2m_pytest_axial.log; tail -3 gpurun_out/r2m_pytest_ref.log; python - <<'P'
import json
for f in ("r2m_ops_fused","r2m_ops_rt4","r2m_ops_rt1","r2m_ops512_fused","r2m_ops512_rt2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k:(v.get("fwd_us_graph"),v.get("bwd_us_graph")) for k,v in d["ops"].items()})
    except Exception as e: print(f, e)
d=json.load(open("gpurun_out/r2m_bench.json")); print(d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d["also"].items()})
P
