mkdir -p gpurun_out
# PDL A/B on one box (gated + MedT only)
for pdl in 0 1; do
  MT_PDL=$pdl timeout 600 python bench.py --configs "" --no-cpu-baseline --steps 30 > gpurun_out/r3g_bench_pdl$pdl.json 2> gpurun_out/r3g_bench_pdl$pdl.err; echo "pdl=$pdl rc=$?"
  tail -2 gpurun_out/r3g_bench_pdl$pdl.err | cut -c1-300
done
MT_PDL=1 timeout 900 python -m pytest tests/test_gpu_shell.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/r3g_tests_pdl1.log 2>&1; echo "tests pdl=1 rc=$?"; tail -3 gpurun_out/r3g_tests_pdl1.log
timeout 900 python -m pytest tests/test_gpu_shell.py -m gpu -x -q > gpurun_out/r3g_tests_shell.log 2>&1; echo "tests shell rc=$?"; tail -3 gpurun_out/r3g_tests_shell.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial -c 2 -o gpurun_out/r3e_c64_l16 python tools/ncu_one_op.py 64 4 16 16 0 1 > gpurun_out/r3e_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3g_bench_pdl0.json","gpurun_out/r3g_bench_pdl1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
        if "tensor_core_kernel" in d.get("roofline",{}): print(json.dumps(d["roofline"]["tensor_core_kernel"])[:900])
    except Exception as e: print(f, "ERR", e)
PY
