mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial -c 2 -o gpurun_out/r3e_c64_l16 python tools/ncu_one_op.py 64 4 16 16 0 1 > gpurun_out/r3e_ncu.log 2>&1; echo "rc=$?"
timeout 600 python bench.py --configs "" > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3e_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
print(json.dumps(d["roofline"].get("tensor_core_kernel"))[:1200])
PY
