mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r3_launches_gated_step_warm.csv python tools/profile_step.py gated > gpurun_out/r3d_prof_gated.log 2>&1; echo "launches gated rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r3_launches_medt_step_warm.csv python tools/profile_step.py MedT > gpurun_out/r3d_prof_medt.log 2>&1; echo "launches medt rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc3 -c 4 -o gpurun_out/r3_conv_tc3 python tools/tc_shapes.py > gpurun_out/r3d_ncu_tc3.log 2>&1; echo "ncu tc3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc3 -s 12 -c 4 -o gpurun_out/r3_conv_tc3_512 python tools/tc_shapes.py > gpurun_out/r3d_ncu_tc3b.log 2>&1; echo "ncu tc3 512 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:k_bn_act -c 4 -o gpurun_out/r3_bn_cluster python tools/profile_step.py gated > gpurun_out/r3d_ncu_bn.log 2>&1; echo "ncu bn rc=$?"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3d_tests.log
timeout 900 python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3d_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d.get("gpu_launches_per_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
print(d.get("cpu_baseline"))
PY
