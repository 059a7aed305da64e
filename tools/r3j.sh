mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3j_bench_ref.json 2> gpurun_out/r3j_bench_ref.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3j_smoke.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r3_launches_gated_step_warm.csv python tools/profile_step.py gated > gpurun_out/r3j_prof_gated.log 2>&1; echo "launches gated rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3j_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d.get("gpu_launches_per_step"), {k:(v["ms_per_step"], round(v["value"],1)) for k,v in d.get("also",{}).items()})
print(d.get("cpu_baseline"))
print(open("gpurun_out/r3j_bench_ref.json").read()[:600])
PY
