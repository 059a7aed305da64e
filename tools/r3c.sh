mkdir -p gpurun_out
for mode in overlap plain; do
  if [ $mode = plain ]; then export MEDT_NO_OVERLAP_REDUCE=1; else unset MEDT_NO_OVERLAP_REDUCE; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --configs "" --no-cpu-baseline > gpurun_out/r3c_n2_$mode.json 2> gpurun_out/r3c_n2_$mode.err; echo "$mode rc=$?"
  tail -2 gpurun_out/r3c_n2_$mode.err | cut -c1-300
done
unset MEDT_NO_OVERLAP_REDUCE
timeout 600 python bench.py --steps 30 --configs "" --no-cpu-baseline > gpurun_out/r3c_n1.json 2> gpurun_out/r3c_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3c_n1.json","gpurun_out/r3c_n2_overlap.json","gpurun_out/r3c_n2_plain.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["config"].get("replicas_lock_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
