mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_models.py::test_dataparallel_two_gpus_like_train_py tests/test_gpu_reference_script.py tests/test_gpu_nccl.py -m gpu -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r3f_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3f_n2.json 2> gpurun_out/r3f_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/r3f_n2.err | cut -c1-200
timeout 600 python bench.py --steps 20 --configs "" --no-cpu-baseline > gpurun_out/r3f_n1.json 2> gpurun_out/r3f_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3f_n1.json","gpurun_out/r3f_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["config"].get("replicas_lock_step"), {k:v["ms_per_step"] for k,v in d.get("also",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
