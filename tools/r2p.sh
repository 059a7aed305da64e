mkdir -p gpurun_out
BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err
tail -c 1500 gpurun_out/r2p_bench_n2.json; tail -5 gpurun_out/r2p_bench_n2.err
(timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -k "dataparallel" 2>&1 | tail -5) > gpurun_out/r2p_pytest_dp.log; tail -3 gpurun_out/r2p_pytest_dp.log
