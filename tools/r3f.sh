mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_models.py::test_dataparallel_two_gpus_like_train_py tests/test_gpu_reference_script.py tests/test_gpu_shell.py tests/test_gpu_nccl.py -m gpu -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r3f_tests.log
