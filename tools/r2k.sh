mkdir -p gpurun_out
(timeout 1700 python -m pytest tests/test_gpu_reference_script.py tests/test_gpu_train_oracle.py tests/test_gpu_shell.py -m gpu -q 2>&1 | tail -60) > gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
