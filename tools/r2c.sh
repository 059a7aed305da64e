mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_axial.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2c_pytest.log
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop_ph > gpurun_out/r2c_ops_coop.json 2> gpurun_out/r2c_ops_coop.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_fwd -s 3 -c 1 -o gpurun_out/r2c_fwd python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2c_ncu2.err
tail -2 gpurun_out/r2c_pytest.log
