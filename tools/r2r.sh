mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_axial -c 2 -o gpurun_out/r2r_c32_l64 python tools/ncu_one_op.py 32 4 64 64 0 1 > gpurun_out/r2r_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial -c 2 -o gpurun_out/r2r_c16_l256 python tools/ncu_one_op.py 16 4 256 256 0 1 >> gpurun_out/r2r_ncu.log 2>&1
tail -4 gpurun_out/r2r_ncu.log; ls -la gpurun_out | grep r2r
