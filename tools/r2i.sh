mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_bwd -s 2 -c 1 -o gpurun_out/r2i_bwd_c32 python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2i_ncu.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_bwd -s 6 -c 1 -o gpurun_out/r2i_bwd_c64 python tools/time_ops.py --reps 1 > /dev/null 2>> gpurun_out/r2i_ncu.err
ls -la gpurun_out | grep r2i
