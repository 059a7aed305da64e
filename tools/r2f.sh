mkdir -p gpurun_out
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop > gpurun_out/r2f_ops_coop.json 2> gpurun_out/r2f_ops_coop.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_bwd -s 2 -c 1 -o gpurun_out/r2f_bwd python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2f_ncu.err
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --px 512 --reps 5 --tag coop512 > gpurun_out/r2f_ops512_coop.json 2> gpurun_out/r2f_ops512_coop.err
MT_DEBUG_FLAGS=32 timeout 300 python tools/time_ops.py --px 512 --reps 5 --tag chain512 > gpurun_out/r2f_ops512_chain.json 2> gpurun_out/r2f_ops512_chain.err
ls -la gpurun_out | grep r2f
