mkdir -p gpurun_out
MT_DEBUG_FLAGS=64 timeout 300 python tools/time_ops.py --tag coop_ph > gpurun_out/r2b_ops_coop.json 2> gpurun_out/r2b_ops_coop.err
MT_DEBUG_FLAGS=32 timeout 300 python tools/time_ops.py --tag chain > gpurun_out/r2b_ops_chain.json 2> gpurun_out/r2b_ops_chain.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_axial_fwd -c 60 --csv --log-file gpurun_out/r2b_launches.csv python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2b_ncu1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_axial_fwd -s 3 -c 2 -o gpurun_out/r2b_fwd python tools/time_ops.py --reps 1 > /dev/null 2> gpurun_out/r2b_ncu2.err
ls -la gpurun_out | tail -8
