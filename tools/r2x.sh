mkdir -p gpurun_out
timeout 900 python tools/time_convs.py > gpurun_out/r2x_convs.json 2> gpurun_out/r2x_convs.err; echo "convs rc=$?"; tail -3 gpurun_out/r2x_convs.err
