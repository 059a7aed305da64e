mkdir -p gpurun_out
timeout 300 python tools/tc_bringup.py 1 > gpurun_out/r2v_mode1.json 2> gpurun_out/r2v_mode1.err; echo "mode 1 rc=$?"
cat gpurun_out/r2v_mode1.json; tail -3 gpurun_out/r2v_mode1.err
timeout 900 python tools/time_convs.py > gpurun_out/r2v_convs.json 2> gpurun_out/r2v_convs.err; echo "convs rc=$?"; tail -3 gpurun_out/r2v_convs.err
