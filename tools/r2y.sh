mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_tc_launches.csv python tools/tc_shapes.py > gpurun_out/r2y.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2y.log
timeout 300 python tools/tc_bringup.py 1 > gpurun_out/r2y_acc.json 2> gpurun_out/r2y_acc.err; echo "bringup rc=$?"; cat gpurun_out/r2y_acc.json | cut -c1-2500
