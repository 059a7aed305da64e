mkdir -p gpurun_out
for m in 1 2 0; do
  timeout 300 python tools/tc_bringup.py $m > gpurun_out/r2t_mode$m.json 2> gpurun_out/r2t_mode$m.err; echo "mode $m rc=$?"
  cat gpurun_out/r2t_mode$m.json; tail -3 gpurun_out/r2t_mode$m.err
done
