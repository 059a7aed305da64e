mkdir -p gpurun_out
for cfg in "rt4ks0:" "rt2ks0:MT_FORCE_RT=2" "rt2ks2:MT_FORCE_RT=2 MT_COOP_KS=2" "rt4ks1:MT_COOP_KS=1" "rt4ks4:MT_COOP_KS=4" "rt2ks4:MT_FORCE_RT=2 MT_COOP_KS=4" "rt1ks4:MT_FORCE_RT=1 MT_COOP_KS=4"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env MT_DEBUG_FLAGS=64 $envs timeout 200 python tools/time_ops.py --tag $tag --reps 10 > gpurun_out/r2d_$tag.json 2> gpurun_out/r2d_$tag.err
done
ls gpurun_out | grep r2d
