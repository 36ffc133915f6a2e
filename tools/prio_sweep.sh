export SWEEP_SHAPES=128x1,64x1,64x5 SWEEP_CTX=100
for cfg in "" "B200W_GROUPS=2" "B200W_GROUPS=2 B200W_PRIO=1" "B200W_GROUPS=3 B200W_PRIO=1" "B200W_GROUPS=4 B200W_PRIO=1" "B200W_PRIO=1"; do
  env $cfg python tools/step_sweep.py 2>&1 | grep "step"
done
