python tools/xattn_bench.py 2>&1 | grep xattn
B200W_XATTN_PERSIST=1 python tools/xattn_bench.py 2>&1 | grep xattn
export SWEEP_SHAPES=128x1,64x1,64x5 SWEEP_CTX=100
for cfg in "B200W_XATTN_PERSIST=1" "B200W_XATTN_PERSIST=1 B200W_GROUPS=2" "B200W_XATTN_PERSIST=1 B200W_GROUPS=2 B200W_PRIO=1" "B200W_XATTN_PERSIST=1 B200W_GROUPS=3"; do
  echo "== $cfg"; env $cfg python tools/step_sweep.py 2>&1 | grep "step"
done
B200W_XATTN_PERSIST=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or large_v3 or concurrent or real_model" 2>&1 | tail -5
