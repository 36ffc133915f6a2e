"""Encoder forward alone (CUDA events inside the library) over batch sizes.
Usage: [B200W_NO_VEC8=1] python tools/encoder_bench.py [model] [batches...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper.backend import B200WhisperBackend  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
batches = [int(v) for v in sys.argv[2:]] or [16, 8, 4, 1]
b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=32, max_sequences=32, max_encoder_batch=max(batches))
tag = "".join(f" {k[6:]}={os.environ[k]}" for k in ("B200W_NO_VEC8", "B200W_ENC_NO_SPLIT") if os.environ.get(k))
for n in batches:
    b.engine.bench_encoder(n, 2)
    ms, fl = b.engine.bench_encoder(n, 10)
    print(f"encoder{tag} batch {n:2d}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  ({fl / ms / 1e9 / 1371.1 * 100:4.1f} % of sustained bf16)", flush=True)
