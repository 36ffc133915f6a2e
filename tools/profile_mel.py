"""Runs the log-mel kernels alone (one 6 s call per iteration) so that ncu can capture them.
Usage: python tools/profile_mel.py [model] [seconds] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper.backend import B200WhisperBackend  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=4, max_sequences=8, max_encoder_batch=1)
ms, by = b.engine.bench_mel(int(seconds * 16000), iters)
print(f"log-mel of a {seconds:.1f} s call: {ms * 1e3:.1f} us, {by / ms / 1e6:.1f} GB/s of algorithmic bytes")
