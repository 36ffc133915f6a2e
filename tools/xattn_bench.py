"""Decoder cross-attention kernel alone (CUDA events inside the library): ms per launch and GB/s of algorithmic bytes.
Usage: python tools/xattn_bench.py [model]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper.backend import B200WhisperBackend  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=128, max_sequences=320, max_encoder_batch=1)
mode = "default"
for seg, grp in ((128, 1), (64, 1), (32, 1), (8, 1), (1, 1), (64, 5), (16, 5)):
    ms, by = b.engine.bench_cross_attention(seg, grp, 64)
    print(f"xattn {mode} {seg:3d} x {grp}: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.1f} GB/s", flush=True)
