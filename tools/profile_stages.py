"""Runs one encoder batch and a few decoder steps of a random-init model so that ncu can list every launch.
Usage: python tools/profile_stages.py [model] [enc_batch] [segments] [n_group] [dec_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper.backend import B200WhisperBackend  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
enc_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
segments = int(sys.argv[3]) if len(sys.argv) > 3 else 128
n_group = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dec_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=segments, max_sequences=max(8, segments * n_group),
                       max_encoder_batch=enc_batch)
eng = b.engine
ms, fl = eng.bench_encoder(enc_batch, 1)
print(f"encoder batch {enc_batch}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
ms, by = eng.bench_decoder_step(segments, n_group, 100, dec_steps)
print(f"decoder step {segments}x{n_group}: {ms:.3f} ms, {by / ms / 1e6:.1f} GB/s")
ms, by = eng.bench_cross_attention(segments, n_group, 32)
print(f"cross attention: {ms:.4f} ms, {by / ms / 1e6:.1f} GB/s")
