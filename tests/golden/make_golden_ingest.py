"""Golden vectors for the audio ingest (utils/audio.py) from the REAL reference functions.
Run in the build container only (needs /root/reference and torchaudio): python tests/golden/make_golden_ingest.py"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from stt_server.utils.audio import ensure_16k, pcm16_to_float32  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def clip(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    x = 0.4 * np.sin(2 * np.pi * (40 + 300 * t) * t) + 0.2 * rng.standard_normal(n)
    x[: n // 10] = 0.0
    x[n // 2] = 1.5   # clips to full scale
    x[min(n - 1, n // 2 + 1)] = -1.5
    return (np.clip(x, -1.0, 32767 / 32768) * 32768).astype(np.int16)


out = {}
for rate, n in ((8000, 2000), (11025, 2756), (16000, 1600), (22050, 3307), (24000, 2400), (32000, 3200), (44100, 4410), (48000, 4800),
                (48000, 1), (8000, 3), (44100, 441)):
    pcm = clip(rate + n, n)
    f = pcm16_to_float32(pcm.tobytes())
    y = ensure_16k(f, rate)
    out[f"pcm_{rate}_{n}"] = pcm
    out[f"f32_{rate}_{n}"] = f
    out[f"y16k_{rate}_{n}"] = np.asarray(y, dtype=np.float32)
np.savez_compressed(os.path.join(HERE, "ingest.npz"), **out)
print("wrote", os.path.join(HERE, "ingest.npz"), {k: v.shape for k, v in out.items() if k.startswith("y16k")})
