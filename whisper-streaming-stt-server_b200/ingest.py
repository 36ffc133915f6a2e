"""Raw audio ingest: the polyphase filter bank `ensure_16k` uses, built on the host for the device resampler.

`ModelWorker._decode` (reference stt_server/model/worker.py:118-121) turns the received bytes into the backend's input
with `pcm16_to_float32` and `ensure_16k` (stt_server/utils/audio.py:6-30): int16 -> float32 / 32768, then
`torchaudio.functional.resample(orig, 16000, lowpass_filter_width=6)` (sinc_interp_hann, rolloff 0.99) when the
stream is not at 16 kHz.  `B200WhisperBackend.transcribe_pcm16` moves both steps onto the GPU; this module builds the
filter taps exactly as torchaudio's `_get_sinc_resample_kernel` does for a float32 waveform (float32 arithmetic
throughout), so the device kernel only has to evaluate the strided correlation.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

LOWPASS_FILTER_WIDTH = 6   # utils/audio.py:24
ROLLOFF = 0.99             # torchaudio default
TARGET_RATE = 16000


def resample_taps(sample_rate: int, target_rate: int = TARGET_RATE) -> Tuple[int, int, int, np.ndarray]:
    """(orig, new, width, taps[new, 2*width+orig] float32) for `sample_rate` -> `target_rate`."""
    if int(sample_rate) != sample_rate or sample_rate <= 0:
        raise ValueError(f"sample_rate must be a positive integer, got {sample_rate!r}")
    g = math.gcd(int(sample_rate), int(target_rate))
    orig, new = int(sample_rate) // g, int(target_rate) // g
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx
    t = t * f32(base_freq)
    t = np.clip(t, f32(-LOWPASS_FILTER_WIDTH), f32(LOWPASS_FILTER_WIDTH))
    window = np.cos(t * f32(math.pi) / f32(LOWPASS_FILTER_WIDTH) / f32(2)) ** 2
    t = t * f32(math.pi)
    scale = f32(base_freq / orig)
    with np.errstate(divide="ignore", invalid="ignore"):
        taps = np.where(t == 0, f32(1.0), np.sin(t) / t).astype(f32)
    taps = taps * (window * scale)
    return orig, new, width, np.ascontiguousarray(taps, dtype=np.float32)


def resampled_length(n_samples: int, sample_rate: int, target_rate: int = TARGET_RATE) -> int:
    """ceil(new * n / orig): what torchaudio keeps of the strided correlation."""
    g = math.gcd(int(sample_rate), int(target_rate))
    orig, new = int(sample_rate) // g, int(target_rate) // g
    return -(-new * int(n_samples) // orig)
