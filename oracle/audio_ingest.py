"""CPU restatement of the reference's audio ingest in front of the backend (TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product).

Reference: /root/reference/stt_server/utils/audio.py
  * pcm16_to_float32 (:6-8)   np.frombuffer(int16).astype(float32) / 32768.0
  * ensure_16k       (:11-30) identity at 16 kHz, else torchaudio.functional.resample(orig, 16000,
                              lowpass_filter_width=6) -- rolloff 0.99, sinc_interp_hann defaults.
torchaudio (pinned 2.x in requirements-lock.txt) is third-party; its published algorithm
(`_get_sinc_resample_kernel` / `_apply_sinc_resample_kernel`) is restated below in numpy float32.
PARITY PINNED: tests/golden/ingest.npz holds outputs of the REAL reference functions (the module imports in the
build container, torchaudio 2.11 is installed there); tests/golden/make_golden_ingest.py generated them.
"""
from __future__ import annotations

import math

import numpy as np


def pcm16_to_float32(pcm_bytes: bytes) -> np.ndarray:
    """utils/audio.py:6-8"""
    return np.frombuffer(pcm_bytes, dtype=np.int16).astype(np.float32) / 32768.0


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio `_get_sinc_resample_kernel` (sinc_interp_hann) evaluated in float32 like resample() does for a
    float32 waveform.  Returns (kernels [new, 2*width+orig] float32, width, orig, new) with orig/new reduced by gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx
    t = t * f32(base_freq)
    t = np.clip(t, f32(-lowpass_filter_width), f32(lowpass_filter_width))
    window = np.cos(t * f32(math.pi) / f32(lowpass_filter_width) / f32(2)) ** 2
    t = t * f32(math.pi)
    scale = f32(base_freq / orig)
    with np.errstate(divide="ignore", invalid="ignore"):
        kernels = np.where(t == 0, f32(1.0), np.sin(t) / t).astype(f32)
    kernels = kernels * (window * scale)
    return kernels.astype(f32), width, orig, new


def ensure_16k(audio: np.ndarray, src_rate: int) -> np.ndarray:
    """utils/audio.py:11-30 (+ torchaudio `_apply_sinc_resample_kernel`: zero pad (width, width+orig), strided
    correlation with the `new` polyphase filters, keep ceil(new * length / orig) samples)."""
    if src_rate == 16000:
        return audio
    kernels, width, orig, new = sinc_resample_kernel(src_rate, 16000)
    x = np.asarray(audio, dtype=np.float32)
    length = x.shape[0]
    xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + orig, np.float32)])
    K = kernels.shape[1]
    n_frames = (xp.shape[0] - K) // orig + 1
    # frames[n, k] = xp[n*orig + k]
    frames = np.lib.stride_tricks.as_strided(xp, shape=(n_frames, K), strides=(xp.strides[0] * orig, xp.strides[0]))
    out = (frames.astype(np.float32) @ kernels.T.astype(np.float32)).reshape(-1)  # [n_frames, new] -> interleaved
    target = int(math.ceil(new * length / orig))
    return np.ascontiguousarray(out[:target], dtype=np.float32)


def ingest(pcm_bytes: bytes, sample_rate: int) -> np.ndarray:
    """What ModelWorker._decode feeds the backend (worker.py:118-121): bytes -> float32 -> 16 kHz."""
    return ensure_16k(pcm16_to_float32(pcm_bytes), sample_rate)
