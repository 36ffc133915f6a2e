"""Slaney mel filterbank (what upstream ships as `assets/mel_filters.npz`, generated there with
`librosa.filters.mel(sr=16000, n_fft=400, n_mels=80|128)`); the asset is not in the image."""
from __future__ import annotations

import numpy as np


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    lin = f / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int, sr: int = 16000, n_fft: int = 400) -> np.ndarray:
    """float32 [n_mels, n_fft // 2 + 1], area-normalised triangles on the Slaney mel scale."""
    n_bins = n_fft // 2 + 1
    fft_f = np.linspace(0.0, sr / 2.0, n_bins)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    width = np.diff(pts)
    out = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        rise = (fft_f - pts[i]) / width[i]
        fall = (pts[i + 2] - fft_f) / width[i + 1]
        out[i] = np.maximum(0.0, np.minimum(rise, fall))
    out *= (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return out
