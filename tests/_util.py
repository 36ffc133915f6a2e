"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
from __future__ import annotations

import functools

import numpy as np

from b200_whisper.synth import MODEL_DIMS, random_state_dict
from oracle import whisper_oracle as wo

REALTIME = {"beam_size": 1, "best_of": 1, "patience": 1.0, "temperature": 0.0, "length_penalty": 1.0,
            "without_timestamps": True, "compression_ratio_threshold": 2.4, "no_speech_threshold": 0.6,
            "log_prob_threshold": -1.0}
ACCURATE = dict(REALTIME, beam_size=5, best_of=5)


@functools.lru_cache(maxsize=8)
def oracle_model(name: str, seed: int = 0, emb_std: float = 0.1, eot_bias: float = 0.0) -> wo.Whisper:
    dims = MODEL_DIMS[name]
    return wo.Whisper(wo.ModelDimensions(**dims.__dict__), random_state_dict(dims, seed, emb_std=emb_std, eot_bias=eot_bias))


def model_spec(name: str, seed: int = 0, emb_std: float = 0.1, eot_bias: float = 0.0) -> str:
    return f"random:{name}:{seed}:{emb_std}:{eot_bias}"


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / (np.linalg.norm(b.astype(np.float64)) + 1e-30))
