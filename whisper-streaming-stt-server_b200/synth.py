"""Model dimension table, seeded random-init weights and synthetic speech-like audio.

There is no network for checkpoints or datasets, so parity tests and the bench use
random-init weights of the named architecture (openai state-dict key layout, the one
`whisper.load_model` in reference `stt_server/model/backends/torch_whisper.py:21` yields)
and synthetic audio.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np


@dataclass(frozen=True)
class ModelDims:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


def _d(n_mels, d, h, layers, vocab) -> ModelDims:
    return ModelDims(n_mels, 1500, d, h, layers, vocab, 448, d, h, layers)


# openai-whisper model zoo
MODEL_DIMS: Dict[str, ModelDims] = {
    "tiny.en": _d(80, 384, 6, 4, 51864),
    "tiny": _d(80, 384, 6, 4, 51865),
    "base.en": _d(80, 512, 8, 6, 51864),
    "base": _d(80, 512, 8, 6, 51865),
    "small.en": _d(80, 768, 12, 12, 51864),
    "small": _d(80, 768, 12, 12, 51865),
    "medium.en": _d(80, 1024, 16, 24, 51864),
    "medium": _d(80, 1024, 16, 24, 51865),
    "large-v1": _d(80, 1280, 20, 32, 51865),
    "large-v2": _d(80, 1280, 20, 32, 51865),
    "large-v3": _d(128, 1280, 20, 32, 51866),
    "large": _d(128, 1280, 20, 32, 51866),
    # 2-layer toys for fast tests (not in the zoo)
    "test-tiny": ModelDims(80, 1500, 128, 2, 2, 51865, 448, 128, 2, 2),
    "test-tiny.en": ModelDims(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2),
    "test-v3": ModelDims(128, 1500, 128, 2, 2, 51866, 448, 128, 2, 2),
}


def state_dict_spec(dims: ModelDims):
    """(name, shape, init kind, fan_in) for every tensor of the openai layout."""
    spec = []
    d = dims.n_audio_state
    spec.append(("encoder.conv1.weight", (d, dims.n_mels, 3), "u", dims.n_mels * 3))
    spec.append(("encoder.conv1.bias", (d,), "u", dims.n_mels * 3))
    spec.append(("encoder.conv2.weight", (d, d, 3), "u", d * 3))
    spec.append(("encoder.conv2.bias", (d,), "u", d * 3))

    def block(p, d, cross):
        names = ["attn"] + (["cross_attn"] if cross else [])
        for a in names:
            spec.append((f"{p}.{a}.query.weight", (d, d), "u", d))
            spec.append((f"{p}.{a}.query.bias", (d,), "u", d))
            spec.append((f"{p}.{a}.key.weight", (d, d), "u", d))
            spec.append((f"{p}.{a}.value.weight", (d, d), "u", d))
            spec.append((f"{p}.{a}.value.bias", (d,), "u", d))
            spec.append((f"{p}.{a}.out.weight", (d, d), "u", d))
            spec.append((f"{p}.{a}.out.bias", (d,), "u", d))
            spec.append((f"{p}.{a}_ln.weight", (d,), "ln_w", 0))
            spec.append((f"{p}.{a}_ln.bias", (d,), "ln_b", 0))
        spec.append((f"{p}.mlp.0.weight", (4 * d, d), "u", d))
        spec.append((f"{p}.mlp.0.bias", (4 * d,), "u", d))
        spec.append((f"{p}.mlp.2.weight", (d, 4 * d), "u", 4 * d))
        spec.append((f"{p}.mlp.2.bias", (d,), "u", 4 * d))
        spec.append((f"{p}.mlp_ln.weight", (d,), "ln_w", 0))
        spec.append((f"{p}.mlp_ln.bias", (d,), "ln_b", 0))

    for i in range(dims.n_audio_layer):
        block(f"encoder.blocks.{i}", d, False)
    spec.append(("encoder.ln_post.weight", (d,), "ln_w", 0))
    spec.append(("encoder.ln_post.bias", (d,), "ln_b", 0))
    dt = dims.n_text_state
    spec.append(("decoder.token_embedding.weight", (dims.n_vocab, dt), "n1", 0))
    spec.append(("decoder.positional_embedding", (dims.n_text_ctx, dt), "pos", 0))
    for i in range(dims.n_text_layer):
        block(f"decoder.blocks.{i}", dt, True)
    spec.append(("decoder.ln.weight", (dt,), "ln_w", 0))
    spec.append(("decoder.ln.bias", (dt,), "ln_b", 0))
    return spec


def random_state_items(dims: ModelDims, seed: int = 0, device: str = "cpu", eot_bias: float = 0.0, emb_std: float = 1.0):
    """Generator over (name, tensor) in `state_dict_spec` order; `random_state_dict` is dict() of it.  Consumers that
    load one tensor at a time (Engine._load) never hold more than one tensor on the host (large-v3: 265 MB instead of
    6.2 GB per process, which matters with one process per GPU)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    eot_row = None
    for name, shape, kind, fan_in in state_dict_spec(dims):
        if kind == "u":
            bound = 1.0 / np.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        elif kind == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif kind == "ln_b":
            t = 0.05 * torch.randn(shape, generator=g, device=device)
        elif kind == "n1":
            t = emb_std * torch.randn(shape, generator=g, device=device)
        elif kind == "pos":
            t = 0.01 * torch.randn(shape, generator=g, device=device)
        else:  # pragma: no cover
            raise AssertionError(kind)
        if eot_bias > 0:
            if name == "decoder.token_embedding.weight":
                eot_row = t[50257 if dims.n_vocab >= 51865 else 50256].clone()
            elif name == "decoder.ln.bias":  # comes after the embedding in the spec
                t = t + eot_bias * eot_row / (eot_row * eot_row).sum()
        yield name, t


class LazyRandomState:
    """dict-like view (`items()` only) of a seeded random checkpoint that materialises one tensor at a time."""

    def __init__(self, dims: ModelDims, seed: int = 0, eot_bias: float = 0.0, emb_std: float = 1.0):
        self.dims, self.seed, self.eot_bias, self.emb_std = dims, seed, eot_bias, emb_std

    def items(self):
        return random_state_items(self.dims, self.seed, eot_bias=self.eot_bias, emb_std=self.emb_std)


def random_state_dict(dims: ModelDims, seed: int = 0, device: str = "cpu", eot_bias: float = 0.0,
                      emb_std: float = 1.0):
    """Seeded random init: torch-default distributions (Linear/Conv U(+-1/sqrt(fan_in)), LayerNorm
    1/0 perturbed so the affine is exercised, Embedding N(0, emb_std^2); a small emb_std breaks the tied-embedding
    self-loop that makes a random model repeat one token), positional embedding N(0, 0.01^2).

    `eot_bias` c > 0 adds c * E[eot] / |E[eot]|^2 to `decoder.ln.bias`, i.e. a constant +c on the
    end-of-text logit (exercises the early-exit path; 0 keeps the no-EOT worst case: 224 steps per window)."""
    return dict(random_state_items(dims, seed, device=device, eot_bias=eot_bias, emb_std=emb_std))


def synth_audio(seed: int, seconds: float, sample_rate: int = 16000) -> np.ndarray:
    """Speech-like bursts (3-6 harmonics of f0~U(90,250) Hz, 4 Hz AM) + pink-ish noise at -30 dB,
    separated by silences; quantised through int16 like `pcm16_to_float32`
    (reference `stt_server/utils/audio.py:6-8`).  Returns float32 in [-1, 1)."""
    rng = np.random.default_rng(1234 + seed)
    n = int(round(seconds * sample_rate))
    t = np.arange(n) / sample_rate
    x = np.zeros(n, dtype=np.float64)
    pos = 0.0
    while pos < seconds:
        dur = rng.uniform(1.0, 8.0)
        gap = rng.uniform(0.6, 1.2)
        a, b = int(pos * sample_rate), min(n, int((pos + dur) * sample_rate))
        if b > a:
            f0 = rng.uniform(90, 250)
            seg_t = t[a:b]
            burst = np.zeros(b - a)
            for h in range(1, rng.integers(3, 7) + 1):
                burst += rng.uniform(0.3, 1.0) / h * np.sin(2 * np.pi * f0 * h * seg_t + rng.uniform(0, 2 * np.pi))
            burst *= 0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * seg_t)
            x[a:b] += burst
        pos += dur + gap
    white = rng.standard_normal(n)
    pink = np.cumsum(white) * 0.02 + white  # crude 1/f tilt
    pink -= pink.mean()
    peak = np.max(np.abs(x)) or 1.0
    x = 0.5 * x / peak + 10 ** (-30 / 20) * pink / (np.max(np.abs(pink)) or 1.0)
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    return pcm.astype(np.float32) / 32768.0


def synth_pcm16(seed: int, seconds: float) -> bytes:
    a = synth_audio(seed, seconds)
    return np.clip(np.round(a * 32768.0), -32768, 32767).astype(np.int16).tobytes()
