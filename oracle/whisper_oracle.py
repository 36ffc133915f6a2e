"""CPU fp32 restatement of the reference's torch_whisper path (TEST INFRASTRUCTURE).

What it follows
---------------
* in-tree wrapper `stt_server/model/backends/torch_whisper.py:49-110` (option
  normalisation, fp16 rule, dict -> Segment mapping) and the caller
  `stt_server/model/worker.py:98-146`, `stt_server/utils/audio.py:6-8`;
* upstream `openai-whisper==20250625` (`requirements-lock.txt:48`), NOT vendored:
  `audio.py` (log_mel_spectrogram, pad_or_trim), `model.py` (AudioEncoder,
  TextDecoder, MultiHeadAttention w/ SDPA), `decoding.py` (DecodingTask, logit
  filters, GreedyDecoder, BeamSearchDecoder, MaximumLikelihoodRanker,
  detect_language), `transcribe.py` (seek loop / segment assembly).  Restated
  from the published algorithm.

PARITY UNPINNED (see `oracle/__init__.py`).  Deviations, all explicit:
* no tokenizer assets -> text is rendered by `render_text` (placeholder) and the
  suppress tables come from `oracle/tables.py`;
* `decoder.positional_embedding` is `torch.empty` upstream; callers must supply it;
* word timestamps are not restated (DTW alignment; out of scope);
* temperature > 0: upstream draws `Categorical(logits / T).sample()` from torch's global
  generator, which no other implementation can reproduce.  With `sample_seed=None` the
  oracle does exactly that; with an integer seed it draws the same distribution by
  Gumbel-max from the counter-based generator `gumbel_noise` below, which is OUR
  definition (restated bit for bit by `sample_uniform` in csrc/sampling.cu) so that the
  device path can be checked token for token.  The fallback ladder itself
  (`decode_with_fallback`) follows upstream transcribe.py.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .tables import LANGUAGES, TokenLayout, layout_for_vocab

# ---- upstream audio.py constants (Appendix A.1) ---------------------------------------
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH  # 100


@dataclass(frozen=True)
class ModelDimensions:
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


# ---- mel frontend ---------------------------------------------------------------------
def mel_filters(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) (Slaney scale + Slaney norm, fmin 0, fmax sr/2);
    regenerates upstream's `assets/mel_filters.npz` entries `mel_80` / `mel_128`."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def log_mel_spectrogram(audio: np.ndarray | torch.Tensor, n_mels: int, padding: int = 0) -> torch.Tensor:
    """upstream audio.py log_mel_spectrogram (Appendix A.2) -> f32 [n_mels, (n+padding)//160]."""
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    audio = audio.to(torch.float32)
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    filters = torch.from_numpy(mel_filters(n_mels))
    mel_spec = filters @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec


def pad_or_trim(array: torch.Tensor, length: int = N_FRAMES, axis: int = -1) -> torch.Tensor:
    if array.shape[axis] > length:
        array = array.index_select(dim=axis, index=torch.arange(length))
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = F.pad(array, [p for sizes in pad[::-1] for p in sizes])
    return array


# ---- model (upstream model.py, Appendix A.3) ------------------------------------------
def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    assert channels % 2 == 0
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


class Whisper:
    """Functional model over an openai-layout state dict (fp32, CPU)."""

    def __init__(self, dims: ModelDimensions, state: Dict[str, torch.Tensor]):
        self.dims = dims
        self.w = {k: v.detach().to(torch.float32) for k, v in state.items()}
        self.layout = layout_for_vocab(dims.n_vocab)
        if "encoder.positional_embedding" not in self.w:
            self.w["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, dims.n_audio_state)

    @property
    def is_multilingual(self) -> bool:
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self) -> int:
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)

    # -- building blocks
    def _ln(self, x, p):
        return F.layer_norm(x.float(), (x.shape[-1],), self.w[p + ".weight"], self.w[p + ".bias"])

    def _lin(self, x, p, bias=True):
        return F.linear(x, self.w[p + ".weight"], self.w[p + ".bias"] if bias else None)

    @staticmethod
    def _sdpa(q, k, v, n_head, causal):
        q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        return a.permute(0, 2, 1, 3).flatten(start_dim=2)

    def _attn(self, x, p, n_head, xa=None, cache=None, causal_mask=False):
        """MultiHeadAttention.forward with the kv-cache hook semantics: self-attn K/V are
        appended along time; cross-attn K/V are computed once and reused."""
        q = self._lin(x, p + ".query")
        if xa is None:
            k = self._lin(x, p + ".key", bias=False)
            v = self._lin(x, p + ".value")
            if cache is not None:
                if p + ".key" in cache:
                    k = torch.cat([cache[p + ".key"], k], dim=1)
                    v = torch.cat([cache[p + ".value"], v], dim=1)
                cache[p + ".key"], cache[p + ".value"] = k, v
        else:
            if cache is not None and p + ".key" in cache:
                k, v = cache[p + ".key"], cache[p + ".value"]
            else:
                k = self._lin(xa, p + ".key", bias=False)
                v = self._lin(xa, p + ".value")
                if cache is not None:
                    cache[p + ".key"], cache[p + ".value"] = k, v
        causal = causal_mask and q.shape[1] > 1
        return self._lin(self._sdpa(q, k, v, n_head, causal), p + ".out")

    def _block(self, x, p, n_head, xa=None, cache=None, causal=False):
        x = x + self._attn(self._ln(x, p + ".attn_ln"), p + ".attn", n_head, cache=cache, causal_mask=causal)
        if xa is not None:
            x = x + self._attn(self._ln(x, p + ".cross_attn_ln"), p + ".cross_attn", n_head, xa=xa, cache=cache)
        h = F.gelu(self._lin(self._ln(x, p + ".mlp_ln"), p + ".mlp.0"))
        return x + self._lin(h, p + ".mlp.2")

    # -- AudioEncoder.forward
    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel f32 [B, n_mels, 3000] -> [B, 1500, d]."""
        d = self.dims
        x = F.gelu(F.conv1d(mel, self.w["encoder.conv1.weight"], self.w["encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, self.w["encoder.conv2.weight"], self.w["encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1)
        assert x.shape[1:] == self.w["encoder.positional_embedding"].shape, "incorrect audio shape"
        x = x + self.w["encoder.positional_embedding"]
        for i in range(d.n_audio_layer):
            x = self._block(x, f"encoder.blocks.{i}", d.n_audio_head)
        return self._ln(x, "encoder.ln_post")

    # -- TextDecoder.forward
    def decode(self, tokens: torch.Tensor, xa: torch.Tensor, cache: Optional[dict] = None) -> torch.Tensor:
        """tokens i64 [B, n] (+cache) -> logits f32 [B, n, V]."""
        d = self.dims
        offset = 0
        if cache:
            offset = cache["decoder.blocks.0.attn.key"].shape[1]
        emb = self.w["decoder.token_embedding.weight"]
        x = F.embedding(tokens, emb) + self.w["decoder.positional_embedding"][offset : offset + tokens.shape[-1]]
        for i in range(d.n_text_layer):
            x = self._block(x, f"decoder.blocks.{i}", d.n_text_head, xa=xa, cache=cache, causal=True)
        x = self._ln(x, "decoder.ln")
        return (x @ emb.transpose(0, 1)).float()


# ---- decoding (upstream decoding.py, Appendix A.4) ------------------------------------
@dataclass
class DecodingOptions:
    task: str = "transcribe"
    language: Optional[str] = None
    temperature: float = 0.0
    sample_len: Optional[int] = None
    best_of: Optional[int] = None
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Sequence[int]] = None
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    sample_seed: Optional[int] = None  # not upstream: see the module docstring (temperature > 0)


_M64 = (1 << 64) - 1


def window_seed(base_seed: int, seek: int, attempt: int) -> int:
    """Seed of one decode attempt (window at `seek`, rung `attempt` of the temperature ladder)."""
    return (base_seed + 0x632BE59BD9B4E019 * (seek + 1) + 0xD1342543DE82EF95 * (attempt + 1)) & _M64


def gumbel_noise(seed: int, stream: int, pos: int, n_vocab: int) -> np.ndarray:
    """-log(-log(u)) for every token id, u = the 23-bit uniform of (seed, hypothesis, position, id):
    splitmix64 finaliser over the packed key.  float64 here; the device evaluates the logs in fp32."""
    ids = np.arange(n_vocab, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64((seed + 0x9E3779B97F4A7C15 * (stream + 1)) & _M64)
        z = z ^ ((np.uint64(pos) << np.uint64(32)) | ids)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(41)).astype(np.float64) + 0.5) / 8388608.0
    return -np.log(-np.log(u))


@dataclass
class DecodingResult:
    language: str
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    temperature: float = float("nan")
    compression_ratio: float = float("nan")
    sum_logprob: float = float("nan")
    # diagnostics for the parity tests: min top1-top2 margin of the filtered logits
    # over all sampled positions of the selected hypothesis' decode.
    min_margin: float = float("inf")


HAS_TEXT = False  # no tokenizer rank file in this image: text is a placeholder, see render_text


def render_text(tokens: Sequence[int], eot: int) -> str:
    """Placeholder detokeniser (vocab assets absent): text tokens as `<id>`."""
    return "".join(f"<{t}>" for t in tokens if t < eot)


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def detect_language(model: Whisper, mel: torch.Tensor) -> Tuple[int, Dict[str, float]]:
    """decoding.py detect_language for one [n_mels, 3000] window."""
    lay = model.layout
    xa = model.encode(mel[None])
    logits = model.decode(torch.tensor([[lay.sot]]), xa)[:, 0]
    mask = torch.ones(logits.shape[-1], dtype=torch.bool)
    mask[list(lay.all_language_tokens)] = False
    logits[:, mask] = -np.inf
    language_token = int(logits.argmax(dim=-1)[0])
    probs = logits.softmax(dim=-1)[0]
    return language_token, {LANGUAGES[i]: float(probs[t]) for i, t in enumerate(lay.all_language_tokens)}


class _Filters:
    def __init__(self, lay: TokenLayout, sample_begin: int, opts: DecodingOptions, n_audio_ctx: int):
        self.lay = lay
        self.sample_begin = sample_begin
        self.opts = opts
        self.suppress = list(lay.suppress_tokens())
        precision = CHUNK_LENGTH / n_audio_ctx
        self.max_initial_timestamp_index = (
            round(opts.max_initial_timestamp / precision) if opts.max_initial_timestamp else None
        )

    def apply(self, logits: torch.Tensor, tokens: torch.Tensor) -> None:
        lay = self.lay
        # SuppressBlank
        if self.opts.suppress_blank and tokens.shape[1] == self.sample_begin:
            logits[:, [lay.blank, lay.eot]] = -np.inf
        # SuppressTokens
        logits[:, self.suppress] = -np.inf
        # ApplyTimestampRules
        if self.opts.without_timestamps:
            return
        tb = lay.timestamp_begin
        logits[:, lay.no_timestamps] = -np.inf
        for k in range(tokens.shape[0]):
            seq = tokens[k, self.sample_begin :].tolist()
            last_was_timestamp = len(seq) >= 1 and seq[-1] >= tb
            penultimate_was_timestamp = len(seq) < 2 or seq[-2] >= tb
            if last_was_timestamp:
                if penultimate_was_timestamp:
                    logits[k, tb:] = -np.inf
                else:
                    logits[k, : lay.eot] = -np.inf
            timestamps = [t for t in seq if t >= tb]
            if timestamps:
                if last_was_timestamp and not penultimate_was_timestamp:
                    timestamp_last = timestamps[-1]
                else:
                    timestamp_last = timestamps[-1] + 1
                logits[k, tb:timestamp_last] = -np.inf
        if tokens.shape[1] == self.sample_begin:
            logits[:, :tb] = -np.inf
            if self.max_initial_timestamp_index is not None:
                last_allowed = tb + self.max_initial_timestamp_index
                logits[:, last_allowed + 1 :] = -np.inf
        logprobs = F.log_softmax(logits.float(), dim=-1)
        for k in range(tokens.shape[0]):
            timestamp_logprob = logprobs[k, tb:].logsumexp(dim=-1)
            max_text_token_logprob = logprobs[k, :tb].max()
            if timestamp_logprob > max_text_token_logprob:
                logits[k, :tb] = -np.inf


class _Greedy:
    def __init__(self, eot: int, temperature: float = 0.0, sample_seed: Optional[int] = None):
        self.eot = eot
        self.temperature = temperature
        self.sample_seed = sample_seed
        self.min_key_margin = float("inf")  # sampling mode: min top1 - top2 of (logits / T + noise)

    def update(self, tokens, logits, sum_logprobs, cache):
        if self.temperature == 0:
            next_tokens = logits.argmax(dim=-1)
        elif self.sample_seed is None:
            next_tokens = torch.distributions.Categorical(logits=logits / self.temperature).sample()
        else:
            pos = tokens.shape[1]
            keys = (logits.double() / self.temperature).numpy().copy()
            for j in range(keys.shape[0]):
                keys[j] += gumbel_noise(self.sample_seed, j, pos, keys.shape[1])
                if int(tokens[j, -1]) != self.eot:
                    top2 = np.partition(keys[j], -2)[-2:]
                    self.min_key_margin = min(self.min_key_margin, float(top2[1] - top2[0]))
            next_tokens = torch.from_numpy(keys.argmax(axis=-1))
        logprobs = F.log_softmax(logits.float(), dim=-1)
        current = logprobs[torch.arange(logprobs.shape[0]), next_tokens]
        sum_logprobs += current * (tokens[:, -1] != self.eot)
        next_tokens[tokens[:, -1] == self.eot] = self.eot
        tokens = torch.cat([tokens, next_tokens[:, None]], dim=-1)
        return tokens, bool((tokens[:, -1] == self.eot).all())

    def finalize(self, tokens, sum_logprobs):
        tokens = F.pad(tokens, (0, 1), value=self.eot)
        return [[t for t in tokens]], [sum_logprobs.tolist()]


class _Beam:
    def __init__(self, beam_size: int, eot: int, patience: Optional[float]):
        self.beam_size = beam_size
        self.eot = eot
        self.patience = patience or 1.0
        self.max_candidates = round(beam_size * self.patience)
        self.finished: Optional[Dict[tuple, float]] = None
        assert self.max_candidates > 0

    def update(self, tokens, logits, sum_logprobs, cache):
        if self.finished is None:
            self.finished = {}
        logprobs = F.log_softmax(logits.float(), dim=-1)
        scores, sources, newly = {}, {}, {}
        for j in range(self.beam_size):
            prefix = tokens[j].tolist()
            for logprob, token in zip(*logprobs[j].topk(self.beam_size + 1)):
                new_logprob = (sum_logprobs[j] + logprob).item()
                sequence = tuple(prefix + [token.item()])
                scores[sequence] = new_logprob
                sources[sequence] = j
        next_tokens, source_indices = [], []
        saved = 0
        for sequence in sorted(scores, key=scores.get, reverse=True):
            if sequence[-1] == self.eot:
                newly[sequence] = scores[sequence]
            else:
                sum_logprobs[len(next_tokens)] = scores[sequence]
                next_tokens.append(sequence)
                source_indices.append(sources[sequence])
                saved += 1
                if saved == self.beam_size:
                    break
        tokens = torch.tensor(next_tokens)
        # rearrange_kv_cache: self-attention K/V only
        for name, t in list(cache.items()):
            if ".attn." in name and t.shape[0] > 1:
                cache[name] = t[source_indices]
        for seq in sorted(newly, key=newly.get, reverse=True):
            if len(self.finished) >= self.max_candidates:
                break
            self.finished[seq] = newly[seq]
        return tokens, len(self.finished) >= self.max_candidates

    def finalize(self, preceding_tokens, sum_logprobs):
        sequences = self.finished if self.finished is not None else {}
        if len(sequences) < self.beam_size:
            for j in list(np.argsort(sum_logprobs.numpy()))[::-1]:
                sequence = preceding_tokens[j].tolist() + [self.eot]
                sequences[tuple(sequence)] = sum_logprobs[j].item()
                if len(sequences) >= self.beam_size:
                    break
        return [[torch.tensor(s) for s in sequences.keys()]], [list(sequences.values())]


def decode_window(model: Whisper, mel_segment: torch.Tensor, opts: DecodingOptions,
                  audio_features: Optional[torch.Tensor] = None) -> DecodingResult:
    """DecodingTask.run for ONE [n_mels, 3000] window (the server never batches)."""
    dims, lay = model.dims, model.layout
    n_group = opts.beam_size or opts.best_of or 1
    n_ctx = dims.n_text_ctx
    sample_len = opts.sample_len or n_ctx // 2
    language = opts.language or "en"
    sot_sequence = lay.sot_sequence(language, opts.task)
    if opts.without_timestamps:
        sot_sequence = sot_sequence + (lay.no_timestamps,)
    initial = list(sot_sequence)
    if opts.prompt:
        initial = [lay.sot_prev] + list(opts.prompt)[-(n_ctx // 2 - 1) :] + initial
    sample_begin = len(initial)
    sot_index = initial.index(lay.sot)
    if opts.beam_size is not None:
        decoder = _Beam(opts.beam_size, lay.eot, opts.patience)
    else:
        decoder = _Greedy(lay.eot, opts.temperature, opts.sample_seed)
    filters = _Filters(lay, sample_begin, opts, dims.n_audio_ctx)

    if audio_features is None:
        audio_features = model.encode(mel_segment[None].float())
    xa = audio_features.repeat_interleave(n_group, dim=0)
    tokens = torch.tensor([initial]).repeat_interleave(n_group, dim=0)
    sum_logprobs = torch.zeros(n_group)
    no_speech_prob = float("nan")
    cache: dict = {}
    min_margin = float("inf")
    for i in range(sample_len):
        inp = tokens if i == 0 else tokens[:, -1:]
        logits = model.decode(inp, xa, cache)
        if i == 0:
            probs_at_sot = logits[:, sot_index].float().softmax(dim=-1)
            no_speech_prob = float(probs_at_sot[0, lay.no_speech])
        logits = logits[:, -1]
        filters.apply(logits, tokens)
        if opts.temperature == 0 or opts.beam_size is not None:
            top2 = logits.topk(2, dim=-1).values
            m = float((top2[:, 0] - top2[:, 1]).min())
            if np.isfinite(m):
                min_margin = min(min_margin, m)
        tokens, completed = decoder.update(tokens, logits, sum_logprobs, cache)
        if completed or tokens.shape[-1] > n_ctx:
            break
    cand_tokens, cand_lp = decoder.finalize(tokens, sum_logprobs)
    cand = []
    for t in cand_tokens[0]:
        t = t[sample_begin:]
        eot_pos = (t == lay.eot).nonzero()
        cand.append(t[: int(eot_pos[0, 0])].tolist())
    # MaximumLikelihoodRanker
    scores = []
    for lp, t in zip(cand_lp[0], cand):
        length = len(t)
        penalty = length if opts.length_penalty is None else ((5 + length) / 6) ** opts.length_penalty
        scores.append(lp / penalty)
    sel = int(np.argmax(scores))
    toks = cand[sel]
    text = render_text(toks, lay.eot).strip()
    return DecodingResult(
        language=language,
        tokens=toks,
        text=text,
        avg_logprob=cand_lp[0][sel] / (len(toks) + 1),
        no_speech_prob=no_speech_prob,
        temperature=opts.temperature,
        compression_ratio=compression_ratio(text),
        sum_logprob=cand_lp[0][sel],
        min_margin=min_margin if (opts.temperature == 0 or opts.beam_size is not None) else decoder.min_key_margin,
    )


# ---- transcribe (upstream transcribe.py seek loop, Appendix A.5) ----------------------
def transcribe(
    model: Whisper,
    audio: np.ndarray,
    *,
    temperature=0.0,
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt_tokens: Optional[Sequence[int]] = None,
    sample_seed: Optional[int] = None,
    **decode_options,
) -> dict:
    """`temperature`: a scalar or upstream's fallback ladder (a tuple, tried in order by `decode_with_fallback`).
    Returns upstream's dict plus `windows` diagnostics (the accepted DecodingResult of every window)."""
    dims, lay = model.dims, model.layout
    decode_options.pop("fp16", None)
    decode_options.pop("word_timestamps", None)
    mel = log_mel_spectrogram(audio, dims.n_mels, padding=N_SAMPLES)
    content_frames = mel.shape[-1] - N_FRAMES
    language_probs = None
    if decode_options.get("language") is None:
        if not model.is_multilingual:
            decode_options["language"] = "en"
        else:
            _, language_probs = detect_language(model, pad_or_trim(mel, N_FRAMES))
            decode_options["language"] = max(language_probs, key=language_probs.get)
    language = decode_options["language"]

    def decode_with_fallback(segment: torch.Tensor, seek: int, prompt: Sequence[int]) -> DecodingResult:
        temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)
        decode_result = None
        audio_features = None
        for attempt, t in enumerate(temperatures):
            kwargs = dict(decode_options)
            if t > 0:  # disable beam_size and patience when t > 0
                kwargs.pop("beam_size", None)
                kwargs.pop("patience", None)
            else:  # disable best_of when t == 0
                kwargs.pop("best_of", None)
            seed = None if sample_seed is None else window_seed(sample_seed, seek, attempt)
            opts = DecodingOptions(**kwargs, temperature=t, prompt=prompt, sample_seed=seed)
            if audio_features is None and len(temperatures) > 1:
                audio_features = model.encode(segment[None].float())  # upstream re-encodes every rung: same numbers
            decode_result = decode_window(model, segment, opts, audio_features)
            needs_fallback = False
            # The repetitiveness check needs real text.  This restatement has no tokenizer assets and renders `<id>`
            # placeholders (render_text), so -- like the product's placeholder mode -- it leaves the check out.
            if HAS_TEXT and compression_ratio_threshold is not None and decode_result.compression_ratio > compression_ratio_threshold:
                needs_fallback = True  # too repetitive
            if logprob_threshold is not None and decode_result.avg_logprob < logprob_threshold:
                needs_fallback = True  # average log probability is too low
            if (no_speech_threshold is not None and decode_result.no_speech_prob > no_speech_threshold
                    and logprob_threshold is not None and decode_result.avg_logprob < logprob_threshold):
                needs_fallback = False  # silence
            if not needs_fallback:
                break
        return decode_result

    seek = 0
    input_stride = N_FRAMES // dims.n_audio_ctx
    time_precision = input_stride * HOP_LENGTH / SAMPLE_RATE
    all_tokens: List[int] = []
    all_segments: List[dict] = []
    windows: List[DecodingResult] = []
    prompt_reset_since = 0
    if initial_prompt_tokens:
        all_tokens.extend(initial_prompt_tokens)
    n_initial = len(all_tokens)

    while seek < content_frames:
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        segment_size = min(N_FRAMES, content_frames - seek)
        mel_segment = pad_or_trim(mel[:, seek : seek + segment_size], N_FRAMES)
        segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
        result = decode_with_fallback(mel_segment, seek, all_tokens[prompt_reset_since:])
        windows.append(result)
        tokens = result.tokens

        if no_speech_threshold is not None:
            should_skip = result.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                should_skip = False
            if should_skip:
                seek += segment_size
                continue

        def new_segment(start, end, toks):
            return {
                "seek": seek, "start": start, "end": end,
                "text": render_text(toks, lay.eot), "tokens": list(toks),
                "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob,
            }

        current_segments: List[dict] = []
        ts = [t >= lay.timestamp_begin for t in tokens]
        single_timestamp_ending = ts[-2:] == [False, True]
        consecutive = [i + 1 for i in range(len(ts) - 1) if ts[i] and ts[i + 1]]
        if consecutive:
            slices = list(consecutive)
            if single_timestamp_ending:
                slices.append(len(tokens))
            last_slice = 0
            for current_slice in slices:
                sliced = tokens[last_slice:current_slice]
                start_pos = sliced[0] - lay.timestamp_begin
                end_pos = sliced[-1] - lay.timestamp_begin
                current_segments.append(
                    new_segment(time_offset + start_pos * time_precision, time_offset + end_pos * time_precision, sliced)
                )
                last_slice = current_slice
            if single_timestamp_ending:
                seek += segment_size
            else:
                last_timestamp_pos = tokens[last_slice - 1] - lay.timestamp_begin
                seek += last_timestamp_pos * input_stride
        else:
            duration = segment_duration
            timestamps = [t for t in tokens if t >= lay.timestamp_begin]
            if timestamps and timestamps[-1] != lay.timestamp_begin:
                duration = (timestamps[-1] - lay.timestamp_begin) * time_precision
            current_segments.append(new_segment(time_offset, time_offset + duration, tokens))
            seek += segment_size

        for segment in current_segments:
            if segment["start"] == segment["end"] or segment["text"].strip() == "":
                segment["text"] = ""
                segment["tokens"] = []
        all_segments.extend({"id": i, **s} for i, s in enumerate(current_segments, start=len(all_segments)))
        all_tokens.extend(t for s in current_segments for t in s["tokens"])
        if not condition_on_previous_text or result.temperature > 0.5:
            prompt_reset_since = len(all_tokens)

    return {
        "text": render_text(all_tokens[n_initial:], lay.eot),
        "segments": all_segments,
        "language": language,
        "language_probs": language_probs,
        "windows": windows,
    }


# ---- in-tree wrapper (torch_whisper.py:49-110) + worker pre-step (worker.py:119-123) ---
SUPPORTED_OPTIONS = {
    "temperature", "compression_ratio_threshold", "logprob_threshold", "no_speech_threshold",
    "condition_on_previous_text", "initial_prompt", "word_timestamps", "prepend_punctuations",
    "append_punctuations", "language", "task", "beam_size", "best_of", "patience", "length_penalty",
    "fp16", "prompt",
}


def normalize_options(options: dict) -> dict:
    """torch_whisper.py:78-110 (`without_timestamps` is converted and vanishes)."""
    opts = dict(options)
    if "log_prob_threshold" in opts and "logprob_threshold" not in opts:
        opts["logprob_threshold"] = opts.pop("log_prob_threshold")
    if "without_timestamps" in opts and "word_timestamps" not in opts:
        opts["word_timestamps"] = not bool(opts.pop("without_timestamps"))
    return {k: v for k, v in opts.items() if k in SUPPORTED_OPTIONS}


def backend_transcribe(model: Whisper, audio: np.ndarray, options: dict, sample_seed: Optional[int] = None):
    """TorchWhisperBackend.transcribe -> ([(start, end, text)], (language, -1.0), raw result)."""
    opts = normalize_options(options)
    for k in ("prepend_punctuations", "append_punctuations", "initial_prompt", "prompt"):
        opts.pop(k, None)  # need tokenizer assets / not reachable through the server's profiles
    result = transcribe(model, audio, sample_seed=sample_seed, **opts)
    segments = [(float(s["start"]), float(s["end"]), str(s["text"])) for s in result["segments"]]
    return segments, (result["language"] or "", -1.0), result


def pcm16_to_float32(pcm_bytes: bytes) -> np.ndarray:
    """stt_server/utils/audio.py:6-8."""
    return np.frombuffer(pcm_bytes, dtype=np.int16).astype(np.float32) / 32768.0
