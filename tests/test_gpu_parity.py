"""Stage-level and end-to-end parity of the CUDA path against the CPU oracle (through the C ABI).

Tolerances (BASELINE.json north_star): log-mel max-abs <= 1e-3; encoder rel-L2 <= 1e-2 in bf16;
greedy / beam token ids bit-exact in the fp32 validation mode (the oracle's minimum top-1/top-2 logit
margin on the path is asserted to be far above fp32 reduction-order noise so the claim is meaningful).
"""
import threading

import numpy as np
import pytest

from tests._util import ACCURATE, REALTIME, model_spec, oracle_model, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

torch = pytest.importorskip("torch")

from b200_whisper.backend import B200WhisperBackend, get_engine  # noqa: E402
from b200_whisper.synth import synth_audio  # noqa: E402
from oracle import whisper_oracle as wo  # noqa: E402


def backend(name="test-tiny", compute="float32", **kw):
    return B200WhisperBackend(model_spec(name, **kw), "cuda:0", compute, max_segments=16, max_sequences=48)


@pytest.mark.parametrize("name,n_mels", [("test-tiny", 80), ("test-v3", 128)])
@pytest.mark.parametrize("seconds,padding", [(30.0, 480000), (3.96, 480000), (1.0, 480000), (0.3, 480000), (2.5, 0), (10.01, 137),
                                             (15.0, 480000), (20.3, 480000)])  # 4, 8 and 16 frames per CTA
def test_log_mel(name, n_mels, seconds, padding):
    eng = backend(name).engine
    audio = synth_audio(int(seconds * 10), seconds)
    got = eng.mel(audio, padding)
    ref = wo.log_mel_spectrogram(audio, n_mels, padding=padding).numpy()
    assert got.shape == ref.shape
    err = float(np.abs(got - ref).max())
    assert err <= 1e-3, f"log-mel max-abs error {err}"


def test_log_mel_silence_and_full_scale():
    eng = backend().engine
    for audio in (np.zeros(16000, np.float32), np.full(8000, 0.999, np.float32),
                  np.sign(np.sin(np.arange(48000) * 0.3)).astype(np.float32) * 0.9):
        got = eng.mel(audio, 480000)
        ref = wo.log_mel_spectrogram(audio, 80, padding=480000).numpy()
        assert float(np.abs(got - ref).max()) <= 1e-3


@pytest.mark.parametrize("compute,tol", [("float32", 2e-4), ("bfloat16", 1e-2)])
@pytest.mark.parametrize("name", ["test-tiny", "test-v3"])
def test_encoder(name, compute, tol):
    b = backend(name, compute)
    model = oracle_model(name)
    mels = []
    for s in range(2):
        audio = synth_audio(20 + s, 6.0 + 3 * s)
        mels.append(wo.pad_or_trim(wo.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000), 3000).numpy())
    mel = np.stack(mels)
    got = b.engine.encode(mel)
    ref = model.encode(torch.from_numpy(mel)).numpy()
    r = rel_l2(got, ref)
    assert r <= tol, f"{name} {compute}: encoder rel-L2 {r}"


def test_decoder_logits_fp32():
    b = backend("test-tiny", "float32")
    model = oracle_model("test-tiny")
    audio = synth_audio(3, 5.0)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(audio, 80, padding=480000), 3000)
    lay = model.layout
    tokens = [lay.sot, lay.language_token("en"), lay.transcribe, lay.timestamp_begin, 1000, 2000, 3000, 400, 50]
    got = b.engine.decode_logits(mel.numpy(), tokens)
    xa = model.encode(mel[None])
    ref = model.decode(torch.tensor([tokens]), xa)[0].numpy()
    err = float(np.abs(got - ref).max())
    assert err <= 2e-3, f"decoder logits max-abs error {err} (scale {np.abs(ref).max()})"


@pytest.mark.parametrize("name", ["test-tiny", "test-v3"])
def test_decoder_logits_bf16(name):
    """bf16 product path (tcgen05 GEMMs, mma.sync cross-attention, split-K clusters) against the fp32 oracle"""
    b = backend(name, "bfloat16")
    model = oracle_model(name)
    audio = synth_audio(3, 5.0)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000), 3000)
    lay = model.layout
    tokens = [lay.sot, lay.language_token("en"), lay.transcribe, lay.timestamp_begin, 1000, 2000, 3000, 400, 50, 7, 11]
    got = b.engine.decode_logits(mel.numpy(), tokens)
    xa = model.encode(mel[None])
    ref = model.decode(torch.tensor([tokens]), xa)[0].numpy()
    r = rel_l2(got, ref)
    assert r <= 3e-2, f"{name}: bf16 decoder logits rel-L2 {r}"
    agree = float((got.argmax(-1) == ref.argmax(-1)).mean())
    print(f"{name}: bf16 logits rel-L2 {r:.4f}, argmax agreement {agree:.2f}")


def _oracle_segments(name, audio, opts, **kw):
    segs, info, raw = wo.backend_transcribe(oracle_model(name, **kw), audio, opts)
    return segs, info, raw


def _check_transcribe(b, name, audio, opts, **kw):
    ref_segs, ref_info, raw = _oracle_segments(name, audio, opts, **kw)
    margins = [w.min_margin for w in raw["windows"]]
    assert min(margins) > 2e-4, f"oracle margin {min(margins)} too small for an exactness claim; pick another seed"
    segs, info = b.transcribe(audio, opts)
    got = [(round(s.start, 3), round(s.end, 3), s.text) for s in segs]
    want = [(round(a, 3), round(e, 3), t) for a, e, t in ref_segs]
    assert got == want
    assert info.language == ref_info[0] and info.language_probability == -1.0
    return raw


@pytest.mark.parametrize("name", ["test-tiny", "test-tiny.en", "test-v3"])
@pytest.mark.parametrize("profile", ["realtime", "accurate"])
def test_transcribe_token_exact_fp32(name, profile):
    opts = dict(REALTIME if profile == "realtime" else ACCURATE, language="en", task="transcribe")
    b = backend(name, "float32")
    for seed, seconds in ((1, 4.0), (2, 11.5)):
        _check_transcribe(b, name, synth_audio(seed, seconds), opts)


def test_transcribe_language_detection_and_greedy_fp32():
    b = backend("test-tiny", "float32")
    audio = synth_audio(5, 6.0)
    raw = _check_transcribe(b, "test-tiny", audio, dict(REALTIME, task="transcribe"))  # language unset -> detect
    assert raw["language_probs"] is not None
    # no beam_size -> GreedyDecoder path
    _check_transcribe(b, "test-tiny", audio, {"temperature": 0.0, "language": "en"})


def test_transcribe_eot_and_multiwindow_fp32():
    """EOT-biased weights end hypotheses early; 41 s of audio exercises the seek loop + prompt carry-over."""
    kw = dict(eot_bias=6.0)
    b = backend("test-tiny", "float32", **kw)
    opts = dict(REALTIME, language="en")
    raw = _check_transcribe(b, "test-tiny", synth_audio(9, 41.0), opts, **kw)
    assert len(raw["windows"]) >= 2
    _check_transcribe(b, "test-tiny", synth_audio(10, 7.0), dict(ACCURATE, language="en"), **kw)


@pytest.mark.parametrize("beam,patience,length_penalty", [(8, 1.0, 1.0), (5, 2.0, None), (3, 1.6, 0.6), (2, 0.5, 1.0),
                                                          (5, 0.5, 1.0), (3, 0.5, None), (5, 0.3, 1.0)])  # x.5 products: round-half-even
def test_beam_patience_and_length_penalty_fp32(beam, patience, length_penalty):
    """BeamSearchDecoder corners outside the two server profiles: the widest beam, patience != 1 (finished pool of
    round(beam * patience) candidates, larger or smaller than the beam -- Python's round(): 5 * 0.5 -> 2, 3 * 0.5 -> 2,
    5 * 0.3 -> 2 (1.5 -> 2) -- ), length normalisation instead of the GNMT penalty.  EOT-biased weights so that hypotheses do finish and the pool / ranker paths are taken."""
    opts = dict(ACCURATE, language="en", beam_size=beam, best_of=beam, patience=patience)
    if length_penalty is None:
        opts.pop("length_penalty")
    else:
        opts["length_penalty"] = length_penalty
    for seed, seconds, eot_bias in ((71, 5.0, 4.0), (72, 9.0, 5.0), (73, 3.0, 4.5)):
        kw = dict(eot_bias=eot_bias)
        _check_transcribe(backend("test-tiny", "float32", **kw), "test-tiny", synth_audio(seed, seconds), opts, **kw)


def test_transcribe_bf16_segments_follow_the_oracle():
    """bf16 product mode through the public transcribe(): the segments' token stream agrees with the fp32 oracle's on (nearly)
    all leading tokens for this model (test-tiny: measured 100 %; the per-token statistics for every model size are
    tests/test_gpu_bf16_decode.py::test_bf16_first_divergence_statistics, profiles/r2_bf16_divergence.json)."""
    b = backend("test-tiny", "bfloat16")
    opts = dict(REALTIME, language="en")
    total = agree = 0
    for seed in range(4):
        audio = synth_audio(30 + seed, 5.0)
        _, _, raw = _oracle_segments("test-tiny", audio, opts)
        res = b.transcribe_raw(audio, **b._normalize_options(opts))
        got = [t for s in res["segments"] for t in s["tokens"]]
        want = [t for s in raw["segments"] for t in s["tokens"]]
        n = min(len(got), len(want))
        first = next((i for i in range(n) if got[i] != want[i]), n)
        total += max(len(want), 1)
        agree += first
    print(f"bf16 first-divergence: {agree}/{total} leading tokens agree with the fp32 oracle")
    assert total > 0 and agree >= 0.75 * total, (agree, total)


def _raw_key(result):
    return [(s["seek"], round(s["start"], 3), round(s["end"], 3), list(s["tokens"]), s["temperature"]) for s in result["segments"]]


@pytest.mark.parametrize("name", ["test-tiny", "test-v3"])
def test_temperature_sampling_token_exact_fp32(name):
    """GreedyDecoder at temperature > 0 with best_of hypotheses: the device draws Categorical(logits / T) by Gumbel-max
    from the counter-based generator the oracle restates (`gumbel_noise`), so the sampled token streams, the ranking
    among the best_of hypotheses and the log-probabilities can be checked exactly.  Cases whose smallest top-1/top-2
    margin of the perturbed logits is within fp32 noise are skipped, and enough must remain."""
    checked = 0
    for kw, extra in ((dict(eot_bias=4.0), {}), (dict(), dict(sample_len=20))):
        b = backend(name, "float32", **kw)
        for seed, temp, best_of in ((11, 0.7, 3), (12, 1.0, 5), (13, 0.2, 1), (14, 0.5, 8), (15, 0.9, None)):
            opts = dict(REALTIME, language="en", temperature=temp, best_of=best_of)
            if best_of is None:
                opts.pop("best_of")
            audio = synth_audio(40 + seed, 6.0)
            model = oracle_model(name, **kw)
            want = wo.transcribe(model, audio, sample_seed=seed, **{k: v for k, v in wo.normalize_options(opts).items()
                                                                   if k not in ("word_timestamps",)}, **extra)
            if min(w.min_margin for w in want["windows"]) < 5e-4:
                continue
            got = b.transcribe_raw(audio, _seed=seed, **b._normalize_options(opts), **extra)
            assert _raw_key(got) == _raw_key(want), f"{name} seed {seed} T {temp} best_of {best_of}"
            for sg, sw in zip(got["segments"], want["segments"]):
                assert abs(sg["avg_logprob"] - sw["avg_logprob"]) < 1e-3 and abs(sg["no_speech_prob"] - sw["no_speech_prob"]) < 1e-4
            # same seed -> same draw, whatever else the engine is batching; another seed -> another draw
            again = b.transcribe_raw(audio, _seed=seed, **b._normalize_options(opts), **extra)
            assert _raw_key(again) == _raw_key(got)
            checked += 1
    assert checked >= 6, f"only {checked} sampling cases had a usable margin"


def test_temperature_fallback_ladder_fp32():
    """decode_with_fallback on the device path: thresholds that no rung can meet walk the whole ladder (beam search at
    T = 0, then sampling rungs) and keep the last result; a ladder whose first rung passes equals the scalar T = 0 run."""
    kw = dict(eot_bias=4.0)
    b = backend("test-tiny", "float32", **kw)
    model = oracle_model("test-tiny", **kw)
    audio = synth_audio(51, 8.0)
    base = {k: v for k, v in wo.normalize_options(dict(ACCURATE, language="en", best_of=3)).items() if k != "word_timestamps"}
    checked = 0
    for seed in (21, 22, 23, 24):
        o = dict(base, temperature=(0.0, 0.4, 0.8), logprob_threshold=10.0)
        want = wo.transcribe(model, audio, sample_seed=seed, **o)
        if min(w.min_margin for w in want["windows"]) < 5e-4:
            continue
        before = b.engine.stats()["windows"]
        got = b.transcribe_raw(audio, _seed=seed, **dict(b._normalize_options(dict(ACCURATE, language="en", best_of=3)),
                                                         temperature=(0.0, 0.4, 0.8), logprob_threshold=10.0))
        assert _raw_key(got) == _raw_key(want)
        assert all(s["temperature"] == 0.8 for s in got["segments"]) and got["segments"]
        assert b.engine.stats()["windows"] - before == 3 * len(want["windows"])  # every rung re-decodes the window
        checked += 1
    assert checked >= 2
    o = dict(base, temperature=(0.0, 0.2, 0.4), compression_ratio_threshold=None, logprob_threshold=None)
    got = b.transcribe_raw(audio, **dict(b._normalize_options(dict(ACCURATE, language="en")), temperature=(0.0, 0.2, 0.4),
                                         compression_ratio_threshold=None, logprob_threshold=None))
    want = wo.transcribe(model, audio, **dict(o, temperature=0.0))
    assert _raw_key(got) == _raw_key(want) and all(s["temperature"] == 0.0 for s in got["segments"])


def test_temperature_sampling_distribution():
    """The draws follow Categorical(softmax(filtered logits / T)): 600 seeds at the first sampled position (whose
    support is the 51 allowed initial timestamps) against the oracle's probabilities, chi-square at ~5 sigma; and the
    reported log-probability is the un-tempered log-softmax of the drawn token (GreedyDecoder.update)."""
    b = backend("test-tiny", "float32")
    model = oracle_model("test-tiny")
    lay = model.layout
    audio = synth_audio(61, 5.0)
    temp = 1.3
    mel = wo.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000)
    mel = wo.pad_or_trim(mel[:, : mel.shape[-1] - 3000], 3000)  # the window transcribe() decodes: content, then zeros
    initial = list(lay.sot_sequence("en", "transcribe"))
    logits = model.decode(torch.tensor([initial]), model.encode(mel[None]))[:, -1]
    wo._Filters(lay, len(initial), wo.DecodingOptions(), model.dims.n_audio_ctx).apply(logits, torch.tensor([initial]))
    logprobs = torch.log_softmax(logits[0].double(), -1).numpy()
    p = torch.softmax(logits[0].double() / temp, -1).numpy()
    n = 600
    counts = np.zeros_like(p)
    with b.engine.open_call(audio) as call:
        for seed in range(n):
            r = call.decode(0, initial, 0, None, None, None, sample_len=1, temperature=temp, best_of=1, seed=1000 + seed)
            assert len(r["tokens"]) == 1
            tok = r["tokens"][0]
            assert p[tok] > 0, "drew a token the logit filters forbid"
            assert abs(r["sum_logprob"] - logprobs[tok]) < 1e-3
            counts[tok] += 1
    order = np.argsort(-p)
    exp = n * p[order]
    keep = exp >= 5
    obs_b = np.append(counts[order][keep], counts[order][~keep].sum())
    exp_b = np.append(exp[keep], exp[~keep].sum())
    if exp_b[-1] < 1e-9:
        obs_b, exp_b = obs_b[:-1], exp_b[:-1]
    chi2 = float(((obs_b - exp_b) ** 2 / exp_b).sum())
    dof = len(exp_b) - 1
    print(f"sampling chi-square {chi2:.1f} over {dof} degrees of freedom ({int(keep.sum())} tokens with expectation >= 5)")
    assert dof >= 3 and chi2 < dof + 5 * np.sqrt(2 * dof)
    # C-ABI contract: sampling is a GreedyDecoder mode
    from b200_whisper._lib import B200WhisperError
    with b.engine.open_call(audio) as call:
        with pytest.raises(B200WhisperError):
            call.decode(0, initial, 0, 5, 1.0, None, temperature=0.5, best_of=2)
        with pytest.raises(B200WhisperError):
            call.decode(0, initial, 0, None, None, None, temperature=0.5, best_of=9)


def test_concurrent_sessions_batch_and_match_serial():
    """Many host threads (= pool handles, model_registry.py:564-606) transcribing at once are coalesced by
    the engine and still return exactly their serial results."""
    b = backend("test-tiny", "float32")
    opts = dict(REALTIME, language="en")
    audios = [synth_audio(100 + i, 2.0 + 0.7 * i) for i in range(12)]
    serial = [b.transcribe(a, opts) for a in audios[:4]]
    out = [None] * len(audios)
    handles = [B200WhisperBackend(b.model_size, "cuda:0", "float32") for _ in audios]  # share the engine
    assert all(h.engine is b.engine for h in handles)

    def work(i):
        out[i] = handles[i].transcribe(audios[i], opts if i % 3 else dict(ACCURATE, language="en"))

    steps0 = b.engine.stats()["decode_steps"]
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(audios))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in (1, 2):
        assert out[i] == serial[i]
    stats = b.engine.stats()
    rows_per_step = (stats["rows"]) / max(1, stats["decode_steps"])
    assert stats["decode_steps"] - steps0 < 224 * len(audios), "no cross-session batching happened"
    for i, a in enumerate(audios):
        o = opts if i % 3 else dict(ACCURATE, language="en")
        ref_segs, _, _ = wo.backend_transcribe(oracle_model("test-tiny"), a, o)
        got = [(round(s.start, 3), round(s.end, 3), s.text) for s in out[i][0]]
        assert got == [(round(x, 3), round(y, 3), t) for x, y, t in ref_segs], f"session {i}"


def test_sampling_is_batch_invariant():
    """Sampling windows decoded next to arg-max and beam-search windows of other sessions (one continuous batch, shared
    decoder steps) return what they return alone: the draw depends on (seed, hypothesis, position, token) only."""
    kw = dict(eot_bias=4.0)
    b = backend("test-tiny", "float32", **kw)
    audios = [synth_audio(200 + i, 3.0 + 0.5 * i) for i in range(9)]
    plans = [dict(REALTIME, language="en", temperature=0.6, best_of=4), dict(REALTIME, language="en"),
             dict(ACCURATE, language="en"), dict(REALTIME, language="en", temperature=(0.0, 0.5, 1.0), logprob_threshold=10.0, best_of=2)]

    def run(i):
        return _raw_key(b.transcribe_raw(audios[i], _seed=500 + i, **b._normalize_options(plans[i % len(plans)])))

    serial = [run(i) for i in range(len(audios))]
    out = [None] * len(audios)

    def work(i):
        out[i] = run(i)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(audios))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert out == serial
    assert any(k[4] > 0 for r in serial for k in r), "no sampled segment in the mix"


def test_errors_and_edge_cases():
    b = backend("test-tiny", "float32")
    segs, info = b.transcribe(np.zeros(0, np.float32), dict(REALTIME, language="en"))
    assert segs == [] and info.language == "en"
    segs, info = b.transcribe(np.zeros(1, np.float32), dict(REALTIME, language="en"))  # n >= 1 reaches the backend
    assert isinstance(segs, list)
    opts = dict(REALTIME, language="en", vad_filter=True, hotwords="x")  # unknown keys: warn-and-drop
    frozen = dict(opts)
    b.transcribe(synth_audio(1, 1.0), opts)
    assert opts == frozen, "options dict must not be mutated"
    with pytest.raises(ValueError):
        B200WhisperBackend(b.model_size, "cpu", "float32")
    with pytest.raises(ValueError):
        b.transcribe(synth_audio(1, 1.0), dict(REALTIME, language="xx"))


def test_large_v3_parity():
    """BASELINE.json configs[3]/[4] architecture at full size (32 + 32 layers, d = 1280, 128 mel bins, random init): the
    shapes the bench runs -- persistent pair GEMMs with K = 1280 / 5120, 20-head persistent attention, LayerNorm-fused
    row GEMMs -- against the fp32 CPU oracle: encoder rel-L2 <= 1e-2 in bf16, bf16 logits, token-exact fp32 decode."""
    name = "large-v3"
    model = oracle_model(name)
    audio = synth_audio(77, 7.0)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000), 3000)
    xa = model.encode(mel[None])
    b16 = B200WhisperBackend(model_spec(name), "cuda:0", "bfloat16", max_segments=8, max_sequences=16, max_encoder_batch=4)
    got = b16.engine.encode(mel.numpy())
    r = rel_l2(got, xa.numpy())
    print(f"large-v3 encoder rel-L2 (bf16 vs fp32 oracle): {r:.2e}")
    assert r <= 1e-2, f"encoder rel-L2 {r}"
    # batch of 4 identical windows through the two-stream split path must give the same rows as a batch of one
    got4 = b16.engine.encode(np.stack([mel.numpy()] * 4))
    assert rel_l2(got4[3], got[0]) <= 1e-6 and rel_l2(got4[1], got[0]) <= 1e-6
    lay = model.layout
    toks = list(lay.sot_sequence("en", "transcribe")) + [lay.timestamp_begin, 500, 900, 12000, 7, 11, 13]
    ref = model.decode(torch.tensor([toks]), xa)[0].numpy()
    lg = b16.engine.decode_logits(mel.numpy(), toks)
    r = rel_l2(lg, ref)
    print(f"large-v3 decoder logits rel-L2 (bf16 vs fp32 oracle): {r:.2e}, argmax agreement {float((lg.argmax(-1) == ref.argmax(-1)).mean()):.2f}")
    assert r <= 3e-2, f"bf16 logits rel-L2 {r}"
    del b16
    b32 = B200WhisperBackend(model_spec(name), "cuda:0", "float32", max_segments=4, max_sequences=8, max_encoder_batch=1)
    opts = dict(REALTIME, language="en")
    raw = wo.transcribe(model, audio, sample_len=48, **wo.normalize_options(opts))  # bounded: 48 CPU decoder steps
    segs = b32.transcribe_raw(audio, sample_len=48, **b32._normalize_options(opts))["segments"]
    got_t = [t for s_ in segs for t in s_["tokens"]]
    want_t = [t for s_ in raw["segments"] for t in s_["tokens"]]
    assert min(w.min_margin for w in raw["windows"]) > 2e-4
    assert got_t == want_t, "fp32 validation mode differs from the oracle at large-v3"


@pytest.mark.parametrize("name", ["tiny.en", "base", "small"])
def test_real_model_sizes(name):
    """BASELINE.json configs[0]/[1] architectures at full size (random init): encoder rel-L2 in bf16, token-exact
    transcribe in the fp32 validation mode, bf16 logits close to the oracle."""
    model = oracle_model(name)
    audio = synth_audio(42, 6.0)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000), 3000)
    xa = model.encode(mel[None])
    b16 = backend(name, "bfloat16")
    r = rel_l2(b16.engine.encode(mel.numpy()), xa.numpy())
    assert r <= 1e-2, f"{name}: encoder rel-L2 {r}"
    lay = model.layout
    toks = list(lay.sot_sequence("en", "transcribe")) + [lay.timestamp_begin, 500, 900, 12000]
    ref = model.decode(torch.tensor([toks]), xa)[0].numpy()
    r = rel_l2(b16.engine.decode_logits(mel.numpy(), toks), ref)
    assert r <= 3e-2, f"{name}: bf16 logits rel-L2 {r}"
    b32 = backend(name, "float32")
    _check_transcribe(b32, name, audio, dict(REALTIME, language="en"))
    _check_transcribe(b32, name, audio, dict(ACCURATE, language="en"))
