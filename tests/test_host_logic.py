"""Host-side logic of the backend without a GPU: option normalisation and result mapping (the template is the
reference's tests/test_mlx_whisper_backend.py), token tables, device parsing, registration, and the seek loop /
segment assembly driven by a fake engine and compared with the oracle's restatement of upstream transcribe()."""
import sys
import types

import numpy as np
import pytest

import b200_whisper.backend as bk
from b200_whisper.backend import B200WhisperBackend
from b200_whisper.melfilters import mel_filterbank
from b200_whisper.synth import MODEL_DIMS, synth_audio
from b200_whisper.vocab import Detokenizer, normalize_language, vocab_for
from oracle import whisper_oracle as wo
from oracle.tables import layout_for_vocab


class FakeCall:
    def __init__(self, engine, audio):
        self.engine = engine
        self.content_frames = (len(audio) + 480000) // 160 - 3000
        self.decodes = []

    def decode(self, seek, initial, sot_index, beam_size, patience, length_penalty, **kw):
        self.decodes.append({"seek": seek, "initial": list(initial), "sot_index": sot_index, "beam": beam_size, **kw})
        self.engine.all_decodes.append(self.decodes[-1])
        return dict(self.engine.script[min(len(self.engine.all_decodes) - 1, len(self.engine.script) - 1)])

    def detect_language(self, seek=0):
        v = self.engine.vocab
        probs = np.full(v.num_languages, 0.001, np.float32)
        probs[5] = 0.9
        return v.first_language_token + 5, probs

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


class FakeEngine:
    def __init__(self, n_vocab=51865, script=None):
        self.dims = MODEL_DIMS["test-tiny"] if n_vocab == 51865 else MODEL_DIMS["test-tiny.en"]
        self.vocab = vocab_for(n_vocab)
        self.script = script or []
        self.all_decodes = []

    def open_call(self, audio, sample_rate=None):
        self.last_sample_rate = sample_rate
        return FakeCall(self, audio)

    def decode_many(self, items):
        self.batches = getattr(self, "batches", []) + [len(items)]
        return [call.decode(seek, **kw) for call, seek, kw in items]


@pytest.fixture
def fake_backend(monkeypatch):
    def make(script, n_vocab=51865):
        eng = FakeEngine(n_vocab, script)
        monkeypatch.setattr(bk, "get_engine", lambda *a, **k: eng)
        return B200WhisperBackend("random:test-tiny", "cuda:0", "bfloat16"), eng

    return make


def res(tokens, avg=-0.5, nsp=0.01):
    return {"tokens": tokens, "sum_logprob": avg * (len(tokens) + 1), "avg_logprob": avg, "no_speech_prob": nsp, "n_steps": len(tokens)}


def test_tables_match_oracle_layout():
    for n in (51864, 51865, 51866):
        v, o = vocab_for(n), layout_for_vocab(n)
        assert v.suppress_tokens() == list(o.suppress_tokens())
        for f in ("eot", "sot", "translate", "transcribe", "sot_lm", "sot_prev", "no_speech", "no_timestamps", "timestamp_begin",
                  "num_languages", "multilingual"):
            assert getattr(v, f) == getattr(o, f), (n, f)
        assert v.sot_sequence("ko", "translate") == list(o.sot_sequence("ko", "translate"))
    assert vocab_for(51866).language_token("yue") == 50358
    with pytest.raises(ValueError):
        vocab_for(51865).language_token("yue")
    assert normalize_language("Korean") == "ko" and normalize_language("EN") == "en"
    with pytest.raises(ValueError):
        normalize_language("klingon")


def test_mel_filterbank_matches_oracle():
    for n in (80, 128):
        assert np.abs(mel_filterbank(n) - wo.mel_filters(n)).max() < 1e-7


def test_option_normalisation_mirrors_torch_whisper(fake_backend):
    b, eng = fake_backend([res([])])
    opts = {"log_prob_threshold": -0.7, "without_timestamps": True, "vad_filter": True, "hotwords": "x", "beam_size": 5,
            "best_of": 5, "patience": 1.0, "temperature": 0.0}
    frozen = dict(opts)
    n = b._normalize_options(opts)
    assert opts == frozen
    assert n == {"logprob_threshold": -0.7, "word_timestamps": False, "beam_size": 5, "best_of": 5, "patience": 1.0, "temperature": 0.0}
    assert n == {k: v for k, v in wo.normalize_options(opts).items()}


def test_result_mapping_and_language(fake_backend):
    v = vocab_for(51865)
    tb = v.timestamp_begin
    b, eng = fake_backend([res([tb, 100, 200, tb + 50, tb + 50, 300, tb + 120])])
    segs, info = b.transcribe(synth_audio(1, 3.0), {"language": "en", "beam_size": 1})
    assert [(s.start, s.end) for s in segs] == [(0.0, 1.0), (1.0, 2.4)]
    assert segs[0].text == "<100><200>" and info.language == "en" and info.language_probability == -1.0
    # language unset -> detection; torch_whisper parity keeps the reported probability at -1.0
    segs, info = b.transcribe(synth_audio(1, 3.0), {"beam_size": 1})
    assert info.language == "ko" and info.language_probability == -1.0 and abs(b.last_language_probability - 0.9) < 1e-6
    assert eng.all_decodes[-1]["initial"][:3] == [v.sot, v.language_token("ko"), v.transcribe]
    # .en vocabulary: language forced to en, sot sequence is a single token
    b2, eng2 = fake_backend([res([])], n_vocab=51864)
    segs, info = b2.transcribe(synth_audio(1, 1.0), {})
    assert info.language == "en" and eng2.all_decodes[-1]["initial"] == [vocab_for(51864).sot]
    assert eng2.all_decodes[-1]["beam"] is None  # no beam_size -> greedy
    # empty audio short-circuit and beam range
    assert b.transcribe(np.zeros(0, np.float32), {"language": "en"})[0] == []
    with pytest.raises(ValueError):
        b.transcribe(synth_audio(1, 1.0), {"language": "en", "beam_size": 9})


def _oracle_with_script(script, audio, n_vocab=51865, **opts):
    """Run the oracle's restatement of upstream transcribe() with decode_window replaced by the same script."""
    dims = MODEL_DIMS["test-tiny"]
    model = types.SimpleNamespace(dims=wo.ModelDimensions(**dims.__dict__), layout=layout_for_vocab(n_vocab), is_multilingual=True,
                                  encode=lambda mel: None)
    calls = []
    _oracle_with_script.options = opt_log = []

    def fake_decode_window(m, mel_segment, o, audio_features=None):
        r = script[min(len(calls), len(script) - 1)]
        calls.append(list(o.prompt or []))
        opt_log.append(o)
        return wo.DecodingResult(language="en", tokens=list(r["tokens"]), text=wo.render_text(r["tokens"], model.layout.eot).strip(),
                                 avg_logprob=r["avg_logprob"], no_speech_prob=r["no_speech_prob"], temperature=o.temperature,
                                 compression_ratio=wo.compression_ratio(wo.render_text(r["tokens"], model.layout.eot).strip()),
                                 sum_logprob=r["sum_logprob"])

    orig = wo.decode_window
    wo.decode_window = fake_decode_window
    try:
        out = wo.transcribe(model, audio, **opts)
    finally:
        wo.decode_window = orig
    return out, calls


@pytest.mark.parametrize("case", ["pairs_then_open", "single_ts_ending", "no_timestamps", "no_speech_skip", "empty"])
def test_seek_loop_matches_oracle(fake_backend, case):
    v = vocab_for(51865)
    tb = v.timestamp_begin
    audio = synth_audio(2, 70.0)  # three 30 s windows worth of content
    scripts = {
        "pairs_then_open": [res([tb, 11, 12, tb + 400, tb + 400, 13, tb + 900, tb + 900, 14]), res([tb + 10, 21, tb + 700, tb + 700]),
                            res([tb, 31, 32, tb + 1500])],
        "single_ts_ending": [res([tb, 11, tb + 200, tb + 200, 12, tb + 1400]), res([tb, 21, tb + 1500])],
        "no_timestamps": [res([11, 12, 13]), res([21, tb + 600]), res([31])],
        "no_speech_skip": [res([tb, 11, tb + 100], avg=-1.5, nsp=0.9), res([tb, 21, tb + 800]), res([tb, 31, tb + 300])],
        "empty": [res([])],
    }
    script = scripts[case]
    b, eng = fake_backend(script)
    for cond in (True, False):
        eng.all_decodes.clear()
        opts = {"language": "en", "beam_size": 1, "condition_on_previous_text": cond}
        got = b.transcribe_raw(audio, **b._normalize_options(opts))
        want, prompts = _oracle_with_script(script, audio, language="en", beam_size=1, condition_on_previous_text=cond)
        assert [(s["seek"], s["start"], s["end"], s["tokens"]) for s in got["segments"]] == \
               [(s["seek"], s["start"], s["end"], s["tokens"]) for s in want["segments"]]
        assert got["text"] == want["text"]
        # prompt carry-over: [sot_prev] + previous tokens precede the sot sequence, sot_index follows
        assert len(eng.all_decodes) == len(prompts)
        for dec, prompt in zip(eng.all_decodes, prompts):
            exp = ([v.sot_prev] + prompt[-223:] if prompt else []) + v.sot_sequence("en", None)
            assert dec["initial"] == exp and dec["initial"][dec["sot_index"]] == v.sot


@pytest.mark.parametrize("case", ["accept_first", "logprob_fallback", "repetitive_to_the_top", "silence_is_not_retried",
                                  "scalar_temperature", "multi_window_prompt_reset"])
def test_temperature_fallback_ladder_matches_oracle(fake_backend, monkeypatch, case):
    """decode_with_fallback of upstream transcribe.py: which rungs run, with which decoder options, which result is kept,
    and the prompt reset after a high-temperature window -- backend vs the oracle's restatement on the same script."""
    v = vocab_for(51865)
    tb = v.timestamp_begin
    ladder = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)
    good = res([tb, 11, 12, tb + 300], avg=-0.4)
    low = res([tb, 13, 14, tb + 200], avg=-1.5)
    rep = res([tb] + [11] * 80 + [tb + 400], avg=-0.3)  # compression ratio of "<11><11>..." is far above 2.4
    assert wo.compression_ratio(wo.render_text(rep["tokens"], v.eot)) > 2.4
    cases = {
        "accept_first": (10.0, ladder, [good], [0.0], 0.0),
        "logprob_fallback": (10.0, ladder, [low, low, good], [0.0, 0.2, 0.4], 0.4),
        "repetitive_to_the_top": (10.0, ladder, [rep], list(ladder), 1.0),
        "silence_is_not_retried": (10.0, ladder, [res([tb, 11, tb + 100], avg=-1.5, nsp=0.9)], [0.0], None),
        "scalar_temperature": (10.0, 0.4, [low], [0.4], 0.4),
        # window 1 falls back to T = 0.6 (> 0.5: its tokens are not fed to window 2 as a prompt); window 2 accepted at T = 0
        "multi_window_prompt_reset": (50.0, (0.0, 0.6), [low, res([tb, 21, tb + 1500], avg=-0.2), good], [0.0, 0.6, 0.0], None),
    }
    seconds, temperature, script, want_rungs, want_temperature = cases[case]
    audio = synth_audio(3, seconds)
    b, eng = fake_backend(script)
    # the repetitiveness check is off in placeholder-text mode (product and oracle alike); switch it on for this host-logic test
    b.check_compression_ratio = True
    monkeypatch.setattr(wo, "HAS_TEXT", True)
    opts = {"language": "en", "beam_size": 5, "best_of": 3, "patience": 1.0, "temperature": temperature}
    got = b.transcribe_raw(audio, _seed=77, **b._normalize_options(opts))
    want, prompts = _oracle_with_script(script, audio, language="en", beam_size=5, best_of=3, patience=1.0,
                                        temperature=temperature, sample_seed=77)
    oracle_opts = _oracle_with_script.options
    assert [d.get("temperature", 0.0) for d in eng.all_decodes] == want_rungs == [o.temperature for o in oracle_opts]
    for dec, o in zip(eng.all_decodes, oracle_opts):
        if o.temperature > 0:  # sampling rung: beam search and patience off, best_of hypotheses, per-attempt seed
            assert dec["beam"] is None and o.beam_size is None and o.patience is None
            assert dec["best_of"] == 3 == o.best_of and dec["seed"] == o.sample_seed
        else:                  # T = 0 rung: beam search, best_of dropped
            assert dec["beam"] == 5 == o.beam_size and o.best_of is None and "best_of" not in dec
    assert len({o.sample_seed for o in oracle_opts}) == len(oracle_opts)  # every attempt draws from its own stream
    key = lambda r: [(s["seek"], s["start"], s["end"], s["tokens"], s["temperature"], round(s["compression_ratio"], 6)) for s in r["segments"]]
    assert key(got) == key(want)
    if want_temperature is not None:
        assert [s["temperature"] for s in got["segments"]] == [want_temperature] * len(got["segments"]) and got["segments"]
    if case == "silence_is_not_retried":
        assert got["segments"] == []  # no_speech_prob > threshold and avg_logprob below: window skipped, no retry
    if case == "multi_window_prompt_reset":
        assert eng.all_decodes[-1]["initial"][0] == v.sot and prompts[-1] == []
    with pytest.raises(ValueError):
        b.transcribe_raw(audio, temperature=-0.1, language="en")
    with pytest.raises(ValueError):
        b.transcribe_raw(audio, temperature=0.2, best_of=9, language="en")


def test_sampling_seeds(fake_backend, monkeypatch):
    """Every call of a backend draws from its own stream; B200_WHISPER_SEED makes a process reproducible; `_seed` pins a call."""
    v = vocab_for(51865)
    script = [res([v.timestamp_begin, 11, v.timestamp_begin + 100])]
    audio = synth_audio(4, 3.0)
    monkeypatch.setenv("B200_WHISPER_SEED", "42")
    b1, eng1 = fake_backend(script)
    b2, eng2 = fake_backend(script)
    for b in (b1, b2):
        for _ in range(2):
            b.transcribe(audio, {"language": "en", "temperature": 0.5, "best_of": 2})
    s1 = [d["seed"] for d in eng1.all_decodes]
    s2 = [d["seed"] for d in eng2.all_decodes]
    assert s1 == s2 and s1[0] != s1[1] and all(0 <= s < 2 ** 64 for s in s1)
    monkeypatch.delenv("B200_WHISPER_SEED")
    b3, eng3 = fake_backend(script)
    b3.transcribe(audio, {"language": "en", "temperature": 0.5})
    assert eng3.all_decodes[0]["seed"] != s1[0] and eng3.all_decodes[0]["best_of"] is None
    b3.transcribe_raw(audio, language="en", temperature=0.5, _seed=7)
    b3.transcribe_raw(audio, language="en", temperature=0.5, _seed=7)
    assert eng3.all_decodes[1]["seed"] == eng3.all_decodes[2]["seed"] == bk.window_seed(7, 0, 0) == wo.window_seed(7, 0, 0)
    # temperature 0 never asks the engine to sample
    b3.transcribe(audio, {"language": "en", "temperature": 0.0, "best_of": 5})
    assert "temperature" not in eng3.all_decodes[-1]


def test_wrapper_matches_the_real_torch_whisper_wrapper(fake_backend):
    """Option normalisation and result mapping against outputs of the REAL `TorchWhisperBackend` methods
    (stt_server/model/backends/torch_whisper.py:49-110, run in the build container by make_golden_server_options.py)."""
    import json
    import os

    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "server_options.json")))
    b, _ = fake_backend([res([])])
    norm = fx["torch_whisper_normalize_options"]
    for given, want in zip(norm["inputs"], norm["outputs"]):
        frozen = json.dumps(given, sort_keys=True)
        assert b._normalize_options(given) == want, given
        assert json.dumps(given, sort_keys=True) == frozen
    mapping = fx["torch_whisper_result_mapping"]
    for given, want in zip(mapping["inputs"], mapping["outputs"]):
        segs, info = b._to_segments(given)
        assert [[s.start, s.end, s.text] for s in segs] == want["segments"]
        assert (info.language, info.language_probability) == (want["language"], want["language_probability"])


def test_server_option_surface_from_the_reference(fake_backend, caplog):
    """The option dicts that really reach `transcribe` -- the server's shipped decode profiles and every key a client may send
    (fixture generated from the real reference by tests/golden/make_golden_server_options.py): the parity tests' REALTIME /
    ACCURATE constants ARE those profiles, and each allowed key is either honoured or warned about and dropped, never an error
    (torch_whisper.py:78-110 precedent), with the caller's dict left untouched."""
    import json
    import logging
    import os

    from tests._util import ACCURATE, REALTIME

    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "server_options.json")))
    assert fx["decode_profiles"]["realtime"] == REALTIME and fx["decode_profiles"]["accurate"] == ACCURATE
    assert fx["default_decode_profile"] == REALTIME and fx["default_task"] == "transcribe"
    v = vocab_for(51865)
    b, eng = fake_backend([res([v.timestamp_begin, 11, v.timestamp_begin + 100])])
    audio = synth_audio(5, 2.0)
    samples = {"append_punctuations": ".", "prepend_punctuations": "(", "chunk_length": 30, "clip_timestamps": "0", "hotwords": "x",
               "initial_prompt": "hello", "language": "en", "max_initial_timestamp": 1.0, "no_repeat_ngram_size": 2, "prefix": "x",
               "prompt_reset_on_temperature": 0.5, "repetition_penalty": 1.1, "suppress_blank": True, "suppress_tokens": [-1],
               "task": "transcribe", "temperature_increment_on_fallback": 0.2, "vad_filter": True, "vad_parameters": {},
               "word_timestamps": False, "condition_on_previous_text": False}
    honoured = bk.SUPPORTED_OPTIONS | {"log_prob_threshold", "without_timestamps"}
    for key in fx["allowed_decode_option_keys"]:
        for profile in ("realtime", "accurate"):
            opts = dict(fx["decode_profiles"][profile], task=fx["default_task"])
            if key not in opts:
                opts[key] = samples[key]
            frozen = json.dumps(opts, sort_keys=True, default=str)
            caplog.clear()
            with caplog.at_level(logging.WARNING, logger="stt_server.model_backend"):
                segs, info = b.transcribe(audio, opts)
            assert json.dumps(opts, sort_keys=True, default=str) == frozen
            assert isinstance(segs, list) and info.language_probability == -1.0
            dropped = any("Dropping unsupported" in r.getMessage() and f" {key}=" in r.getMessage() for r in caplog.records)
            assert dropped == (key not in honoured), key
    # the profiles decode with the decoder upstream would pick: beam search with 1 / 5 beams, best_of dropped at temperature 0
    assert {d["beam"] for d in eng.all_decodes} == {1, 5} and all("best_of" not in d for d in eng.all_decodes)


def test_device_parsing_and_compute_types(monkeypatch):
    assert bk.parse_device("cuda", 3) == 0 and bk.parse_device("cuda:5", 0) == 5
    monkeypatch.setattr(bk, "get_engine", lambda *a, **k: FakeEngine())
    for ct, want in (("float32", "fp32"), ("fp32", "fp32"), ("bfloat16", "bf16"), ("float16", "bf16"), ("int8", "bf16"), ("weird", "bf16")):
        assert B200WhisperBackend("random:test-tiny", "cuda:0", ct).compute == want
    for dev in ("cpu", "mps", "mlx"):
        with pytest.raises(ValueError):
            B200WhisperBackend("random:test-tiny", dev, "bfloat16")


def test_checkpoint_resolution(tmp_path, monkeypatch):
    import torch

    from b200_whisper.synth import random_state_dict

    dims = MODEL_DIMS["test-tiny"]
    path = tmp_path / "tiny-test.pt"
    torch.save({"dims": dims.__dict__, "model_state_dict": random_state_dict(dims, 1)}, path)
    d, sd, name = bk.load_checkpoint(str(path))
    assert d == dims and "decoder.token_embedding.weight" in sd
    monkeypatch.setenv("B200_WHISPER_MODEL_DIR", str(tmp_path))
    d2, _, _ = bk.load_checkpoint("tiny-test")
    assert d2 == dims
    with pytest.raises(RuntimeError):
        bk.load_checkpoint("definitely-not-a-model")
    d3, sd3, _ = bk.load_checkpoint("random:test-v3:3")
    assert d3.n_mels == 128 and dict(sd3.items())["decoder.token_embedding.weight"].shape == (51866, 128)
    # the lazy view yields exactly the eager state dict (same generator order)
    eager = random_state_dict(MODEL_DIMS["test-v3"], 3, emb_std=0.1)
    assert all(torch.equal(t, eager[k]) for k, t in sd3.items())


def test_registration_wraps_reference_get_backend(monkeypatch):
    """register.install() against a stand-in for stt_server.model.{backends,worker} (the real package needs
    grpc stubs that are not generated here)."""
    pkg = types.ModuleType("stt_server"); model = types.ModuleType("stt_server.model")
    backends = types.ModuleType("stt_server.model.backends"); worker = types.ModuleType("stt_server.model.worker")

    def get_backend(name):
        if name in ("torch_whisper", "faster_whisper"):
            return "REF:" + name
        raise ValueError(f"Unknown model backend: {name}")

    backends.get_backend = get_backend
    worker.get_backend = get_backend
    for n, m in (("stt_server", pkg), ("stt_server.model", model), ("stt_server.model.backends", backends), ("stt_server.model.worker", worker)):
        monkeypatch.setitem(sys.modules, n, m)
    from b200_whisper import register

    register.install()
    register.install()  # idempotent
    for alias in ("b200_whisper", "B200", "blackwell", "b200-whisper"):
        assert backends.get_backend(alias) is B200WhisperBackend and worker.get_backend(alias) is B200WhisperBackend
    assert backends.get_backend("torch_whisper") == "REF:torch_whisper"
    with pytest.raises(ValueError):
        backends.get_backend("nope")
    assert "faster_whisper" in sys.modules  # stubbed when absent so the registry's eager import works


def _write_synthetic_rank_file(path, n_ranks):
    """A tiktoken rank file of the real SIZE (the special-token ids depend on it) with made-up merges: 256 bytes, a chain that
    turns " hello" / " world" into single tokens, then filler sequences of non-ASCII bytes."""
    import base64

    toks = [bytes([b]) for b in range(256)]
    toks += [b"he", b"ll", b"hell", b"hello", b" hello", b"wo", b"rl", b"worl", b"world", b" world"]
    a = 128
    while len(toks) < n_ranks:
        for b in range(128, 256):
            for c in range(128, 256):
                if len(toks) < n_ranks:
                    toks.append(bytes([a, b, c]))
        a += 1
    with open(path, "w") as fh:
        for rank, tok in enumerate(toks):
            fh.write(f"{base64.b64encode(tok).decode()} {rank}\n")
    return {tok: rank for rank, tok in enumerate(toks)}


def test_text_rendering_and_initial_prompt_with_a_rank_file(fake_backend, tmp_path, monkeypatch):
    """With B200_WHISPER_VOCAB_DIR the backend renders real text and honours `initial_prompt` (upstream transcribe.py: the
    prompt is encoded with a leading space and fed as [sot_prev] + tokens in front of the sot sequence).  The real asset ships
    with openai-whisper; a synthetic rank file of the real size stands in for it here."""
    ranks = _write_synthetic_rank_file(tmp_path / "multilingual.tiktoken", 50257)
    monkeypatch.setenv("B200_WHISPER_VOCAB_DIR", str(tmp_path))
    v = vocab_for(51865)
    d = Detokenizer(v)
    assert d.has_text and d.encoding.n_vocab == 51865
    hello, world = ranks[b" hello"], ranks[b" world"]
    assert d.encode(" hello world") == [hello, world] and d.decode([hello, world, v.timestamp_begin + 7]) == " hello world"
    # special tokens sit where the engine's token tables expect them
    for name, tok in (("<|endoftext|>", v.eot), ("<|startoftranscript|>", v.sot), ("<|en|>", v.language_token("en")),
                      ("<|transcribe|>", v.transcribe), ("<|startofprev|>", v.sot_prev), ("<|nospeech|>", v.no_speech),
                      ("<|notimestamps|>", v.no_timestamps), ("<|0.00|>", v.timestamp_begin), ("<|30.00|>", v.timestamp_begin + 1500)):
        assert d.encoding.encode(name, allowed_special="all") == [tok], name
    tb = v.timestamp_begin
    b, eng = fake_backend([res([tb, hello, world, tb + 100])])
    segs, info = b.transcribe(synth_audio(6, 3.0), {"language": "en", "beam_size": 1, "initial_prompt": "hello"})
    assert [(s.start, s.end, s.text) for s in segs] == [(0.0, 2.0, " hello world")]
    assert eng.all_decodes[-1]["initial"] == [v.sot_prev, hello] + v.sot_sequence("en", None)
    assert eng.all_decodes[-1]["initial"][eng.all_decodes[-1]["sot_index"]] == v.sot
    raw = b.transcribe_raw(synth_audio(6, 3.0), language="en", initial_prompt="hello")
    assert raw["text"] == " hello world" and raw["segments"][0]["compression_ratio"] == bk.compression_ratio("hello world")


def test_real_checkpoint_without_rank_file_is_refused(monkeypatch, tmp_path):
    """a real checkpoint served with `<id>` placeholder text is a silent failure -- the backend must refuse to load
    unless the model is random-init or the operator opts in; and the rank file is found under B200_WHISPER_VOCAB_DIR or in an
    installed openai-whisper's assets."""
    from b200_whisper.vocab import find_rank_file

    monkeypatch.delenv("B200_WHISPER_VOCAB_DIR", raising=False)
    monkeypatch.delenv("B200_WHISPER_ALLOW_PLACEHOLDER_TEXT", raising=False)
    assert find_rank_file("multilingual") is None
    with pytest.raises(RuntimeError, match="tiktoken"):
        Detokenizer(vocab_for(51865), allow_placeholders=False)
    eng = FakeEngine()
    monkeypatch.setattr(bk, "get_engine", lambda *a, **k: eng)
    with pytest.raises(RuntimeError, match="placeholder"):
        B200WhisperBackend("/models/large-v3.pt", "cuda:0", "bfloat16")
    monkeypatch.setenv("B200_WHISPER_ALLOW_PLACEHOLDER_TEXT", "1")
    assert not B200WhisperBackend("/models/large-v3.pt", "cuda:0", "bfloat16").detok.has_text
    # auto-location: a package called `whisper` with assets/<name>.tiktoken on the import path
    pkg = tmp_path / "whisper"
    (pkg / "assets").mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    (pkg / "assets" / "gpt2.tiktoken").write_text("")
    monkeypatch.syspath_prepend(str(tmp_path))
    import importlib

    importlib.invalidate_caches()
    assert find_rank_file("gpt2") == str(pkg / "assets" / "gpt2.tiktoken")


def test_placeholder_mode_skips_the_repetitiveness_check(fake_backend):
    v = vocab_for(51865)
    tb = v.timestamp_begin
    rep = res([tb] + [11] * 80 + [tb + 400], avg=-0.3)
    b, eng = fake_backend([rep])
    assert not b.detok.has_text and not b.check_compression_ratio
    got = b.transcribe_raw(synth_audio(3, 10.0), language="en", beam_size=5, temperature=(0.0, 0.4, 0.8))
    assert len(eng.all_decodes) == 1 and got["segments"][0]["temperature"] == 0.0


def test_compute_type_remap_is_logged(fake_backend, caplog):
    import logging

    bk._REMAP_WARNED.clear()
    with caplog.at_level(logging.WARNING, logger="stt_server.model_backend"):
        b, _ = fake_backend([res([])])
        assert B200WhisperBackend("random:test-tiny", "cuda:0", "int8").compute == "bf16"
        assert B200WhisperBackend("random:test-tiny", "cuda:0", "float16").compute == "bf16"
        assert B200WhisperBackend("random:test-tiny", "cuda:0", "float32").compute == "fp32"
    msgs = [r.getMessage() for r in caplog.records]
    assert any("compute_type=int8 runs as bfloat16" in m for m in msgs) and any("compute_type=float16" in m for m in msgs)


def test_detokenizer_placeholder():
    d = Detokenizer(vocab_for(51865))
    v = vocab_for(51865)
    assert d.decode([5, v.timestamp_begin + 3, 7]) == "<5><7>" and d.encode("hi") is None


def test_side_doors_route_through_the_same_seek_loop(fake_backend):
    """transcribe_pcm16 hands the raw bytes + rate to the engine (no host-side conversion); transcribe_many keeps the
    input order, accepts one options dict or one per call, and re-raises the first failure after all calls finished"""
    v = vocab_for(51865)
    tb = v.timestamp_begin
    res = {"tokens": [tb, 100, 200, tb + 50], "sum_logprob": -1.0, "avg_logprob": -0.2, "no_speech_prob": 0.1, "n_steps": 4,
           "t_queue": 0.0, "t_encode": 0.0, "t_decode": 0.0}
    b, eng = fake_backend([res])
    pcm = (np.arange(48000) % 200 - 100).astype(np.int16).tobytes()
    segs, info = b.transcribe_pcm16(pcm, 48000, {"language": "en", "beam_size": 1})
    assert eng.last_sample_rate == 48000 and len(segs) == 1 and info.language == "en"
    assert b.transcribe_pcm16(b"", 8000, {"language": "en"})[0] == []
    with pytest.raises(ValueError):
        b.transcribe_pcm16(pcm, -1, {})
    audios = [np.zeros(16000 * (i + 1), np.float32) for i in range(3)]
    out = b.transcribe_many(audios + [pcm], {"language": "en"}, [None, None, None, 48000])
    assert len(out) == 4 and all(len(s) == 1 for s, _ in out)
    assert eng.batches == [4], "one bw_decode_many for the whole batch (one window each), not a thread per item"
    out2 = b.transcribe_many(audios, [{"language": "en"}, {"language": "de"}, {"language": "en"}])
    assert [i.language for _, i in out2] == ["en", "de", "en"]
    # items with several windows stay in lockstep: round k carries the k-th window of every item that still has one
    eng.batches = []
    long_audios = [np.zeros(16000 * 70, np.float32), np.zeros(16000 * 5, np.float32), np.zeros(16000 * 40, np.float32)]
    b.transcribe_many(long_audios, {"language": "en"})
    assert eng.batches == [3, 2, 1]
    with pytest.raises(ValueError):
        b.transcribe_many(audios, {"language": "en", "beam_size": 99})  # every call fails on the option check
    with pytest.raises(ValueError):
        b.transcribe_many(audios, [{}], None)
