"""Raw-PCM ingest on the device (bw_resample_pcm16 / bw_call_open_pcm16 through the C ABI) against the golden vectors of
the real reference functions and the CPU oracle; `transcribe_pcm16` and `transcribe_many` against `transcribe`."""
import os

import numpy as np
import pytest

from tests._util import ACCURATE, REALTIME, model_spec

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

from b200_whisper.backend import B200WhisperBackend  # noqa: E402
from b200_whisper.synth import synth_audio  # noqa: E402
from oracle import audio_ingest as ai  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest.npz"))
CASES = sorted((int(k.split("_")[1]), int(k.split("_")[2])) for k in GOLD.files if k.startswith("pcm_"))


@pytest.fixture(scope="module")
def be():
    return B200WhisperBackend(model_spec("test-tiny"), "cuda:0", "float32", max_segments=16, max_sequences=48)


@pytest.mark.parametrize("rate,n", CASES)
def test_device_resampler_matches_reference_golden(be, rate, n):
    pcm = GOLD[f"pcm_{rate}_{n}"]
    y = be.engine.resample_pcm16(pcm.tobytes(), rate)
    ref = GOLD[f"y16k_{rate}_{n}"]
    assert y.shape == ref.shape
    # int16 -> float is exact; the polyphase sums differ from torch's conv1d only in fp32 summation order
    assert float(np.abs(y - ref).max()) <= (0.0 if rate == 16000 else 2e-6)


@pytest.mark.parametrize("rate", [8000, 22050, 44100, 48000])
def test_device_resampler_long_clip_matches_oracle(be, rate):
    n = int(rate * 7.3) + 5
    rng = np.random.default_rng(rate)
    pcm = (np.clip(0.3 * rng.standard_normal(n), -1, 0.9999) * 32768).astype(np.int16)
    y = be.engine.resample_pcm16(pcm, rate)
    ref = ai.ingest(pcm.tobytes(), rate)
    assert y.shape == ref.shape and float(np.abs(y - ref).max()) <= 2e-6


def test_transcribe_pcm16_equals_transcribe_of_reference_ingest(be):
    """the side door gives what the unchanged worker path gives: transcribe(ensure_16k(pcm16_to_float32(b), rate))"""
    opts = dict(REALTIME, language="en", task="transcribe")
    for rate, seconds in ((16000, 5.0), (48000, 4.0), (8000, 6.5), (44100, 3.0)):
        x = synth_audio(70 + rate // 1000, seconds)  # 16 kHz float in [-1, 1)
        # a clip "recorded" at `rate`: linear interpolation is enough, it only has to be some int16 signal at that rate
        t = np.arange(int(seconds * rate)) * (16000.0 / rate)
        pcm = (np.clip(np.interp(t, np.arange(x.size), x), -1, 0.9999) * 32768).astype(np.int16)
        want_segs, want_info = be.transcribe(ai.ingest(pcm.tobytes(), rate), opts)
        got_segs, got_info = be.transcribe_pcm16(pcm.tobytes(), rate, opts)
        assert got_info == want_info
        assert [(s.start, s.end, s.text) for s in got_segs] == [(s.start, s.end, s.text) for s in want_segs], rate
    with pytest.raises(ValueError):
        be.transcribe_pcm16(b"\x00\x00", 0, opts)
    with pytest.raises(ValueError):
        be.transcribe_pcm16(b"\x00\x00\x00", 16000, opts)  # odd byte count
    assert be.transcribe_pcm16(b"", 16000, opts)[0] == []


def test_transcribe_many_matches_serial_calls(be):
    opts_a = dict(REALTIME, language="en", task="transcribe")
    opts_b = dict(ACCURATE, language="en", task="transcribe")
    audios = [synth_audio(90 + i, 2.0 + i) for i in range(5)]
    pcm48 = (np.clip(np.interp(np.arange(48000 * 3) / 3.0, np.arange(16000 * 3), synth_audio(99, 3.0)), -1, 0.9999) * 32768).astype(np.int16)
    items = audios + [pcm48.tobytes()]
    rates = [None] * 5 + [48000]
    options = [opts_a, opts_b, opts_a, opts_b, opts_a, opts_a]
    serial = [be.transcribe(a, o) if r is None else be.transcribe_pcm16(a, r, o) for a, r, o in zip(items, rates, options)]
    before = be.engine.stats()
    many = be.transcribe_many(items, options, rates)
    after = be.engine.stats()
    assert len(many) == len(serial)
    for (gs, gi), (ws, wi) in zip(many, serial):
        assert gi == wi and [(s.start, s.end, s.text) for s in gs] == [(s.start, s.end, s.text) for s in ws]
    # the six calls were batched: fewer encoder launches than windows
    assert after["encoder_batches"] - before["encoder_batches"] < after["windows"] - before["windows"]
    with pytest.raises(ValueError):
        be.transcribe_many(audios, [opts_a])
