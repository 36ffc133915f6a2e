"""The drop-in claim against the REAL caller: the reference's own `ModelWorker` (stt_server/model/worker.py) resolves
`backend="b200_whisper"` through its registry, constructs `B200WhisperBackend` the way it constructs its own backends and
drives it through `decode_sync` -- bytes in, `DecodeResult` out.  Runs wherever the reference tree is mounted (the build
container); the GPU box has no /root/reference, so the engine below the backend is the host-logic fake (no CUDA here)."""
import os
import sys
import time

import numpy as np
import pytest

import b200_whisper.backend as bk
from b200_whisper.synth import synth_audio
from b200_whisper.vocab import vocab_for
from tests._util import ACCURATE, REALTIME
from tests.test_host_logic import FakeEngine, res

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "stt_server")), reason="reference tree not mounted")


@pytest.fixture
def reference_worker(monkeypatch):
    monkeypatch.syspath_prepend(REF)
    for name in [n for n in sys.modules if n == "stt_server" or n.startswith("stt_server.")]:
        monkeypatch.delitem(sys.modules, name)
    from b200_whisper import register

    register.install()  # wraps the real get_backend in stt_server.model.backends AND stt_server.model.worker
    import stt_server.model.worker as worker

    v = vocab_for(51865)
    tb = v.timestamp_begin
    eng = FakeEngine(51865, [res([tb, 11, 12, tb + 150, tb + 150, 13, tb + 260])])
    audios = []
    real_open = eng.open_call

    def open_call(audio, sample_rate=None):
        audios.append(np.array(audio, dtype=np.float32, copy=True))
        return real_open(audio, sample_rate)

    eng.open_call = open_call
    monkeypatch.setattr(bk, "get_engine", lambda *a, **k: eng)
    yield worker, eng, audios
    for name in [n for n in sys.modules if n == "stt_server" or n.startswith("stt_server.")]:
        sys.modules.pop(name, None)


def _pcm(seed, seconds, rate=16000):
    a = synth_audio(seed, seconds)
    if rate != 16000:
        a = a[:: 16000 // rate] if 16000 % rate == 0 else a
    return (np.clip(a, -1, 1) * 32767).astype(np.int16).tobytes()


def test_real_model_worker_drives_the_backend(reference_worker):
    worker, eng, audios = reference_worker
    assert worker.get_backend("b200_whisper") is bk.B200WhisperBackend
    assert worker.get_backend("faster_whisper").__name__ == "FasterWhisperBackend"  # the reference's own entries survive
    with pytest.raises(RuntimeError, match="openai-whisper"):                       # ... including their own failures
        worker.get_backend("torch_whisper")
    with pytest.raises(ValueError):
        worker.get_backend("no_such_backend")
    w = worker.ModelWorker("random:test-tiny", "cuda:0", "bfloat16", "en", False, base_options={"task": "transcribe"},
                           backend="b200_whisper")
    assert isinstance(w.backend, bk.B200WhisperBackend) and w.backend_name == "b200_whisper"
    pcm = _pcm(3, 4.0)
    opts = dict(REALTIME)
    r = w.decode_sync(pcm, 16000, opts, time.perf_counter())
    assert isinstance(r, worker.DecodeResult) and opts == REALTIME
    assert [(s.start, s.end, s.text) for s in r.segments] == [(0.0, 3.0, "<11><12>"), (3.0, 5.2, "<13>")]
    assert r.language_code == "en" and r.language_probability == -1.0 and abs(r.audio_duration - 4.0) < 1e-6
    assert r.latency_sec > 0 and r.rtf > 0 and r.queue_wait_sec >= 0
    # what reached the backend is exactly pcm16_to_float32(bytes) (utils/audio.py:6-8); language / task came from the worker
    np.testing.assert_array_equal(audios[-1], np.frombuffer(pcm, np.int16).astype(np.float32) / 32768.0)
    d = eng.all_decodes[-1]
    v = vocab_for(51865)
    assert d["initial"] == v.sot_sequence("en", "transcribe") and d["beam"] == 1
    # accurate profile -> beam 5; a per-request language overrides the worker's
    r = w.decode_sync(pcm, 16000, dict(ACCURATE, language="ko"), time.perf_counter())
    assert eng.all_decodes[-1]["beam"] == 5 and eng.all_decodes[-1]["initial"][1] == v.language_token("ko") and r.language_code == "ko"
    # 8 kHz stream: the worker resamples with torchaudio and hands a float array of twice the length
    r = w.decode_sync(_pcm(3, 2.0, 8000), 8000, dict(REALTIME), time.perf_counter())
    assert abs(len(audios[-1]) - 32000) <= 2 and abs(r.audio_duration - 2.0) < 1e-3
    # empty PCM never reaches the backend (worker.py:108-117)
    n = len(eng.all_decodes)
    r = w.decode_sync(b"", 16000, dict(REALTIME), time.perf_counter())
    assert r.segments == [] and r.rtf == -1.0 and len(eng.all_decodes) == n
    # backend errors surface as the exception types the scheduler maps to ERR2002 (decode_scheduler.py:632-644)
    with pytest.raises((RuntimeError, ValueError, TypeError, OSError, TimeoutError)):
        w.decode_sync(pcm, 16000, dict(REALTIME, beam_size=99), time.perf_counter())
    w.close()


def test_real_worker_pool_shares_one_engine(reference_worker, monkeypatch):
    """model_registry.py:230-247 builds pool_size workers with identical arguments: every handle must share one engine."""
    worker, eng, _ = reference_worker
    made = []
    monkeypatch.setattr(bk, "get_engine", lambda *a, **k: (made.append(a), eng)[1])
    pool = [worker.ModelWorker("random:test-tiny", "cuda:0", "bfloat16", None, False, backend="b200_whisper") for _ in range(4)]
    assert len({id(p.backend.engine) for p in pool}) == 1 and len({a for a in made}) == 1
    with pytest.raises(ValueError):  # constructor failures are of a type load_model aborts on cleanly (model_registry.py:281-289)
        worker.ModelWorker("random:test-tiny", "cpu", "bfloat16", None, False, backend="b200_whisper")
    for p in pool:
        p.close()
