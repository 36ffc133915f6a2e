"""Engine / call lifetimes: the reference has no backend close hook (worker.py:160-169), so the engine is
refcounted -- pool handles and OPEN CALLS hold references, `bw_engine_destroy` drops one, the last one tears the engine down.
A destroy while calls are open (atexit while a worker thread is still inside transcribe) must leave those calls usable."""
import ctypes as C
import threading

import pytest

from tests._util import REALTIME, model_spec

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from b200_whisper import _lib as L  # noqa: E402
from b200_whisper.backend import B200WhisperBackend, _ENGINES, _ENGINES_LOCK  # noqa: E402
from b200_whisper.synth import synth_audio  # noqa: E402


def test_destroy_with_open_calls_defers_the_teardown():
    spec = model_spec("test-tiny", seed=41)
    b = B200WhisperBackend(spec, "cuda:0", "float32", max_segments=4, max_sequences=8)
    eng = b.engine
    v = b.vocab
    initial = v.sot_sequence("en", "transcribe")
    audio = synth_audio(950, 3.0)
    want = eng.open_call(audio)
    ref = want.decode(0, initial, 0, None, None, None, sample_len=12)
    want.close()
    calls = [eng.open_call(audio) for _ in range(3)]
    # drop the engine's own reference (what the atexit hook / Engine.close does) while three calls are open
    with _ENGINES_LOCK:
        _ENGINES.pop((spec, 0, "fp32"), None)
    handle = eng.handle
    eng.handle = C.c_void_p()  # the Python wrapper forgets the engine: nobody calls destroy twice
    assert eng.lib.bw_engine_destroy(handle) == 0
    out = [None] * 3

    def work(i):
        out[i] = calls[i].decode(0, initial, 0, None, None, None, sample_len=12)

    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(o["tokens"] == ref["tokens"] for o in out), "open calls must keep working after bw_engine_destroy"
    for c in calls:  # the last close releases the last reference: scheduler joined, memory freed, no crash
        c.close()
    # a fresh engine for the same model can be built afterwards
    b2 = B200WhisperBackend(spec, "cuda:0", "float32", max_segments=4, max_sequences=8)
    assert b2.engine is not eng
    segs, _ = b2.transcribe(audio, dict(REALTIME, language="en"))
    assert isinstance(segs, list)


def test_failed_window_does_not_poison_the_batch():
    """a window the C ABI rejects (token id out of range) inside a bw_decode_many batch fails alone"""
    b = B200WhisperBackend(model_spec("test-tiny", seed=42), "cuda:0", "float32", max_segments=4, max_sequences=8)
    eng = b.engine
    initial = b.vocab.sot_sequence("en", "transcribe")
    audio = synth_audio(951, 2.0)
    with eng.open_call(audio) as c1, eng.open_call(audio) as c2:
        good = dict(initial=initial, sot_index=0, beam_size=None, patience=None, length_penalty=None, sample_len=8)
        bad = dict(good, initial=[10 ** 6])
        with pytest.raises(L.B200WhisperError):
            eng.decode_many([(c1, 0, good), (c2, 0, bad)])  # argument errors are raised before anything is queued
        outs = eng.decode_many([(c1, 0, good), (c2, 0, good)])
        assert all(isinstance(o, dict) and len(o["tokens"]) == 8 for o in outs) and outs[0]["tokens"] == outs[1]["tokens"]
