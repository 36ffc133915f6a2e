"""Test driver (not product): starts the REAL reference server through `b200_whisper.launcher` with the engine below the
backend replaced by the host-logic fake (this container has no GPU); B200_TEST_REAL_ENGINE=1 keeps the real engine (GPU box).
argv is passed to stt_server.main unchanged.  The reference is taken from PYTHONPATH / $STT_SERVER_ROOT / /root/reference.

B200_TEST_ENERGY_VAD=1 also provides a stand-in for the `silero_vad` package (absent here, weights not vendored): a frame is
speech when its RMS exceeds 0.01.  It only exists so that the server's own VAD gate, endpointing and partial-decode
schedule run (--vad-threshold > 0); it says nothing about Silero's decisions."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]  # ours first: the reference has a `tests` package too
sys.path.insert(1, os.environ.get("STT_SERVER_ROOT") or "/root/reference")

import b200_whisper.backend as bk  # noqa: E402
from b200_whisper.vocab import vocab_for  # noqa: E402

if os.environ.get("B200_TEST_REAL_ENGINE") != "1":
    from tests.test_host_logic import FakeEngine, res  # noqa: E402

    v = vocab_for(51865)
    tb = v.timestamp_begin
    ENGINE = FakeEngine(51865, [res([tb, 11, 12, tb + 100])])
    bk.get_engine = lambda *a, **k: ENGINE

if os.environ.get("B200_TEST_ENERGY_VAD") == "1":
    import types

    import torch

    class EnergyVAD:
        def __call__(self, audio_tensor, sample_rate):
            rms = float(torch.sqrt(torch.mean(audio_tensor.float() ** 2)))
            return torch.tensor(1.0 if rms > 0.01 else 0.0)

        def reset_states(self):
            pass

    stub = types.ModuleType("silero_vad")
    stub.load_silero_vad = lambda onnx=False: EnergyVAD()
    sys.modules["silero_vad"] = stub

from b200_whisper.launcher import main  # noqa: E402

main()
