"""Test driver (not product): starts the REAL reference server through `b200_whisper.launcher` with the engine below the
backend replaced by the host-logic fake (this container has no GPU).  argv is passed to stt_server.main unchanged."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, "/root/reference"]  # ours first: the reference has a `tests` package too

import b200_whisper.backend as bk  # noqa: E402
from b200_whisper.vocab import vocab_for  # noqa: E402
from tests.test_host_logic import FakeEngine, res  # noqa: E402

v = vocab_for(51865)
tb = v.timestamp_begin
ENGINE = FakeEngine(51865, [res([tb, 11, 12, tb + 100])])
bk.get_engine = lambda *a, **k: ENGINE

from b200_whisper.launcher import main  # noqa: E402

main()
