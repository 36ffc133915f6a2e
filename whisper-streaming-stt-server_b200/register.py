"""Registration of `b200_whisper` behind the reference's backend registry.

The reference resolves `--model-backend` through a hard-coded switch with no plugin hook
(stt_server/model/backends/__init__.py:9-33) and `ModelWorker` imports that function by name
(stt_server/model/worker.py:8), so `install()` wraps `get_backend` in BOTH namespaces.  It also
tolerates `faster_whisper` being absent (the registry imports it eagerly at
stt_server/model/backends/__init__.py:6).  INTEGRATION.md has the 3-line upstream patch that makes
this wrapper unnecessary.
"""
from __future__ import annotations

import importlib
import sys
import types

BACKEND_NAMES = {"b200_whisper", "b200-whisper", "b200", "blackwell"}


def _stub_faster_whisper_if_missing() -> None:
    try:
        importlib.import_module("faster_whisper")
        return
    except ImportError:
        pass
    fw = types.ModuleType("faster_whisper")
    tr = types.ModuleType("faster_whisper.transcribe")

    def _missing(*_a, **_k):
        raise RuntimeError("faster_whisper is not installed in this environment")

    fw.WhisperModel = _missing
    tr.BatchedInferencePipeline = _missing
    fw.transcribe = tr
    sys.modules["faster_whisper"] = fw
    sys.modules["faster_whisper.transcribe"] = tr


def install() -> None:
    """Make `get_backend("b200_whisper")` return `B200WhisperBackend` inside the reference server."""
    _stub_faster_whisper_if_missing()
    backends = importlib.import_module("stt_server.model.backends")
    original = backends.get_backend
    if getattr(original, "_b200_wrapped", False):
        return

    def get_backend(name: str):
        if (name or "").lower() in BACKEND_NAMES:
            from .backend import B200WhisperBackend

            return B200WhisperBackend
        return original(name)

    get_backend._b200_wrapped = True  # type: ignore[attr-defined]
    backends.get_backend = get_backend
    worker = importlib.import_module("stt_server.model.worker")
    worker.get_backend = get_backend
