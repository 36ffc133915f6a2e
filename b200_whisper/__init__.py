"""Import alias for the product package.

The product lives in `whisper-streaming-stt-server_b200/` (a directory name Python cannot
import directly); this shim points `b200_whisper` at it so `import b200_whisper.backend`
resolves to `whisper-streaming-stt-server_b200/backend.py`.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_os.path.join(_os.path.dirname(_here), "whisper-streaming-stt-server_b200")]

from .version import __version__  # noqa: E402,F401
