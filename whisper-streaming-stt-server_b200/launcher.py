"""`python -m b200_whisper.launcher --model-backend b200_whisper --device cuda:0 ...`

Installs the backend registration, then hands over to the reference server's own entry point
(stt_server/main.py:645) with the unmodified command line.  When the server's generated protobuf modules are absent and
`grpc_tools` is not there to make them (gen/stt/python/v1/__init__.py:3-9), they are built at run time from the tree's
own `proto/stt.proto` (`protostubs.py`).
"""
from __future__ import annotations

import importlib.util
import os


def ensure_proto_stubs() -> None:
    """Make `gen.stt.python.v1.stt_pb2[_grpc]` importable: the generated files if present, run-time stubs otherwise."""
    spec = importlib.util.find_spec("stt_server")
    if spec is None or not spec.submodule_search_locations:
        raise RuntimeError("stt_server is not importable: put the whisper-streaming-stt-server checkout on PYTHONPATH")
    root = os.path.dirname(list(spec.submodule_search_locations)[0])
    if os.path.isfile(os.path.join(root, "gen", "stt", "python", "v1", "stt_pb2.py")):
        return
    from .protostubs import install

    # a checkout keeps proto/ next to the packages; an installed server (pip --target) does not ship it: $B200_WHISPER_PROTO,
    # or the `_tree/` directory tools/install_reference.sh puts next to the packages
    candidates = [os.environ.get("B200_WHISPER_PROTO"), os.path.join(root, "proto", "stt.proto"), os.path.join(root, "_tree", "proto", "stt.proto")]
    for path in candidates:
        if path and os.path.isfile(path):
            install(path)
            return
    raise RuntimeError(f"stt.proto not found (looked at {[c for c in candidates if c]}); set B200_WHISPER_PROTO")


def main() -> None:
    from .register import install

    ensure_proto_stubs()
    install()
    from stt_server.main import main as server_main  # type: ignore

    server_main()


if __name__ == "__main__":
    main()
