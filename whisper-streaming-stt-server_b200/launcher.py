"""`python -m b200_whisper.launcher --model-backend b200_whisper --device cuda:0 ...`

Installs the backend registration, then hands over to the reference server's own entry point
(stt_server/main.py:645) with the unmodified command line.
"""
from __future__ import annotations


def main() -> None:
    from .register import install

    install()
    from stt_server.main import main as server_main  # type: ignore

    server_main()


if __name__ == "__main__":
    main()
