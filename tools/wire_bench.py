"""Wire-level load test: the UNMODIFIED reference server (started through b200_whisper.launcher) under the reference's own
load generator `tools/bench/grpc_load_test.py`, run unchanged.

  python tools/wire_bench.py [--fake-engine] [--channels 64] [--seconds 10] [--pool-size 16] [--model random:large-v3]
                             [--server-root /root/reference] [-- extra grpc_load_test.py arguments]

--energy-vad turns the server's VAD gate on over an energy stand-in for silero_vad (absent here), so that its endpointing and
partial-decode schedule run.  --fake-engine replaces the engine below the backend by the host-logic fake (tests/_ref_server_driver.py): no GPU needed; what
is measured then is the ceiling of the server's Python control plane in front of a zero-cost backend.  Without it the real
engine runs (needs a B200 and the reference checkout on the same box)."""
import argparse
import contextlib
import io
import json
import os
import runpy
import socket
import subprocess
import sys
import tempfile
import time
import wave

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def reference_layout(root=None):
    """Where the unmodified reference lives: a checkout (packages, config/, proto/, tools/ side by side) or the install
    tools/install_reference.sh makes under baseline/_ref (packages at the top, the non-package files under _tree/).
    Returns (package root for PYTHONPATH, directory holding config/ proto/ tools/) or None."""
    for cand in ([root] if root else []) + [os.environ.get("STT_SERVER_ROOT"), "/root/reference", os.path.join(REPO, "baseline", "_ref")]:
        if cand and os.path.isdir(os.path.join(cand, "stt_server")):
            tree = os.path.join(cand, "_tree") if os.path.isdir(os.path.join(cand, "_tree")) else cand
            if os.path.isfile(os.path.join(tree, "proto", "stt.proto")):
                return cand, tree
    return None


def parse_load_test_summary(text: str) -> dict:
    """grpc_load_test.py prints `* Section` headers followed by `    key: value` lines; keep them as {section: {key: value}}"""
    out, section = {}, None
    for line in text.splitlines():
        if line.startswith("* "):
            section = line[2:].strip()
            out[section] = {}
        elif section and line.startswith("    ") and ":" in line:
            k, v = line.strip().split(":", 1)
            v = v.strip()
            try:
                out[section][k] = float(v[:-1]) if v.endswith("s") else float(v)
            except ValueError:
                out[section][k] = v
    return out


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--fake-engine", action="store_true")
    ap.add_argument("--energy-vad", action="store_true", help="energy stand-in for silero_vad, VAD gate on (partial decodes run)")
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--pool-size", type=int, default=16)
    ap.add_argument("--model", default="random:large-v3")
    ap.add_argument("--decode-profile", default="realtime")
    ap.add_argument("--server-root", default=None, help="reference checkout or baseline/_ref install (default: first one found)")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--compute-type", default="bfloat16")
    ap.add_argument("--json", action="store_true", help="last stdout line = one JSON object with the parsed load-test summary")
    ap.add_argument("--server-log", default=None, help="file for the server's stdout/stderr (default: discarded)")
    ap.add_argument("--ready-timeout", type=float, default=600.0, help="seconds to wait for the server's gRPC port")
    args, extra = ap.parse_known_args()
    extra = [a for a in extra if a != "--"]
    layout = reference_layout(args.server_root)
    if layout is None:
        msg = "no reference tree (looked at --server-root, $STT_SERVER_ROOT, /root/reference, baseline/_ref: run tools/install_reference.sh)"
        print(json.dumps({"unavailable": msg}) if args.json else msg)
        return
    pkg_root, tree = layout
    sys.path.append(pkg_root)

    from b200_whisper import protostubs
    from b200_whisper.synth import synth_audio

    tmp = tempfile.mkdtemp(prefix="wire_bench_")
    wav = os.path.join(tmp, "speech.wav")
    pcm = (np.clip(synth_audio(11, args.seconds), -1, 1) * 32767).astype(np.int16)
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    port, mport, wport = free_port(), free_port(), free_port()
    cfg = os.path.join(tmp, "server.yaml")
    with open(cfg, "w") as fh:
        fh.write(open(os.path.join(tree, "config", "server.yaml")).read())
        fh.write(f"\nws_host: 127.0.0.1\nws_port: {wport}\nmax_sessions: {max(256, 2 * args.channels)}\n"
                 "max_sessions_per_ip: 0\nmax_sessions_per_api_key: 0\ncreate_session_rps: 0\nmax_audio_bytes_per_sec: 0\n"
                 "max_audio_bytes_per_sec_burst: 0\n")  # one load generator = one client IP: lift the per-IP limits
    # the energy stand-in for silero_vad lives in the test driver; with the real engine the driver keeps the engine
    use_driver = args.fake_engine or args.energy_vad
    entry = [os.path.join(REPO, "tests", "_ref_server_driver.py")] if use_driver else ["-m", "b200_whisper.launcher"]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REPO, pkg_root, os.environ.get("PYTHONPATH", "")]),
               B200_WHISPER_PROTO=os.path.join(tree, "proto", "stt.proto"), STT_SERVER_ROOT=pkg_root)
    if not args.fake_engine:
        env["B200_TEST_REAL_ENGINE"] = "1"
    if args.energy_vad:
        env["B200_TEST_ENERGY_VAD"] = "1"
    vad_threshold = "0.5" if args.energy_vad else "0"
    log = open(args.server_log, "w") if args.server_log else subprocess.DEVNULL
    server = subprocess.Popen([sys.executable, *entry, "--config", cfg, "--model-backend", "b200_whisper", "--model", args.model,
                               "--device", args.device, "--compute-type", args.compute_type, "--port", str(port),
                               "--metrics-port", str(mport), "--vad-threshold", vad_threshold,
                               "--model-pool-size", str(args.pool_size), "--language", "en", "--log-level", "WARNING"],
                              cwd=REPO, env=env, stdout=log, stderr=subprocess.STDOUT)
    try:
        import grpc

        protostubs.install(os.path.join(tree, "proto", "stt.proto"))
        grpc.channel_ready_future(grpc.insecure_channel(f"127.0.0.1:{port}")).result(timeout=args.ready_timeout)
        sys.argv = ["grpc_load_test.py", "--target", f"127.0.0.1:{port}", "--channels", str(args.channels), "--iterations", "1",
                    "--audio", wav, "--chunk-ms", "100", "--realtime", "--decode-profile", args.decode_profile, "--language", "en",
                    "--vad-mode", "continue", *extra]
        t0 = time.time()
        buf = io.StringIO()
        code = 0
        try:
            with contextlib.redirect_stdout(buf):
                runpy.run_path(os.path.join(tree, "tools", "bench", "grpc_load_test.py"), run_name="__main__")
        except SystemExit as exc:
            code = exc.code
        wall = time.time() - t0
        print(buf.getvalue())
        if code:
            print(f"grpc_load_test.py exited with {code}")
        print(f"wire_bench: {args.channels} channels x {args.seconds:.0f} s of audio in {wall:.1f} s wall "
              f"({args.channels * args.seconds / wall:.1f} audio-s/s through the wire), engine = {'fake' if args.fake_engine else 'B200'}")
        if args.json:
            summary = parse_load_test_summary(buf.getvalue())
            print(json.dumps({"channels": args.channels, "audio_s_per_channel": args.seconds, "wall_s": wall, "pool_size": args.pool_size,
                              "model": args.model, "engine": "fake" if args.fake_engine else "b200", "decode_profile": args.decode_profile,
                              "audio_s_per_s": args.channels * args.seconds / wall, "load_test_exit": code, "summary": summary}))
    finally:
        server.terminate()
        try:
            server.wait(timeout=40)
        except subprocess.TimeoutExpired:
            server.kill()


if __name__ == "__main__":
    main()
