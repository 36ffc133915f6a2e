"""ONE server process, one engine per GPU (north_star: "the model pool is sharded one replica per GPU ... sessions are
placed on GPUs by the pool"): the deployment shape behind `device="cuda:N"` / `"cuda:auto"` (SURVEY 8(e) options A / B),
as opposed to bench.py's process-per-GPU torchrun line.  Runs bench.py's configs[4] step -- 128 sessions per GPU, one
2-10 s partial window each, full 224-step decodes -- on every visible GPU from this one Python process, two ways:

  threads : one host thread per session calling `transcribe` (what the reference's pool does: pool_size threads)
  many    : one host thread per GPU calling `transcribe_many` (one blocking bw_decode_many per round of windows)

and prints audio-s/s for each, next to the same step on GPU 0 alone, so the efficiency of the in-process shape and its
limiter (the GIL / thread fan-out on the host) can be read off.

Usage: python tools/multi_gpu_inproc.py [--gpus N] [--sessions 128] [--steps 3] [--model large-v3]"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ENC_BATCH, REALTIME, window_lengths  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--sessions", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--device-mode", default="explicit", choices=["explicit", "auto"])
    args = ap.parse_args()
    from b200_whisper import _lib
    from b200_whisper.backend import B200WhisperBackend
    from b200_whisper.synth import synth_audio

    n_gpus = args.gpus or _lib.load().bw_device_count()
    S = args.sessions
    spec = f"random:{args.model}:0:0.1"
    t0 = time.time()
    handles = []  # handles[g][i]
    if args.device_mode == "auto":
        # option B: every pool handle is created with the SAME arguments (model_registry.py:230-247) and the backend
        # spreads them round-robin over the GPUs
        flat = [B200WhisperBackend(spec, "cuda:auto", "bfloat16", max_segments=S, max_sequences=2 * S, max_encoder_batch=min(ENC_BATCH, S))
                for _ in range(S * n_gpus)]
        handles = [[h for h in flat if h.device_index == g] for g in range(n_gpus)]
        assert all(len(hs) == S for hs in handles), [len(hs) for hs in handles]
    else:
        for g in range(n_gpus):
            handles.append([B200WhisperBackend(spec, f"cuda:{g}", "bfloat16", max_segments=S, max_sequences=2 * S,
                                               max_encoder_batch=min(ENC_BATCH, S)) for _ in range(S)])
    load_s = time.time() - t0
    audios = [[synth_audio(g * 100000 + i, l) for i, l in enumerate(window_lengths(g, S))] for g in range(n_gpus)]
    audio_s = [sum(a.size for a in au) / 16000.0 for au in audios]

    def step_threads(gpus):
        th = [threading.Thread(target=handles[g][i].transcribe, args=(audios[g][i], REALTIME)) for g in gpus for i in range(S)]
        t = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        return time.perf_counter() - t

    def step_many(gpus):
        th = [threading.Thread(target=handles[g][0].transcribe_many, args=(audios[g], REALTIME)) for g in gpus]
        t = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        return time.perf_counter() - t

    out = {"gpus": n_gpus, "sessions_per_gpu": S, "model": args.model, "device_mode": args.device_mode, "load_s": load_s,
           "host_cores": os.cpu_count()}
    for name, fn in (("threads", step_threads), ("many", step_many)):
        for label, gpus in (("1gpu", [0]), ("all", list(range(n_gpus)))):
            for _ in range(args.warmup):
                fn(gpus)
            dt = [fn(gpus) for _ in range(args.steps)]
            out[f"{name}_{label}_audio_s_per_s"] = sum(audio_s[g] for g in gpus) * len(dt) / sum(dt)
        out[f"{name}_efficiency"] = out[f"{name}_all_audio_s_per_s"] / (n_gpus * out[f"{name}_1gpu_audio_s_per_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
