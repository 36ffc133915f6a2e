"""Streaming-session load against one GPU: the decode schedule the reference's stream orchestrator produces
(SURVEY 8(d) configs 2 / 5), driven at the `ModelBackend.transcribe` boundary in real time.

Per session (one host thread = one pool handle, model_registry.py:564-606): speech bursts of U(1, 8) s separated by
U(0.6, 1.2) s gaps; while a burst is running a PARTIAL decode of its last <= 10 s every 1.5 s
(config/server.yaml:31-32), at most one decode in flight per session (a partial whose slot was missed is skipped,
model_registry.py:631-680); a FINAL decode of the whole burst when it ends.  Latency = submit -> result
(queue wait + inference, what the server reports as stt-decode-* metadata); p95 is nearest-rank
(tools/bench/grpc_load_test.py:1028-1032).  Random-init weights never emit EOT (and with an EOT logit offset they
emit it either always or never: their logits hardly depend on the token history), so every decode is bounded with
upstream's `DecodingOptions.sample_len` at `tokens_per_s` tokens per second of audio (3.5 ~ conversational English
incl. timestamp tokens) + 4; `--tokens-per-s 0` removes the bound (224 steps per decode, the worst case of bench.py).

Usage: python tools/stream_bench.py [--model large-v3] [--sessions 64 128] [--seconds 20] [--tokens-per-s 3.5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

REALTIME = {"beam_size": 1, "best_of": 1, "patience": 1.0, "temperature": 0.0, "length_penalty": 1.0,
            "without_timestamps": True, "compression_ratio_threshold": 2.4, "no_speech_threshold": 0.6,
            "log_prob_threshold": -1.0, "language": "en", "task": "transcribe"}


def nearest_rank(values, q):
    if not values:
        return None
    v = sorted(values)
    return v[max(0, int(np.ceil(q * len(v))) - 1)]


def run_stream_sim(handles, seconds: float, seed: int = 0, partial_every: float = 1.5, window_s: float = 10.0, opts=None,
                   tokens_per_s: float = 3.5, finals_only: bool = False):
    """handles[i] serves session i.  Returns a dict of latency / throughput statistics.
    finals_only: VAD endpointing without interim results (BASELINE configs[2]): one decode per burst, at its end."""
    from b200_whisper.synth import synth_audio

    opts = dict(opts or REALTIME)
    n = len(handles)
    rng = np.random.default_rng(seed)
    # schedule: list of (burst_start, burst_end) per session within [0, seconds)
    plans = []
    for i in range(n):
        t = float(rng.uniform(0.0, partial_every))
        bursts = []
        while t < seconds:
            length = float(rng.uniform(1.0, 8.0))
            bursts.append((t, min(t + length, seconds)))
            t += length + float(rng.uniform(0.6, 1.2))
        plans.append(bursts)
    audio = [synth_audio(seed * 1000 + i, 9.0) for i in range(n)]  # sliced / tiled per decode

    def clip(i, dur):
        k = max(1, int(round(dur * 16000)))
        a = audio[i]
        return a[:k] if k <= a.size else np.resize(a, k)

    def slen(dur):
        return int(tokens_per_s * dur) + 4 if tokens_per_s > 0 else 0

    partial_lat, final_lat, steps, audio_s = [], [], [], [0.0]
    skipped = [0]
    lock = threading.Lock()
    t0 = time.perf_counter() + 0.2

    def session(i):
        h = handles[i]
        for (b0, b1) in plans[i]:
            k = 1
            while not finals_only:
                due = b0 + k * partial_every
                if due >= b1:
                    break
                now = time.perf_counter() - t0
                if now > due + 0.05:  # the previous decode overran this slot: skip it (one in flight per session)
                    with lock:
                        skipped[0] += 1
                    k = int((now - b0) / partial_every) + 1
                    continue
                time.sleep(max(0.0, due - now))
                dur = min(window_s, due - b0)
                ts = time.perf_counter()
                res = h.transcribe_raw(clip(i, dur), sample_len=slen(dur), **h._normalize_options(opts))
                lat = time.perf_counter() - ts
                with lock:
                    partial_lat.append(lat)
                    audio_s[0] += dur
                    steps.append(sum(len(s["tokens"]) for s in res["segments"]))
                k += 1
            now = time.perf_counter() - t0
            time.sleep(max(0.0, b1 - now))
            dur = min(20.0, b1 - b0)  # max_buffer_sec
            ts = time.perf_counter()
            h.transcribe_raw(clip(i, dur), sample_len=slen(dur), **h._normalize_options(opts))
            lat = time.perf_counter() - ts
            with lock:
                final_lat.append(lat)
                audio_s[0] += dur

    s0 = handles[0].engine.stats()
    threads = [threading.Thread(target=session, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    wall = time.perf_counter() - t0
    s1 = handles[0].engine.stats()
    ds = {k: s1[k] - s0[k] for k in ("windows", "encoder_batches", "decode_steps", "rows")}
    sched = {"windows_per_encoder_batch": ds["windows"] / max(1, ds["encoder_batches"]),
             "rows_per_decoder_step": ds["rows"] / max(1, ds["decode_steps"]), "decoder_steps_per_s": ds["decode_steps"] / wall,
             "encoder_batches_per_s": ds["encoder_batches"] / wall}
    return {"scheduler": sched,"sessions": n, "stream_s": seconds, "wall_s": wall, "partials": len(partial_lat), "finals": len(final_lat),
            "skipped_partial_slots": skipped[0],
            "p50_partial_latency_s": nearest_rank(partial_lat, 0.50), "p95_partial_latency_s": nearest_rank(partial_lat, 0.95),
            "p99_partial_latency_s": nearest_rank(partial_lat, 0.99), "p95_final_latency_s": nearest_rank(final_lat, 0.95),
            "decoded_audio_s_per_s": audio_s[0] / wall, "mean_tokens_per_partial": float(np.mean(steps)) if steps else None,
            "tokens_per_audio_s": tokens_per_s,
            "schedule": f"partial every {partial_every} s over the last <= {window_s:.0f} s of a burst, final at burst end"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--sessions", type=int, nargs="+", default=[128])
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--tokens-per-s", type=float, default=3.5)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args()
    from b200_whisper.backend import B200WhisperBackend

    smax = max(args.sessions)
    spec = f"random:{args.model}:0:0.1"
    handles = [B200WhisperBackend(spec, args.device, "bfloat16", max_segments=min(smax, 192), max_sequences=max(2 * min(smax, 192), 8),
                                  max_encoder_batch=16) for _ in range(smax)]
    run_stream_sim(handles[: min(8, smax)], 3.0, seed=99)  # warm-up (graphs, lazy init)
    for n in args.sessions:
        out = run_stream_sim(handles[:n], args.seconds, seed=n, tokens_per_s=args.tokens_per_s)
        out["model"] = args.model
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
