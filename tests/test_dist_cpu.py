"""N > 1 path on CPU: world_size-2 gloo run of bench.py's sharding + reduction logic (sessions shard across ranks,
no data-path collective; whole-job RTFx = sum of audio / slowest rank)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    lengths = bench.window_lengths(rank, 16)
    audio = float(sum(lengths))
    seconds = 1.0 + rank  # rank 1 is the slow one
    rtfx = bench.whole_job_rtfx(dist, rank, audio, seconds)
    mx = bench.barrier_max(dist, rank, seconds)
    out.put((rank, lengths, audio, rtfx, mx))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, l0, a0, v0, m0), (r1, l1, a1, v1, m1) = res
    assert l0 != l1 and all(2.0 <= x <= 10.0 for x in l0 + l1)  # each rank owns different sessions
    assert m0 == m1 == 2.0
    assert v0 == pytest.approx((a0 + a1) / 2.0) and v1 == pytest.approx(v0)


def test_reference_arm_is_rank0_only(monkeypatch):
    """under torchrun only rank 0 runs the CPU reference; other ranks exit without work"""
    sys.path.insert(0, ROOT)
    import bench

    called = []
    monkeypatch.setattr(bench, "cpu_oracle_window", lambda *a, **k: called.append(1) or (1.0, 1))
    monkeypatch.setattr(bench, "emit", lambda line: called.append(line))
    args = bench.resolve_config(type("A", (), {"gpus": 2, "steps": 1, "warmup": 0, "model": "test-tiny", "cpu_threads": 1, "sessions": 128,
                                               "config": 4})())
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    bench.run_reference(args)
    assert called == []
    monkeypatch.setenv("RANK", "0")  # rank 0 runs it alone, with no process group, and prints one line on our arm's workload
    bench.run_reference(args)
    assert called[0] == 1 and len(called) == 2
    import json

    line = json.loads(called[1])
    assert line["impl"] == "reference" and line["config"]["workload"] == bench.workload_name(args) and line["n_gpus"] == 2
    assert "configs[4]" in line["config"]["workload"] and "faster_whisper_int8" in line
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
