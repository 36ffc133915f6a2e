"""In-situ timeline of the batched decoder step: which phase of which kernel the time goes to.
Usage: B200W_NO_GRAPH=1 python tools/trace_step.py [model] [segments] [n_group] [context] > gpurun_out/trace.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200W_NO_GRAPH", "1")
from b200_whisper.backend import B200WhisperBackend  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
segments = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n_group = int(sys.argv[3]) if len(sys.argv) > 3 else 1
context = int(sys.argv[4]) if len(sys.argv) > 4 else 100
b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=segments, max_sequences=max(8, segments * n_group),
                       max_encoder_batch=1)
eng = b.engine
ms, by = eng.bench_decoder_step(segments, n_group, context, 4)
print(f"# untraced decoder step {segments}x{n_group}: {ms:.3f} ms")
eng.trace_begin()
ms, by = eng.bench_decoder_step(segments, n_group, context, 2)
rec = eng.trace_end()
print(f"# traced decoder step: {ms:.3f} ms, {len(rec)} records")
NAMES = {1: "gemm_splitk", 2: "layernorm", 3: "self_attn", 4: "xattn_first", 5: "xattn_last", 6: "gemm_plain"}
order = np.argsort(rec[:, 1], kind="stable")
rec = rec[order]
t0 = int(rec[0, 1]) if len(rec) else 0
rows = []
for tag, t in rec:
    tag = int(tag); t = int(t)
    rows.append((t - t0, NAMES.get((tag >> 24) & 0xff, "?"), (tag >> 8) & 0xffff, tag & 0xff, tag >> 32))
# print the third quarter of the records (a steady-state stretch in the middle of the last step)
lo, hi = len(rows) * 5 // 8, len(rows) * 5 // 8 + 260
prev = None
for r in rows[lo:hi]:
    dt = 0 if prev is None else r[0] - prev
    prev = r[0]
    print(f"{r[0] / 1e3:10.2f} us  +{dt / 1e3:6.2f}  {r[1]:12s} grid.x={r[2]:<4d} phase={r[3]} sm={r[4]}")
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace_step.npy"), rec)
