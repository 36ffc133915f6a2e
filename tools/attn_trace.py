"""Timeline of CTA 0 of the persistent encoder attention kernel (clock64 stores into fixed slots).
Usage: python tools/attn_trace.py [batch] [T]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper import _lib as L  # noqa: E402

lib = L.load()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
H = 20
d = 64 * H
qkv = torch.randn((batch * T, 3 * d), device="cuda").bfloat16()
out = torch.zeros((batch * T, d), device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "attn")
torch.cuda.synchronize()
L.check(lib.bw_debug_trace(None, 1, None, 0, None), "trace on")
L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "attn")
torch.cuda.synchronize()
cap = 1 << 15
rec = np.zeros((cap, 2), dtype=np.uint64)
L.check(lib.bw_debug_trace(None, 2, rec.ctypes.data_as(C.POINTER(C.c_uint64)), cap, None), "trace dump")
raw = rec.reshape(-1)[1: 1 + 5 * 40 * 20].reshape(5, 40, 20).astype(np.int64)
EV = {1: "tma Q issued", 2: "tma kv stage free", 3: "mma S issued", 4: "mma P ready", 10: "enter", 11: "S ready", 12: "S loaded",
      13: "max exch", 14: "O ready", 15: "pingpong go", 16: "exp done", 17: "P arrived", 18: "epi O ready", 19: "epi done"}
WARP = ["tma", "mmaA", "smA", "smB", "mmaB"]
ev = [(raw[w, t, e], WARP[w], EV.get(e, e), t) for w in range(5) for t in range(40) for e in range(20) if raw[w, t, e] > 0]
ev.sort()
t0 = ev[0][0]
prev = {}
for c, w, name, t in ev:
    d = c - prev.get(w, c)
    prev[w] = c
    print(f"{c - t0:8d} cyc  (+{d:5d})  {w:5s} {name:18s} tile#{t}")
