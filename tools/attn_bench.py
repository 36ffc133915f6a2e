"""Encoder attention kernel alone: CUDA-event timing + a target for `ncu --set full`.
Usage: python tools/attn_bench.py [batch] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper import _lib as L  # noqa: E402

lib = L.load()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
H = 20
d = 64 * H
g = torch.Generator(device="cuda").manual_seed(5)
qkv = torch.randn((batch * T, 3 * d), device="cuda", generator=g).bfloat16()
out = torch.zeros((batch * T, d), device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "attn")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "attn")
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
flops = 4.0 * batch * H * T * T * 64
print(f"attention batch {batch} T {T}: {ms * 1e3:.1f} us, {flops / ms / 1e9:.1f} TFLOP/s, {batch * H * T * 1536 / ms / 1e9 * 1e3 / 1e3:.2f} Gexp/s")
