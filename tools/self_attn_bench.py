"""Decoder self-attention kernels in isolation through bw_test_dec_self_attention (run under ncu: the launch list gives
each kernel's own duration).  Usage: ncu --metrics gpu__time_duration.sum --csv python tools/self_attn_bench.py [rows] [beam] [ctx]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper import _lib as L  # noqa: E402

n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 100
modes = [int(m) for m in os.environ.get("MODES", "1,2,3").split(",")]
n_head, n_layer, n_ctx, P = 20, 4, 448, 16
d = 64 * n_head
dev = "cuda:0"
lib = L.load()
S = n_req * G
nb = (n_ctx + P - 1) // P
used = (ctx + P) // P
gen = torch.Generator(device=dev).manual_seed(1)
n_pages = S * used
pool = (torch.randn((n_pages, n_layer, 2, P, d), device=dev, generator=gen)).bfloat16()
pt = torch.zeros((S, nb), dtype=torch.int32)
pt[:, :used] = torch.randperm(n_pages, generator=torch.Generator().manual_seed(2)).reshape(S, used).to(torch.int32)
anc = torch.randint(0, G, (S, n_ctx), generator=torch.Generator().manual_seed(3)).to(torch.uint8)
if G == 1:
    anc.zero_()
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
seq = list(range(S))
rs, rp, rb = i32(seq), i32([ctx] * S), i32([ctx] * S)
rpage = i32([int(pt[s, ctx // P]) for s in seq])
sf = i32([(s // G) * G | (0x40000000 if G == 1 else 0) for s in seq])
anc_d, pt_d = anc.to(dev), pt.to(dev)
qkv = torch.randn((S, 3 * d), device=dev, generator=gen)
out = torch.empty((S, d), device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
for mode in modes:
    L.check(lib.bw_test_self_attention_mode(mode), "mode")
    for layer in (1, 2, 3):
        L.check(lib.bw_test_dec_self_attention(S, rs.data_ptr(), rp.data_ptr(), rb.data_ptr(), rpage.data_ptr(), qkv.data_ptr(), pool.data_ptr(),
                                               n_layer, n_ctx, S, pt_d.data_ptr(), sf.data_ptr(), anc_d.data_ptr(), layer, d, n_head, ctx + 1,
                                               out.data_ptr(), None), "bw_test_dec_self_attention")
torch.cuda.synchronize()
print("ok", S, "rows", ctx, "ctx", 2 * 2 * ctx * d * S / 1e6, "MB per layer")
