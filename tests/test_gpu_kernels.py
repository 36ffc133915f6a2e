"""Kernel-level GPU tests through the C ABI: tcgen05 GEMM / attention vs torch on the same inputs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

torch = pytest.importorskip("torch")


def _lib():
    from b200_whisper import _lib as L

    return L, L.load()


def _gemm(impl, A, B, bias=None, residual=None, gelu=False, out_fp32=False):
    L, lib = _lib()
    M, K = A.shape
    N = B.shape[0]
    Cout = torch.empty((M, N), device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    torch.cuda.synchronize()
    st = lib.bw_gemm_bf16(impl, A.data_ptr(), B.data_ptr(), Cout.data_ptr(), bias.data_ptr() if bias is not None else None,
                          residual.data_ptr() if residual is not None else None, M, N, K, int(gelu), int(out_fp32), None)
    L.check(st, "bw_gemm_bf16")
    torch.cuda.synchronize()
    return Cout


def _ref(A, B, bias, residual, gelu):
    r = A.float() @ B.float().t()
    if bias is not None:
        r = r + bias
    if gelu:
        r = torch.nn.functional.gelu(r)
    if residual is not None:
        r = r + residual
    return r


SHAPES = [(128, 128, 64), (256, 256, 128), (1500, 384, 384), (3000, 384, 240), (200, 1536, 384), (1500, 1280, 1280),
          (333, 2560, 640), (1500, 5120, 1280), (4500, 1280, 5120)]


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_plain(impl, shape):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn((N, K), device="cuda", generator=g) * 0.5).bfloat16()
    out = _gemm(impl, A, B, out_fp32=True)
    ref = _ref(A, B, None, None, False)
    err = (out - ref).abs().max().item()
    assert err <= 2e-2 * (K ** 0.5) * 0.25 + 1e-3, f"impl {impl} shape {shape}: max abs err {err}"
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, f"impl {impl} shape {shape}: rel-L2 {rel}"


@pytest.mark.parametrize("impl", [0, 1, 2])
def test_gemm_epilogues(impl):
    M, N, K = 1500, 768, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.3).bfloat16()
    B = (torch.randn((N, K), device="cuda", generator=g) * 0.3).bfloat16()
    bias = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((M, N), device="cuda", generator=g)
    for gelu in (False, True):
        for use_res in (False, True):
            for out_fp32 in (True, False):
                out = _gemm(impl, A, B, bias, res if use_res else None, gelu, out_fp32).float()
                ref = _ref(A, B, bias, res if use_res else None, gelu)
                tol = 3e-2 if not out_fp32 else 5e-3
                rel = ((out - ref).norm() / ref.norm()).item()
                assert rel < tol, f"impl {impl} gelu={gelu} res={use_res} fp32={out_fp32}: rel-L2 {rel}"


@pytest.mark.parametrize("rows", [1, 5, 16, 37, 128, 300])
def test_gemm_swap_ab_skinny(rows):
    """decoder path: few rows against a big [N, K] weight (incl. the 51866-row tied embedding)"""
    for N, K in ((1280, 1280), (51866, 384)):
        g = torch.Generator(device="cuda").manual_seed(rows + N)
        A = (torch.randn((rows, K), device="cuda", generator=g) * 0.3).bfloat16()
        B = (torch.randn((N, K), device="cuda", generator=g) * 0.3).bfloat16()
        out = _gemm(2, A, B, out_fp32=True)
        ref = _ref(A, B, None, None, False)
        rel = ((out - ref).norm() / ref.norm()).item()
        assert rel < 1e-3, f"rows {rows} N {N}: rel-L2 {rel}"


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("cfg", [(1, 1500, 2), (2, 1500, 6), (1, 300, 1)])
def test_encoder_attention(impl, cfg):
    L, lib = _lib()
    batch, T, H = cfg
    d = 64 * H
    g = torch.Generator(device="cuda").manual_seed(11 + T + H)
    qkv = (torch.randn((batch * T, 3 * d), device="cuda", generator=g)).bfloat16()
    out = torch.zeros((batch * T, d), device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    L.check(lib.bw_attention_bf16(impl, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "bw_attention_bf16")
    torch.cuda.synchronize()
    q, k, v = [t.reshape(batch, T, H, 64).permute(0, 2, 1, 3).float() for t in qkv.split(d, dim=1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(batch * T, d)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 2e-2, f"impl {impl} cfg {cfg}: rel-L2 {rel}"


def _attention_ref(qkv, batch, T, H):
    d = 64 * H
    q, k, v = [t.reshape(batch, T, H, 64).permute(0, 2, 1, 3).float() for t in qkv.split(d, dim=1)]
    return torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(batch * T, d)


@pytest.mark.parametrize("cfg", [(4, 1500, 8), (3, 700, 20), (2, 129, 3), (5, 1024, 7)])
def test_encoder_attention_persistent_many_items(cfg):
    """more (window, head, query-pair) items than SMs: every persistent CTA walks over several items, with ragged
    last key / query tiles (the barrier phases and the K / V rings run across item boundaries)"""
    L, lib = _lib()
    batch, T, H = cfg
    d = 64 * H
    g = torch.Generator(device="cuda").manual_seed(100 + T + H)
    qkv = torch.randn((batch * T, 3 * d), device="cuda", generator=g).bfloat16()
    out = torch.full((batch * T, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    for _ in range(2):  # twice: results must not depend on what the previous launch left in TMEM / shared memory
        L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "bw_attention_bf16")
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, batch, T, H)
    assert torch.isfinite(out.float()).all()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 2e-2, f"cfg {cfg}: rel-L2 {rel}"
    # per-window error too: a wrong item mapping would hide in the global norm of a large batch
    per = (out.float() - ref).reshape(batch, -1).norm(dim=1) / ref.reshape(batch, -1).norm(dim=1)
    assert per.max().item() < 2e-2, per


def test_encoder_attention_lazy_rescale():
    """scores whose row maximum keeps jumping by far more than 2^8 from key tile to key tile: the accumulators in
    TMEM are rescaled in place (tcgen05.ld / st) whenever the reference maximum is raised"""
    L, lib = _lib()
    batch, T, H = 2, 1500, 4
    d = 64 * H
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn((batch * T, 3 * d), device="cuda", generator=g)
    ramp = torch.linspace(0.5, 8.0, T, device="cuda").repeat(batch)[:, None]
    qkv[:, :d] *= 3.0
    qkv[:, d:2 * d] *= ramp  # later keys score higher: the running maximum grows all the way through
    qkv = qkv.bfloat16()
    out = torch.zeros((batch * T, d), device="cuda", dtype=torch.bfloat16)
    L.check(lib.bw_attention_bf16(0, qkv.data_ptr(), out.data_ptr(), batch, T, H, None), "bw_attention_bf16")
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, batch, T, H)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 2e-2, f"rel-L2 {rel}"


@pytest.mark.parametrize("shape", [(128, 1280, 1280), (100, 5120, 1280), (128, 1280, 5120), (300, 1280, 512), (7, 384, 1536), (1, 3840, 1280)])
def test_gemm_cluster_splitk(shape):
    """few output tiles -> K is split over a thread-block cluster and reduced through DSMEM (decoder GEMMs)"""
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.3).bfloat16()
    B = (torch.randn((N, K), device="cuda", generator=g) * 0.3).bfloat16()
    bias = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((M, N), device="cuda", generator=g)
    for gelu, use_res, out_fp32 in ((False, False, True), (True, False, False), (False, True, True), (False, True, False)):
        out = _gemm(0, A, B, bias, res if use_res else None, gelu, out_fp32).float()
        ref = _ref(A, B, bias, res if use_res else None, gelu)
        rel = ((out - ref).norm() / ref.norm()).item()
        assert rel < (5e-3 if out_fp32 else 3e-2), f"{shape} gelu={gelu} res={use_res} fp32={out_fp32}: rel-L2 {rel}"
    # in-place residual (C == residual), as the decoder uses it
    x = res.clone()
    L, lib = _lib()
    L.check(lib.bw_gemm_bf16(0, A.data_ptr(), B.data_ptr(), x.data_ptr(), bias.data_ptr(), x.data_ptr(), M, N, K, 0, 1, None), "gemm")
    torch.cuda.synchronize()
    ref = _ref(A, B, bias, res, False)
    assert ((x - ref).norm() / ref.norm()).item() < 5e-3


@pytest.mark.parametrize("shape", [(3000, 1280, 1280), (6000, 3840, 1280), (2999, 1000, 200), (24000, 1280, 384), (4500, 5120, 1280)])
def test_gemm_cta_pair_persistent(shape):
    """>= 60 pair tiles -> persistent cta_group::2 kernel (256x256 per CTA pair, double-buffered TMEM)"""
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.3).bfloat16()
    B = (torch.randn((N, K), device="cuda", generator=g) * 0.3).bfloat16()
    bias = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((M, N), device="cuda", generator=g)
    for gelu, use_res, out_fp32 in ((False, False, True), (True, False, False), (False, True, True)):
        out = _gemm(0, A, B, bias, res if use_res else None, gelu, out_fp32).float()
        ref = _ref(A, B, bias, res if use_res else None, gelu)
        rel = ((out - ref).norm() / ref.norm()).item()
        assert rel < (5e-3 if out_fp32 else 3e-2), f"{shape} gelu={gelu} res={use_res} fp32={out_fp32}: rel-L2 {rel}"
        assert torch.isfinite(out).all()


def test_gemm_throughput_report():
    """not a pass/fail perf gate: prints TFLOP/s of the encoder GEMM shapes (CUDA events)"""
    L, lib = _lib()
    for M, N, K in ((24000, 5120, 1280), (24000, 1280, 5120), (24000, 3840, 1280), (24000, 1280, 1280), (6000, 5120, 1280)):
        A = torch.randn((M, K), device="cuda").bfloat16()
        B = torch.randn((N, K), device="cuda").bfloat16()
        Cout = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            lib.bw_gemm_bf16(0, A.data_ptr(), B.data_ptr(), Cout.data_ptr(), None, None, M, N, K, 0, 0, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            lib.bw_gemm_bf16(0, A.data_ptr(), B.data_ptr(), Cout.data_ptr(), None, None, M, N, K, 0, 0, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            torch.matmul(A, B.t(), out=Cout)
        t1.record()
        torch.cuda.synchronize()
        ms_ref = t0.elapsed_time(t1) / 10
        print(f"GEMM {M}x{N}x{K}: ours {2 * M * N * K / ms / 1e9:.0f} TFLOP/s ({ms:.3f} ms), cuBLAS {2 * M * N * K / ms_ref / 1e9:.0f} TFLOP/s")


@pytest.mark.parametrize("shape", [(128, 1280, 1280, 3840), (37, 1280, 5120, 1280), (300, 384, 384, 1536), (128, 1280, 0, 5120), (5, 512, 512, 512),
                                   # beam batches: several M tiles x a wide consumer -> the unsplit (KS = 1) single-wave path
                                   (320, 1280, 1280, 5120), (320, 1280, 0, 3840), (200, 1280, 1280, 5120), (257, 1280, 0, 3840)])
@pytest.mark.parametrize("gelu", [False, True])
def test_layernorm_fused_row_gemms(shape, gelu):
    """decoder LayerNorm fusion: producer GEMM (x = res + A.Wp^T + b, plus bf16(x) and LayerNorm partials) feeding a
    consumer GEMM whose weight carries gamma and whose epilogue applies mean / rstd, vs torch LayerNorm + Linear"""
    L, lib = _lib()
    M, d, Kp, N = shape
    g = torch.Generator(device="cuda").manual_seed(M + d + Kp + N)
    res = torch.randn((M, d), device="cuda", generator=g) * 2.0 + 0.7          # non-zero mean rows
    res[:, 5] += 30.0                                                            # an outlier channel, as in real residual streams
    gamma = 1.0 + 0.3 * torch.randn((d,), device="cuda", generator=g)
    beta = 0.2 * torch.randn((d,), device="cuda", generator=g)
    Wc = torch.randn((N, d), device="cuda", generator=g) * 0.05
    bc = torch.randn((N,), device="cuda", generator=g) * 0.1
    if Kp:
        A = (torch.randn((M, Kp), device="cuda", generator=g) * 0.5).bfloat16()
        Wp = (torch.randn((d, Kp), device="cuda", generator=g) * 0.05).bfloat16()
        bp = torch.randn((d,), device="cuda", generator=g) * 0.1
        x_ref = res + A.float() @ Wp.float().T + bp
    else:
        A = Wp = bp = None
        x_ref = res.clone()
    x_out = torch.empty((M, d), device="cuda")
    out = torch.empty((M, N), device="cuda")
    torch.cuda.synchronize()
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    L.check(lib.bw_test_ln_chain(ptr(A), ptr(Wp), ptr(bp), ptr(res), ptr(gamma), ptr(beta), ptr(Wc), ptr(bc), M, d, Kp, N, int(gelu),
                                 ptr(x_out), ptr(out), None), "bw_test_ln_chain")
    torch.cuda.synchronize()
    assert ((x_out - x_ref).norm() / x_ref.norm()).item() < 2e-3
    # reference on the same bf16 operand precision the product path has (bf16 activations and weights, fp32 accumulate)
    y = torch.nn.functional.layer_norm(x_ref, (d,), gamma, beta, 1e-5)
    ref = y.bfloat16().float() @ Wc.bfloat16().float().T + bc
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1.5e-2, f"{shape} gelu={gelu}: rel-L2 {rel}"
