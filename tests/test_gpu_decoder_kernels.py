"""Kernel-level parity of the bf16 PRODUCT decoder kernels against fp32 torch, through the C-ABI test hooks
(include/b200_whisper_hooks.h): the kernels every bench number rides on are checked in isolation --

* `dec_cross_attention_mma_kernel` (+ `dec_cross_combine_kernel`): hypotheses per segment 1 / 5 / 8 (and ragged groups),
  1 / 7 / 128 segments, every T-split count 1..8 incl. splits whose last tile is ragged or that are empty;
* `dec_self_attention_kernel<bf16>`: cached decode through a non-trivial ancestry table (beam reorder) at context
  1 / 100 / 447, the single-hypothesis fast path, a multi-row prefill, and the fused K/V append;
* `sample_topk_kernel`: SuppressBlank / SuppressTokens / ApplyTimestampRules + log-softmax + top-(beam + 1) on random
  logits and histories for the three vocabularies, against the oracle's filters + torch.topk.

Upstream semantics: whisper/model.py MultiHeadAttention (cached cross / self attention), whisper/decoding.py logit
filters and BeamSearchDecoder.update's topk (reached from reference stt_server/model/backends/torch_whisper.py:55).
"""
import ctypes as C

import numpy as np
import pytest

from tests._util import model_spec

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

torch = pytest.importorskip("torch")

from b200_whisper import _lib as L  # noqa: E402
from b200_whisper.backend import B200WhisperBackend  # noqa: E402
from oracle import whisper_oracle as wo  # noqa: E402
from oracle.tables import layout_for_vocab  # noqa: E402

DEV = "cuda"


def _i32(x):
    return torch.tensor(x, dtype=torch.int32, device=DEV)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


# ------------------------------------------------------------------------------------------------ cross attention
def _cross_case(n_groups, nq, n_head, n_layer, layer, force_split, seed, ragged=False, T_enc=1500):
    lib = L.load()
    d = 64 * n_head
    g = torch.Generator(device=DEV).manual_seed(seed)
    n_slots = n_groups + 3
    cache = (torch.randn((n_slots, n_layer, T_enc, 2 * d), device=DEV, generator=g) * 1.5).bfloat16()
    sizes = [(1 + (i * 5 + 2) % nq) if ragged else nq for i in range(n_groups)]
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    R = int(sum(sizes))
    perm = torch.randperm(n_slots, generator=torch.Generator().manual_seed(seed))[:n_groups].to(torch.int32)
    q = torch.randn((R, d), device=DEV, generator=g) * 2.0
    out = torch.full((R, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    gf, gn, gx = _i32(first), _i32(sizes), perm.to(DEV)
    torch.cuda.synchronize()
    L.check(lib.bw_test_dec_cross_attention(cache.data_ptr(), n_slots, n_layer, layer, T_enc, d, n_head, q.data_ptr(), gf.data_ptr(),
                                            gn.data_ptr(), gx.data_ptr(), n_groups, max(sizes), R, force_split, out.data_ptr(), None),
            "bw_test_dec_cross_attention")
    torch.cuda.synchronize()
    # fp32 reference: softmax(q . K^T / 8) . V per head, on the cached K/V of the row's segment
    ref = torch.empty((R, d), device=DEV)
    for gi in range(n_groups):
        kv = cache[int(perm[gi]), layer].float()
        k = kv[:, :d].reshape(T_enc, n_head, 64).permute(1, 0, 2)
        v = kv[:, d:].reshape(T_enc, n_head, 64).permute(1, 0, 2)
        rows = slice(int(first[gi]), int(first[gi]) + sizes[gi])
        qq = q[rows].reshape(-1, n_head, 64).permute(1, 0, 2)
        w = torch.softmax(qq @ k.transpose(1, 2) / 8.0, dim=-1)
        ref[rows] = (w @ v).permute(1, 0, 2).reshape(-1, d)
    assert torch.isfinite(out.float()).all(), "rows left unwritten"
    return out, ref, first, sizes


@pytest.mark.parametrize("nq", [1, 5, 8])
@pytest.mark.parametrize("n_groups", [1, 7, 128])
def test_cross_attention_mma_rows_and_segments(nq, n_groups):
    out, ref, first, sizes = _cross_case(n_groups, nq, n_head=6, n_layer=2, layer=1, force_split=0, seed=nq * 131 + n_groups)
    r = _rel(out, ref)
    assert r < 1.5e-2, f"nq {nq} segments {n_groups}: rel-L2 {r}"
    per_row = (out.float() - ref).norm(dim=1) / ref.norm(dim=1)
    assert per_row.max().item() < 4e-2, f"worst row {per_row.max().item()}"  # a wrong group mapping hides in the global norm


@pytest.mark.parametrize("n_split", [1, 2, 3, 4, 5, 6, 7, 8])
def test_cross_attention_mma_every_split_count(n_split):
    """T = 1500 keys over n_split chunks of whole 64-key tiles: 1500 = 23 tiles + 28 keys (ragged last tile); 5 and 7
    splits leave a trailing EMPTY split (chunk 320 / 256 keys) whose (-inf, 0) partial the combine kernel must ignore."""
    out, ref, _, _ = _cross_case(3, 5, n_head=20, n_layer=3, layer=2, force_split=n_split, seed=500 + n_split, ragged=True)
    r = _rel(out, ref)
    assert r < 1.5e-2, f"{n_split} splits: rel-L2 {r}"
    base, _, _, _ = _cross_case(3, 5, n_head=20, n_layer=3, layer=2, force_split=1, seed=500 + n_split, ragged=True)
    assert _rel(out, base) < 6e-3, "split partials do not recombine to the single-pass result"


def test_cross_attention_mma_short_encoder_context():
    """T_enc below one split chunk and not a multiple of the tile: masking of the keys past T_enc inside the last tile"""
    for T_enc, split in ((100, 0), (65, 2), (1500 - 37, 8)):
        out, ref, _, _ = _cross_case(4, 8, n_head=6, n_layer=1, layer=0, force_split=split, seed=T_enc, T_enc=T_enc)
        assert _rel(out, ref) < 1.5e-2, f"T_enc {T_enc}"


# ------------------------------------------------------------------------------------------------- self attention
P = 16  # kPageTokens


def _paged_pool(S, n_layer, n_ctx, d, gen, seed):
    """pool [n_pages][L][2][P][d] bf16 + a page table [S][n_blocks] that scatters every (slot, block) onto a random page"""
    nb = (n_ctx + P - 1) // P
    n_pages = S * nb + 5
    pool = (torch.randn((n_pages, n_layer, 2, P, d), device=DEV, generator=gen) * 1.3).bfloat16()
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(seed))[: S * nb].reshape(S, nb).to(torch.int32)
    return pool, perm, nb


def _kv_at(pool, pt, u, layer, kv, t):
    return pool[int(pt[u, t // P]), layer, kv, t % P]


def _self_ref(qkv, pool_before, pt, rows, seq_first, anc, layer, d, n_head):
    """fp32 reference of one decoder self-attention step over the paged pool: row r attends positions [0, pos_r]; positions
    < bpos come from the pool through the ancestry table + page table, the others from this step's qkv rows (bf16-rounded
    like the append)"""
    R = qkv.shape[0]
    out = torch.empty((R, d), device=DEV)
    kb = qkv[:, d:2 * d].bfloat16().float()
    vb = qkv[:, 2 * d:].bfloat16().float()
    for r in range(R):
        s, pos, bpos = rows["seq"][r], rows["pos"][r], rows["bpos"][r]
        first = seq_first[s]
        ks, vs = [], []
        for t in range(pos + 1):
            if t < bpos:
                u = first + int(anc[s, t])
                ks.append(_kv_at(pool_before, pt, u, layer, 0, t).float())
                vs.append(_kv_at(pool_before, pt, u, layer, 1, t).float())
            else:
                rr = r - (pos - t)
                ks.append(kb[rr])
                vs.append(vb[rr])
        K = torch.stack(ks).reshape(-1, n_head, 64).permute(1, 0, 2)
        V = torch.stack(vs).reshape(-1, n_head, 64).permute(1, 0, 2)
        qq = qkv[r, :d].reshape(n_head, 1, 64)
        w = torch.softmax(qq @ K.transpose(1, 2) / 8.0, dim=-1)
        out[r] = (w @ V).reshape(d)
    return out


def _run_self(rows, qkv, pool, pt, seq_first_dev_vals, anc, layer, d, n_head, n_ctx, n_layer, max_ctx=None):
    lib = L.load()
    R = qkv.shape[0]
    out = torch.full((R, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    rs, rp, rb = _i32(rows["seq"]), _i32(rows["pos"]), _i32(rows["bpos"])
    rpage = _i32([int(pt[s, t // P]) for s, t in zip(rows["seq"], rows["pos"])])  # the row's k / v goes to its slot's page
    sf = _i32(seq_first_dev_vals)
    anc_d = anc.to(torch.uint8).to(DEV).contiguous()
    pt_d = pt.to(DEV).contiguous()
    torch.cuda.synchronize()
    L.check(lib.bw_test_dec_self_attention(R, rs.data_ptr(), rp.data_ptr(), rb.data_ptr(), rpage.data_ptr(), qkv.data_ptr(), pool.data_ptr(),
                                           n_layer, n_ctx, pt.shape[0], pt_d.data_ptr(), sf.data_ptr(), anc_d.data_ptr(), layer, d, n_head,
                                           max(rows["pos"]) + 1 if max_ctx is None else max_ctx, out.data_ptr(), None),
            "bw_test_dec_self_attention")
    torch.cuda.synchronize()
    return out


SELF_MODES = {"staged": 1, "warp": 2, "persistent": 3}  # bw_test_self_attention_mode


@pytest.fixture(params=sorted(SELF_MODES))
def self_mode(request):
    lib = L.load()
    L.check(lib.bw_test_self_attention_mode(SELF_MODES[request.param]), "bw_test_self_attention_mode")
    yield request.param
    L.check(lib.bw_test_self_attention_mode(0), "bw_test_self_attention_mode")


@pytest.mark.parametrize("ctx", [1, 31, 32, 63, 64, 100, 193, 447])  # 1 .. 15 ring items / 1 .. 4 staging passes, page and chunk edges
@pytest.mark.parametrize("G", [1, 5, 8])
def test_self_attention_bf16_cached_decode_through_ancestry(ctx, G, self_mode):
    """one new token per hypothesis at position `ctx` (the step's row), `ctx` cached positions behind it.  G > 1: every
    cached position of every hypothesis lives in a RANDOM beam slot of its request (what beam reordering leaves behind);
    every (slot, block) sits on a random page of the pool."""
    n_head, n_layer, layer, n_ctx = 6, 3, 1, 448
    d = 64 * n_head
    n_req = 3
    S = n_req * G + 2
    g = torch.Generator(device=DEV).manual_seed(ctx * 17 + G)
    pool, pt, nb = _paged_pool(S, n_layer, n_ctx, d, g, ctx + 31 * G)
    pool_before = pool.clone()
    cg = torch.Generator().manual_seed(ctx + G)
    anc = torch.randint(0, G, (S, n_ctx), generator=cg)
    seq, first_of = [], {}
    for rq in range(n_req):
        f = 1 + rq * G  # requests do not start at unit 0
        for j in range(G):
            seq.append(f + j)
            first_of[f + j] = f
    rows = {"seq": seq, "pos": [ctx] * len(seq), "bpos": [ctx] * len(seq)}
    seq_first = [first_of.get(s, 0) for s in range(S)]
    flag = 0x40000000 if G == 1 else 0  # kSingleBeamFlag: ancestry identically 0, the kernel skips the table
    if G == 1:
        anc.zero_()
    qkv = torch.randn((len(seq), 3 * d), device=DEV, generator=g) * 1.5
    out = _run_self(rows, qkv, pool, pt, [f | flag for f in seq_first], anc, layer, d, n_head, n_ctx, n_layer)
    ref = _self_ref(qkv, pool_before, pt, rows, seq_first, anc, layer, d, n_head)
    assert torch.isfinite(out.float()).all()
    r = _rel(out, ref)
    assert r < 5e-3, f"ctx {ctx} G {G}: rel-L2 {r}"  # fp32 math on bf16 K/V; only the bf16 output rounding is left
    per_row = (out.float() - ref).norm(dim=1) / ref.norm(dim=1)
    assert per_row.max().item() < 2.5e-2
    # fused append: this step's k / v rows landed on the hypothesis' OWN page at position ctx, nothing else changed
    changed = (pool != pool_before)
    for i, s in enumerate(seq):
        pg = int(pt[s, ctx // P])
        assert torch.equal(pool[pg, layer, 0, ctx % P], qkv[i, d:2 * d].bfloat16()) and torch.equal(pool[pg, layer, 1, ctx % P], qkv[i, 2 * d:].bfloat16())
        changed[pg, layer, :, ctx % P] = False
    assert not changed.any(), "the kernel wrote outside this step's (page, layer, position) rows"


def test_self_attention_bf16_many_units_per_cta(self_mode):
    """more (row, head) units than resident warps: the persistent kernel walks several units per warp, with contexts of
    different lengths (1 .. 10 ring items) interleaved, so items cross unit boundaries in the ring"""
    n_head, n_layer, layer, n_ctx = 20, 2, 1, 448  # 6000 units: 3 - 4 per resident warp
    d = 64 * n_head
    S = 300
    g = torch.Generator(device=DEV).manual_seed(5)
    pool, pt, nb = _paged_pool(S, n_layer, n_ctx, d, g, 11)
    pool_before = pool.clone()
    ctxs = [(7 * i * i + 3 * i) % 200 for i in range(S)]
    ctxs[3] = 0; ctxs[50] = 0; ctxs[51] = 64; ctxs[52] = 128; ctxs[99] = 299
    rows = {"seq": list(range(S)), "pos": ctxs, "bpos": ctxs}
    anc = torch.zeros((S, n_ctx), dtype=torch.int64)
    qkv = torch.randn((S, 3 * d), device=DEV, generator=g) * 1.5
    out = _run_self(rows, qkv, pool, pt, [s | 0x40000000 for s in range(S)], anc, layer, d, n_head, n_ctx, n_layer)
    ref = _self_ref(qkv, pool_before, pt, rows, list(range(S)), anc, layer, d, n_head)
    assert torch.isfinite(out.float()).all()
    per_row = (out.float() - ref).norm(dim=1) / ref.norm(dim=1)
    assert _rel(out, ref) < 5e-3 and per_row.max().item() < 2.5e-2, (self_mode, _rel(out, ref), per_row.max().item())


def test_self_attention_bf16_prefill_rows(self_mode):
    """prefill: n rows of one sequence fed in one step (bpos = 0): causal attention among this step's own rows, plus a
    second sequence continuing from a cached prefix in the same launch"""
    n_head, n_layer, layer, n_ctx = 20, 2, 0, 448
    d = 64 * n_head
    g = torch.Generator(device=DEV).manual_seed(9)
    S = 4
    pool, pt, nb = _paged_pool(S, n_layer, n_ctx, d, g, 77)
    pool_before = pool.clone()
    n_init = 19  # spans two pages
    rows = {"seq": [2] * n_init + [0] * 3, "pos": list(range(n_init)) + [130, 131, 132], "bpos": [0] * n_init + [130] * 3}
    anc = torch.zeros((S, n_ctx), dtype=torch.int64)
    flag = 0x40000000
    qkv = torch.randn((n_init + 3, 3 * d), device=DEV, generator=g)
    ref = _self_ref(qkv, pool_before, pt, rows, list(range(S)), anc, layer, d, n_head)
    for max_ctx in (None, 0):  # the exact longest context, and "unknown" (largest staging)
        pool.copy_(pool_before)
        out = _run_self(rows, qkv, pool, pt, [s | flag for s in range(S)], anc, layer, d, n_head, n_ctx, n_layer, max_ctx)
        r = _rel(out, ref)
        assert r < 5e-3, f"prefill rel-L2 {r}"


# ---------------------------------------------------------------------------------------------------- sample_topk
def _state_from_history(lay, sample_begin, sampled, G, greedy, without_ts, suppress_blank, max_initial_ts):
    """what beam_update_kernel leaves in ReqState / SeqState after `sampled` tokens"""
    tb = lay.timestamp_begin
    last = sampled[-1] if sampled else lay.transcribe
    prev = sampled[-2] if len(sampled) >= 2 else -1
    stamps = [t for t in sampled if t >= tb]
    return [G, int(greedy), sample_begin + len(sampled), sample_begin, int(without_ts), int(suppress_blank),
            -1 if max_initial_ts is None else max_initial_ts, last, prev, stamps[-1] if stamps else -1]


@pytest.mark.parametrize("name", ["test-tiny.en", "test-tiny", "test-v3"])
def test_sample_topk_matches_oracle_filters_and_torch_topk(name):
    b = B200WhisperBackend(model_spec(name, seed=7), "cuda:0", "bfloat16", max_segments=2, max_sequences=8)  # own engine
    eng = b.engine
    V = eng.dims.n_vocab
    lay = layout_for_vocab(V)
    tb = lay.timestamp_begin
    rng = np.random.default_rng(V)
    sample_begin = 4
    histories = [
        [],                                              # first sampled position: text masked, initial-timestamp ceiling, blanks
        [tb + 5],                                        # one timestamp
        [tb + 5, 300],                                   # timestamp then text
        [tb + 5, 300, 301, tb + 40],                     # text then timestamp: must be followed by a timestamp or EOT
        [tb + 5, 300, tb + 40, tb + 40],                 # pair: timestamps masked entirely
        [tb + 5, 300, tb + 40, tb + 40, 17, 18, 19],     # after a pair: timestamps >= last allowed
        [300, 301],                                      # no timestamp so far (without_timestamps path too)
        [tb + 1500],                                     # last timestamp token of the vocabulary
    ]
    cases = []
    for hist in histories:
        for G, greedy in ((1, True), (1, False), (5, False), (8, False)):
            for without_ts in (False, True):
                for scale in (1.0, 6.0):
                    cases.append((hist, G, greedy, without_ts, scale))
    n = len(cases)
    logits = np.empty((n, V), np.float32)
    state = np.empty((n, 10), np.int32)
    for i, (hist, G, greedy, without_ts, scale) in enumerate(cases):
        x = rng.standard_normal(V).astype(np.float32) * scale
        if i % 3 == 0:
            x[tb:] += 3.0 * scale  # push probability mass onto the timestamps: "sum of timestamp probs > max text prob" rule
        if i % 5 == 0:
            x[rng.integers(0, V, 4)] = x.max() + 1.0  # exact ties for the top: lowest id must win (torch.topk order is checked by value)
        logits[i] = x
        state[i] = _state_from_history(lay, sample_begin, hist, G, greedy, without_ts, i % 2 == 0, 50 if i % 4 else None)
    cand_tok = np.full((n, 9), -7, np.int32)
    cand_lp = np.full((n, 9), np.nan, np.float32)
    L.check(eng.lib.bw_test_sample_topk(eng.handle, logits.ctypes.data_as(L.c_f32_p), n, state.ctypes.data_as(L.c_i32_p),
                                        cand_tok.ctypes.data_as(L.c_i32_p), cand_lp.ctypes.data_as(L.c_f32_p)), "bw_test_sample_topk")
    checked_masks = set()
    for i, (hist, G, greedy, without_ts, scale) in enumerate(cases):
        opts = wo.DecodingOptions(without_timestamps=without_ts, suppress_blank=bool(state[i][5]),
                                  max_initial_timestamp=1.0 if state[i][6] >= 0 else None)
        filt = wo._Filters(lay, sample_begin, opts, 1500)
        x = torch.from_numpy(logits[i : i + 1].copy())
        tokens = torch.tensor([[lay.sot, lay.sot + 1, lay.transcribe, 0][:sample_begin] + hist])
        assert tokens.shape[1] == sample_begin + len(hist)
        filt.apply(x, tokens)
        lp = torch.log_softmax(x[0].float(), -1)
        K = 1 if greedy else G + 1
        n_allowed = int(torch.isfinite(lp).sum())
        k_eff = min(K, n_allowed)
        want_v, _ = lp.topk(k_eff)
        got_t, got_v = cand_tok[i, :K], cand_lp[i, :K]
        assert (got_t[:k_eff] >= 0).all() and len(set(got_t[:k_eff].tolist())) == k_eff, f"case {i}: duplicate / missing candidates"
        assert torch.isfinite(lp[got_t[:k_eff].tolist()]).all(), f"case {i}: the kernel picked a token the filters forbid"
        np.testing.assert_allclose(got_v[:k_eff], want_v.numpy(), rtol=0, atol=2e-4 * max(1.0, scale), err_msg=f"case {i}")
        np.testing.assert_allclose(lp[got_t[:k_eff].tolist()].numpy(), got_v[:k_eff], rtol=0, atol=2e-4 * max(1.0, scale))
        # ties: candidates of equal value come out in ascending id order
        for a in range(k_eff - 1):
            if got_v[a] == got_v[a + 1]:
                assert got_t[a] < got_t[a + 1]
        assert (cand_tok[i, K:] == -7).all(), "wrote past the request's candidate count"
        checked_masks.add((bool((lp[:tb] == -np.inf).all()), bool((lp[tb:] == -np.inf).all())))
    assert len(checked_masks) >= 3, "the random cases did not reach the text-masked, timestamp-masked and mixed branches"
