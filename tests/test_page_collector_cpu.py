"""Host logic of the paged self-KV pool, without a GPU: the scheduler's page bookkeeping for one beam-search request is
replayed through the C ABI (`bw_test_page_collector`, no CUDA call) and checked against an exact token-level model of what
beam reordering does (upstream BeamSearchDecoder.update -> rearrange_kv_cache; here: ancestry rows instead of K/V copies).

Invariants: (safety) every (beam slot, block) that a surviving hypothesis' ancestry references still holds its page;
(tightness) a COMPLETE block holds a page for exactly the slots some survivor references; nothing leaks at the end."""
import ctypes as C

import numpy as np
import pytest

from b200_whisper import _lib as L

P, NB = 16, 28


def replay(G, n_init, parents):
    lib = L.load()
    n_steps = len(parents)
    par = np.ascontiguousarray(np.array(parents, dtype=np.uint8).reshape(n_steps, G))
    masks = np.zeros((n_steps, NB), dtype=np.uint8)
    used = np.zeros(n_steps, dtype=np.int32)
    L.check(lib.bw_test_page_collector(G, n_init, n_steps, par.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       masks.ctypes.data_as(C.POINTER(C.c_uint8)), used.ctypes.data_as(L.c_i32_p)), "bw_test_page_collector")
    return masks, used


def model(G, n_init, parents):
    """exact ancestry: anc[j][t] = beam slot whose page holds position t of hypothesis j"""
    anc = [[0] * n_init for _ in range(G)]
    out = []
    for k, par in enumerate(parents):
        written = n_init + k  # positions [0, written) hold K/V after this step
        if k > 0:
            for j in range(G):
                assert anc[j][written - 1] == j  # slot j wrote the position it had been assigned
        anc = [list(anc[par[j]]) + [j] for j in range(G)]  # reorder; position `written` will be written by slot j
        ref = {}
        for j in range(G):
            for t in range(written):
                ref.setdefault(t // P, set()).add(anc[j][t])
        out.append((ref, written))
    return out


@pytest.mark.parametrize("G,n_init,n_steps,style", [(5, 3, 224, "merge"), (8, 4, 200, "random"), (5, 3, 120, "identity"), (2, 40, 60, "random"),
                                                    (1, 3, 100, "identity"), (5, 223, 224, "merge"), (3, 17, 50, "collapse")])
def test_page_collector_matches_exact_ancestry(G, n_init, n_steps, style):
    rng = np.random.default_rng(G * 1000 + n_init + n_steps)
    parents = []
    for k in range(n_steps):
        if k == 0:
            par = [0] * G  # every hypothesis starts from the prompt in slot 0
        elif style == "identity":
            par = list(range(G))  # greedy / independent samples
        elif style == "collapse":
            par = [int(rng.integers(0, G))] * G  # every survivor descends from one hypothesis
        elif style == "merge":
            par = sorted(int(x) for x in rng.choice(G, size=G, p=np.array([0.5] + [0.5 / (G - 1)] * (G - 1))))  # trained-model-like
        else:
            par = [int(x) for x in rng.integers(0, G, G)]
        parents.append(par)
    masks, used = replay(G, n_init, parents)
    exact = model(G, n_init, parents)
    for k, (ref, written) in enumerate(exact):
        cur_block = written // P  # block of the position the next step writes
        for b in range(NB):
            have = {j for j in range(G) if masks[k, b] >> j & 1}
            need = ref.get(b, set())
            assert need <= have, f"step {k} block {b}: pages of referenced slots {sorted(need - have)} were freed"
            if b < cur_block and G > 1:
                assert have == need, f"step {k} block {b}: slots {sorted(have - need)} keep a page nobody references"
        assert used[k] == int(sum(bin(int(m)).count("1") for m in masks[k]))
    # beam search holds far fewer pages than hypotheses x blocks once the beams share ancestors
    if style in ("merge", "collapse") and G > 1:
        worst = G * (-(-(n_init + n_steps) // P))
        assert used.max() < worst, (used.max(), worst)


def test_page_collector_rejects_bad_geometry():
    lib = L.load()
    par = np.zeros((4, 9), np.uint8)
    masks = np.zeros((4, NB), np.uint8)
    used = np.zeros(4, np.int32)
    st = lib.bw_test_page_collector(9, 3, 4, par.ctypes.data_as(C.POINTER(C.c_uint8)), masks.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    used.ctypes.data_as(L.c_i32_p))
    assert st != 0
