"""Temperature sampling in the oracle (CPU): the counter-based Gumbel-max draw used for device parity
(`gumbel_noise`, our definition -- see oracle/whisper_oracle.py) is pinned by known answers and shown to be the same
distribution as upstream's `Categorical(logits / T).sample()`, which the oracle uses when no seed is given."""
import numpy as np
import torch

from oracle import whisper_oracle as wo


def test_gumbel_noise_known_answers():
    g = wo.gumbel_noise(123, 0, 3, 8)
    # splitmix64 finaliser over (seed + golden * (stream + 1)) ^ (pos << 32 | id); 23-bit uniforms
    z = []
    for i in range(8):
        x = (123 + 0x9E3779B97F4A7C15) & (2**64 - 1)
        x ^= (3 << 32) | i
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        x ^= x >> 31
        z.append(((x >> 41) + 0.5) / 8388608.0)
    assert np.allclose(g, -np.log(-np.log(np.array(z))), rtol=0, atol=1e-12)
    assert len(set(np.round(g, 9))) == 8
    # distinct streams / positions / seeds give unrelated noise
    for other in (wo.gumbel_noise(123, 1, 3, 8), wo.gumbel_noise(123, 0, 4, 8), wo.gumbel_noise(124, 0, 3, 8)):
        assert not np.allclose(other, g)
    u = np.exp(-np.exp(-wo.gumbel_noise(5, 0, 0, 200000)))
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert wo.window_seed(1, 0, 0) != wo.window_seed(1, 0, 1) != wo.window_seed(1, 100, 0)


def _chi2(counts, p, n):
    exp = n * p
    keep = exp >= 5
    obs_b = np.append(counts[keep], counts[~keep].sum())
    exp_b = np.append(exp[keep], max(exp[~keep].sum(), 1e-12))
    return float(((obs_b - exp_b) ** 2 / exp_b).sum()), int(keep.sum())


def test_seeded_gumbel_max_is_categorical():
    torch.manual_seed(0)
    V, eot, temp, n = 40, 39, 0.8, 4000
    logits = torch.randn(1, V) * 2.0
    logits[0, 5] = -np.inf  # a filtered token is never drawn
    p = torch.softmax(logits[0].double() / temp, -1).numpy()
    logprobs = torch.log_softmax(logits[0], -1)
    tokens = torch.tensor([[1, 2, 3]])
    for seeded in (True, False):
        counts = np.zeros(V)
        for i in range(n):
            dec = wo._Greedy(eot, temp, sample_seed=i if seeded else None)
            s = torch.zeros(1)
            new, done = dec.update(tokens, logits.clone(), s, {})
            tok = int(new[0, -1])
            counts[tok] += 1
            assert abs(float(s[0]) - float(logprobs[tok])) < 1e-6  # un-tempered log-probability is accumulated
            assert done == (tok == eot)
        assert counts[5] == 0
        chi2, dof = _chi2(counts, p, n)
        assert chi2 < dof + 5 * np.sqrt(2 * dof), (seeded, chi2, dof)


def test_finished_hypotheses_keep_emitting_eot():
    eot = 9
    dec = wo._Greedy(eot, 1.0, sample_seed=3)
    tokens = torch.tensor([[1, eot], [1, 2]])
    s = torch.tensor([-1.0, -2.0])
    new, done = dec.update(tokens, torch.zeros(2, 10), s, {})
    assert int(new[0, -1]) == eot and float(s[0]) == -1.0 and float(s[1]) < -2.0
