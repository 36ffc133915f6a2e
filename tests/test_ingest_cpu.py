"""Audio ingest (reference stt_server/utils/audio.py:6-30) on the CPU: the oracle restatement against golden vectors
produced by the REAL reference functions (tests/golden/make_golden_ingest.py), and the product's host-side filter
bank against the oracle's."""
import os

import numpy as np
import pytest

from oracle import audio_ingest as ai

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest.npz"))
CASES = sorted((int(k.split("_")[1]), int(k.split("_")[2])) for k in GOLD.files if k.startswith("pcm_"))


@pytest.mark.parametrize("rate,n", CASES)
def test_oracle_matches_reference_golden(rate, n):
    pcm = GOLD[f"pcm_{rate}_{n}"]
    f = ai.pcm16_to_float32(pcm.tobytes())
    assert f.dtype == np.float32 and np.array_equal(f, GOLD[f"f32_{rate}_{n}"])  # exact: int16 / 2^15
    y = ai.ensure_16k(f, rate)
    ref = GOLD[f"y16k_{rate}_{n}"]
    assert y.shape == ref.shape and y.dtype == np.float32
    # same taps, different summation order (torch conv1d vs numpy matmul): a few ulps of a unit-scale signal
    assert float(np.abs(y - ref).max()) <= 1e-6
    if rate == 16000:
        assert y is f  # identity, no copy (utils/audio.py:13-14)


def test_resampled_length_and_taps_match_oracle():
    from b200_whisper.ingest import resample_taps, resampled_length

    for rate in (8000, 11025, 12000, 22050, 24000, 32000, 44100, 48000, 96000):
        orig, new, width, taps = resample_taps(rate)
        k, w, o, nw = ai.sinc_resample_kernel(rate, 16000)
        assert (orig, new, width) == (o, nw, w)
        assert taps.dtype == np.float32 and taps.shape == (new, 2 * width + orig) and np.array_equal(taps, k)
        for n in (1, 2, 159, 160, 4410, 48000):
            assert resampled_length(n, rate) == ai.ensure_16k(np.zeros(n, np.float32), rate).shape[0]
    with pytest.raises(ValueError):
        resample_taps(0)
    with pytest.raises(ValueError):
        resample_taps(44100.5)


def test_ingest_edge_cases():
    assert ai.pcm16_to_float32(b"").shape == (0,)
    full = np.array([-32768, 32767, 0, 1, -1], dtype=np.int16)
    f = ai.pcm16_to_float32(full.tobytes())
    assert f[0] == -1.0 and f[1] == np.float32(32767 / 32768) and f[2] == 0.0
    # DC gain of every polyphase filter is ~1 (rolloff 0.99 keeps 0 Hz): a constant stays a constant away from the edges
    y = ai.ensure_16k(np.full(4800, 0.25, np.float32), 48000)
    assert abs(float(y[200:-200].mean()) - 0.25) < 1e-3
