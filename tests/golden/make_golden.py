"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference's arithmetic (openai-whisper 20250625) is not installable here and the reference's own tests hold
no numeric vectors for this path, so these fixtures pin OUR oracle against regressions;
independent agreement with HF transformers' port of the same model is checked in tests/test_oracle_crosscheck.py.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from b200_whisper.synth import MODEL_DIMS, random_state_dict, synth_audio  # noqa: E402
from oracle import whisper_oracle as wo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REALTIME = {"beam_size": 1, "best_of": 1, "patience": 1.0, "temperature": 0.0, "length_penalty": 1.0,
            "without_timestamps": True, "language": "en"}
ACCURATE = dict(REALTIME, beam_size=5, best_of=5)


def main():
    src = open(os.path.join(ROOT, "oracle", "whisper_oracle.py"), "rb").read()
    meta = {"oracle_sha256": hashlib.sha256(src).hexdigest(), "torch": torch.__version__}
    mel = {}
    for n_mels in (80, 128):
        for seed, seconds in ((1, 3.96), (2, 10.0), (3, 0.25)):
            audio = synth_audio(seed, seconds)
            m = wo.log_mel_spectrogram(audio, n_mels, padding=wo.N_SAMPLES).numpy()
            mel[f"mel{n_mels}_s{seed}"] = m[:, ::7].astype(np.float32)  # every 7th frame
            mel[f"mel{n_mels}_s{seed}_shape"] = np.array(m.shape)
    np.savez_compressed(os.path.join(OUT, "mel.npz"), **mel)

    dec = {}
    for name in ("test-tiny", "test-tiny.en", "test-v3"):
        dims = MODEL_DIMS[name]
        model = wo.Whisper(wo.ModelDimensions(**dims.__dict__), random_state_dict(dims, 0, emb_std=0.1))
        audio = synth_audio(1, 4.0)
        melw = wo.pad_or_trim(wo.log_mel_spectrogram(audio, dims.n_mels, padding=wo.N_SAMPLES), 3000)
        xa = model.encode(melw[None])
        dec[f"{name}_enc_sample"] = xa[0, ::97, ::5].numpy().astype(np.float32)
        for pname, prof in (("realtime", REALTIME), ("accurate", ACCURATE)):
            segs, info, raw = wo.backend_transcribe(model, audio, prof)
            w = raw["windows"][0]
            dec[f"{name}_{pname}_tokens"] = np.array(w.tokens, dtype=np.int32)
            dec[f"{name}_{pname}_stats"] = np.array([w.sum_logprob, w.avg_logprob, w.no_speech_prob, w.min_margin], dtype=np.float64)
            dec[f"{name}_{pname}_segments"] = np.array([[s[0], s[1]] for s in segs], dtype=np.float64)
    # temperature > 0: seeded Gumbel-max draws (OUR counter-based generator, restated on the device) and the fallback ladder
    for name in ("test-tiny", "test-v3"):
        dims = MODEL_DIMS[name]
        model = wo.Whisper(wo.ModelDimensions(**dims.__dict__), random_state_dict(dims, 0, emb_std=0.1, eot_bias=4.0))
        audio = synth_audio(51, 6.0)
        for tag, temperature, extra in (("t07", 0.7, {}), ("ladder", (0.0, 0.4, 0.8), {"logprob_threshold": 10.0})):
            opts = dict(ACCURATE, best_of=3, temperature=temperature, **extra)
            segs, info, raw = wo.backend_transcribe(model, audio, opts, sample_seed=11)
            w = raw["windows"][0]
            dec[f"{name}_sample_{tag}_tokens"] = np.array(w.tokens, dtype=np.int32)
            dec[f"{name}_sample_{tag}_stats"] = np.array([w.sum_logprob, w.avg_logprob, w.temperature, w.min_margin], dtype=np.float64)
    dec["gumbel_seed123_stream2_pos7"] = wo.gumbel_noise(123, 2, 7, 64)
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **dec)
    with open(os.path.join(OUT, "META.txt"), "w") as fh:
        for k, v in meta.items():
            fh.write(f"{k}: {v}\n")
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
