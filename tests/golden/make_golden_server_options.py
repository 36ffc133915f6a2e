"""Regenerates tests/golden/server_options.json from the REAL reference (run in the build container, where /root/reference
exists: python tests/golden/make_golden_server_options.py): the two decode profiles the server ships
(config/model.yaml:42-65), its built-in default profile and every option key a client may send
(stt_server/config/default/model.py:17-65).  These are the option dicts that actually reach `ModelBackend.transcribe`."""
import json
import logging
import os
import sys
import types

import yaml

REF = "/root/reference"
sys.path.insert(0, REF)
from stt_server.config.default import model as m  # noqa: E402

# the REAL wrapper class (stt_server/model/backends/torch_whisper.py) imports once `whisper` / `faster_whisper` exist as modules;
# its option normalisation and result mapping do not touch them
sys.modules.setdefault("whisper", types.ModuleType("whisper"))
fw = types.ModuleType("faster_whisper"); fw.WhisperModel = object
fwt = types.ModuleType("faster_whisper.transcribe"); fwt.BatchedInferencePipeline = object
sys.modules.setdefault("faster_whisper", fw); sys.modules.setdefault("faster_whisper.transcribe", fwt)
from stt_server.model.backends.torch_whisper import TorchWhisperBackend  # noqa: E402

cfg = yaml.safe_load(open(os.path.join(REF, "config", "model.yaml")))
profiles = cfg["decode_profiles"]
OPTION_CASES = [
    {}, dict(profiles["realtime"], task="transcribe"), dict(profiles["accurate"], task="translate", language="ko"),
    {"log_prob_threshold": -0.5, "logprob_threshold": -0.7}, {"without_timestamps": True, "word_timestamps": True},
    {"without_timestamps": False}, {"temperature": [0.0, 0.2, 0.4], "best_of": 3, "beam_size": 2, "patience": 2.0},
    {"vad_filter": True, "hotwords": "x", "suppress_tokens": [-1], "initial_prompt": "hi", "fp16": False, "prompt": [1, 2]},
    {"condition_on_previous_text": False, "no_speech_threshold": None, "compression_ratio_threshold": 1.5, "length_penalty": None},
]
RESULT_CASES = [
    {"segments": [{"start": 0.0, "end": 1.5, "text": " hello"}, {"start": "2.5", "end": None, "text": None}, "junk",
                  {"start": "x", "end": 3, "text": 5}], "language": "en"},
    {"segments": [], "language": None}, {"language": 7}, {"segments": [{"id": 0}], "language": "ko", "text": "ignored"},
]


class _Model:
    def __init__(self, result):
        self.result, self.calls = result, []

    def transcribe(self, audio, **opts):
        self.calls.append(opts)
        return self.result


logging.getLogger("stt_server.model_backend").setLevel(logging.ERROR)
wrapper = object.__new__(TorchWhisperBackend)
wrapper.device, wrapper.compute_type = "cuda", "float16"
normalised = [TorchWhisperBackend._normalize_options(wrapper, dict(o)) for o in OPTION_CASES]
mapped = []
for res in RESULT_CASES:
    wrapper.model = _Model(res)
    segs, info = TorchWhisperBackend.transcribe(wrapper, None, {"beam_size": 1})
    mapped.append({"segments": [[s.start, s.end, s.text] for s in segs], "language": info.language,
                   "language_probability": info.language_probability, "passed_to_upstream": wrapper.model.calls[0]})
out = {
    "torch_whisper_normalize_options": {"inputs": OPTION_CASES, "outputs": normalised},
    "torch_whisper_result_mapping": {"inputs": RESULT_CASES, "outputs": mapped},
    "decode_profiles": cfg["decode_profiles"],
    "default_decode_profile": dict(m.DEFAULT_DECODE_PROFILE),
    "allowed_decode_option_keys": sorted(m.ALLOWED_DECODE_OPTION_KEYS),
    "default_task": m.DEFAULT_TASK,
    "source": ["config/model.yaml decode_profiles", "stt_server/config/default/model.py",
               "stt_server/model/backends/torch_whisper.py (_normalize_options, transcribe result mapping, run here)"],
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "server_options.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path)
