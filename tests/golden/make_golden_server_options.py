"""Regenerates tests/golden/server_options.json from the REAL reference (run in the build container, where /root/reference
exists: python tests/golden/make_golden_server_options.py): the two decode profiles the server ships
(config/model.yaml:42-65), its built-in default profile and every option key a client may send
(stt_server/config/default/model.py:17-65).  These are the option dicts that actually reach `ModelBackend.transcribe`."""
import json
import os
import sys

import yaml

REF = "/root/reference"
sys.path.insert(0, REF)
from stt_server.config.default import model as m  # noqa: E402

cfg = yaml.safe_load(open(os.path.join(REF, "config", "model.yaml")))
out = {
    "decode_profiles": cfg["decode_profiles"],
    "default_decode_profile": dict(m.DEFAULT_DECODE_PROFILE),
    "allowed_decode_option_keys": sorted(m.ALLOWED_DECODE_OPTION_KEYS),
    "default_task": m.DEFAULT_TASK,
    "source": ["config/model.yaml decode_profiles", "stt_server/config/default/model.py"],
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "server_options.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path)
