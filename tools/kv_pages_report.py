"""Self-KV memory actually held by concurrent beam-search decodes (paged pool + prefix sharing) vs the static
reservation of round 1 (448 positions per hypothesis).  N windows x beam G decode `steps` tokens through the real
scheduler; the engine's page statistics give the peak.

Usage: python tools/kv_pages_report.py [model] [windows] [beam] [steps]"""
import json
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200_whisper.backend import B200WhisperBackend  # noqa: E402
from b200_whisper.synth import synth_audio  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
G = int(sys.argv[3]) if len(sys.argv) > 3 else 5
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 97  # 3 initial tokens + 97 = context 100

b = B200WhisperBackend(f"random:{model}:0:0.1", "cuda:0", "bfloat16", max_segments=N, max_sequences=N * G, max_encoder_batch=16)
eng = b.engine
v = b.vocab
initial = v.sot_sequence("en", "transcribe")
audios = [synth_audio(900 + i, 3.0 + (i % 7)) for i in range(N)]
calls = [eng.open_call(a) for a in audios]
out = [None] * N


def work(i):
    out[i] = calls[i].decode(0, initial, 0, G if G > 1 else None, 1.0, 1.0, sample_len=steps)


th = [threading.Thread(target=work, args=(i,)) for i in range(N)]
[t.start() for t in th]
[t.join() for t in th]
[c.close() for c in calls]
st = eng.stats()
dims = eng.dims
per_token = dims.n_text_layer * 2 * dims.n_text_state * 2
rep = {"model": model, "windows": N, "beam": G, "context": len(initial) + steps,
       "page_tokens": 16, "page_bytes": st["kv_page_bytes"], "pool_pages": st["kv_pages_total"],
       "pool_gb": st["kv_pages_total"] * st["kv_page_bytes"] / 1e9,
       "peak_pages_in_use": st["kv_pages_peak"], "peak_gb_in_use": st["kv_pages_peak"] * st["kv_page_bytes"] / 1e9,
       "pages_in_use_after": st["kv_pages_in_use"],
       "worst_case_reserved_gb": N * G * ((len(initial) + steps + 15) // 16) * st["kv_page_bytes"] / 1e9,
       "round1_static_gb": N * G * 448 * per_token / 1e9,
       "tokens_if_no_sharing_gb": N * G * (len(initial) + steps) * per_token / 1e9,
       "all_decoded": all(o and o["n_steps"] >= 1 for o in out)}
print(json.dumps(rep))
