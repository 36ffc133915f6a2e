"""CPU oracle for the Whisper hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`whisper-streaming-stt-server_b200/`, imported as `b200_whisper`) never does and fails loudly
when its CUDA library is missing.

PARITY UNPINNED: the arithmetic the reference runs lives in the third-party
package `openai-whisper==20250625` (reference `requirements-lock.txt:48`, call
sites `stt_server/model/backends/torch_whisper.py:21,55`), which is neither
under `/root/reference` nor installable here.  This package restates its
published algorithm in plain torch fp32; the reference's own tests hold no
numeric vectors for the path.  Independent cross-checks
against HF `transformers` (same published model, different code) are in
`tests/test_oracle_crosscheck.py`.
"""
