"""The C-ABI library loads and exports every symbol include/b200_whisper.h declares; struct layouts seen by
ctypes match the C compiler's; without a GPU the engine refuses loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from b200_whisper import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", "b200_whisper.h"), os.path.join(ROOT, "include", "b200_whisper_hooks.h")]


def declared_symbols(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    boundary, hooks = declared_symbols(HEADERS[0]), declared_symbols(HEADERS[1])
    assert len(boundary) >= 22 and len(hooks) >= 10
    for n in boundary + hooks:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    # the drop-in boundary carries no bench / test / debug entry point
    assert not [n for n in boundary if n.startswith(("bw_bench_", "bw_test_", "bw_debug_"))]
    assert sorted(L.SIGNATURES) == sorted(set(boundary + hooks)), "ctypes signatures without a declaration in include/"
    assert lib.bw_version() >= 100


def test_hooks_header_compiles_as_c(tmp_path):
    prog = tmp_path / "hooks.c"
    prog.write_text('#include "b200_whisper_hooks.h"\nint main(void){return 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(prog), "-o", str(tmp_path / "h.o")],
                   check=True)


def test_struct_layouts_match_c(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b200_whisper.h"\nint main(){'
        'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bw_model_dims), sizeof(bw_engine_config), sizeof(bw_tensor_desc),'
        ' sizeof(bw_token_tables), sizeof(bw_decode_opts), sizeof(bw_result), sizeof(bw_lang_result),'
        ' offsetof(bw_result, sum_logprob), offsetof(bw_decode_opts, patience)); return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(L.ModelDimsC), C.sizeof(L.EngineConfigC), C.sizeof(L.TensorDescC), C.sizeof(L.TokenTablesC),
            C.sizeof(L.DecodeOptsC), C.sizeof(L.ResultC), C.sizeof(L.LangResultC), L.ResultC.sum_logprob.offset,
            L.DecodeOptsC.patience.offset]
    assert got == want


@pytest.mark.skipif(L.load().bw_device_count() > 0, reason="GPU present")
def test_no_gpu_means_loud_failure():
    """no compute calls without a GPU: creation must fail with BW_ERR_NO_DEVICE, the backend must raise"""
    lib = L.load()
    dims = L.ModelDimsC(80, 1500, 128, 2, 2, 51865, 448, 128, 2, 2)
    cfg = L.EngineConfigC(0, 0, 0, 0, 0, 0)
    h = C.c_void_p()
    st = lib.bw_engine_create(C.byref(dims), C.byref(cfg), C.byref(h))
    assert st == -5 and b"CUDA device" in lib.bw_last_error()
    from b200_whisper.backend import B200WhisperBackend

    with pytest.raises(RuntimeError):
        B200WhisperBackend("random:test-tiny", "cuda:0", "bfloat16")
    with pytest.raises(ValueError):
        B200WhisperBackend("random:test-tiny", "cpu", "int8")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "whisper-streaming-stt-server_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    code = "import sys; import b200_whisper.backend, b200_whisper.engine, b200_whisper.register; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported by the product'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
