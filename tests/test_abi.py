"""CPU: the C-ABI library builds, loads without a GPU, and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    syms = []
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            syms += re.findall(r"RL_API\s+[\w\s\*]+?\b(rl_\w+)\s*\(", text)
    return syms


def test_library_exports_every_declared_symbol():
    from realise_b200 import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert lib.rl_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    from realise_b200 import _lib
    lib = _lib.lib()
    d = _lib.GemmDesc()
    rc = lib.rl_gemm_bf16(ctypes.byref(d), None)
    assert rc < 0
    assert b"rl_gemm_bf16" in lib.rl_last_error()
    assert lib.rl_attention_fwd(None, None, None, 1, 1, 1, 64, None) < 0
    assert lib.rl_layernorm_fwd(None, None, None, None, None, 1, 768, ctypes.c_float(1e-12), None) < 0


def test_gemm_desc_layout_matches_header():
    """ctypes mirror vs the C struct: field order / count taken from the header text."""
    from realise_b200 import _lib
    text = open(os.path.join(ROOT, "include", "realise_b200.h")).read()
    body = text.split("typedef struct rl_gemm_desc {", 1)[1].split("} rl_gemm_desc;", 1)[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(\[\d+\])?\s*$", part.strip())
            names.append(m.group(1))
    assert names == [f[0] for f in _lib.GemmDesc._fields_]
