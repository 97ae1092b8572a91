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
    i64, i32, f32, u64, u32 = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint32
    assert lib.rl_attention_fwd(None, None, None, None, i64(1), i64(1), i64(1), i64(64), i32(0), f32(0), u64(0), u32(0), None,
                                None) < 0
    assert lib.rl_attention_bwd(None, None, None, None, None, None, None, i64(1), i64(1), i64(1), i64(64), i32(2), f32(0), u64(0),
                                u32(0), None, None) < 0
    assert lib.rl_layernorm_fwd(None, None, None, None, None, i64(1), i64(768), f32(1e-12), f32(0), u64(0), u32(0), None, i32(0),
                                i32(0), None) < 0


def test_no_process_wide_switches_in_the_abi():
    """re-entrant, no mutable global state — formats, the dropout counter and tuning knobs are per call."""
    syms = declared_symbols()
    assert not [s for s in syms if s.startswith("rl_set_") or "_set_" in s], syms
    src = "".join(open(os.path.join(ROOT, "realise_b200", "csrc", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "realise_b200", "csrc")))
    assert "rl_set_half_format" not in src and "rl_set_dropout_seed_ptr" not in src and "set_debug_mode" not in src


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


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Implicit-im2col GEMM operand, format / seed switches and the fused passes reject bad arguments on the host."""
    from realise_b200 import _lib
    lib = _lib.lib()
    d = _lib.GemmDesc()
    d.a, d.b, d.out = 16, 16, 16           # non-null, never dereferenced: validation fails first
    d.M, d.N, d.K = 64, 576, 1 << 20
    d.lda, d.ldb, d.ldo = 64, 64, 576
    d.b_mode, d.b_major, d.a_major = 1, 0, 1
    assert lib.rl_gemm_bf16(ctypes.byref(d), None) < 0
    assert b"b_mode" in lib.rl_last_error()
    d.b_mode, d.b_major, d.a_major, d.a_dtype = 0, 0, 0, 1        # f32 is not an MMA operand format
    assert lib.rl_gemm_bf16(ctypes.byref(d), None) < 0
    assert b"a_dtype" in lib.rl_last_error()
    assert lib.rl_gelu_fwd(None, None, ctypes.c_int64(8), ctypes.c_int32(0), None) < 0
    assert lib.rl_gelu_bwd_colsum(None, None, None, ctypes.c_int64(8), ctypes.c_int64(8), ctypes.c_int64(8), ctypes.c_int32(0),
                                  None) < 0
    assert lib.rl_split3_bf16(None, None, ctypes.c_int64(1), ctypes.c_int64(8), ctypes.c_int32(0), None) < 0
    assert lib.rl_workspace_bytes(b"gate_fuse_bwd", ctypes.c_int64(2), ctypes.c_int64(16), ctypes.c_int64(768)) == (2 * 16 * 3 + 2 * 2 * 768) * 4
    assert lib.rl_workspace_bytes(b"nope", ctypes.c_int64(1), ctypes.c_int64(1), ctypes.c_int64(1)) == -1
    assert lib.rl_mt_adamw_dev(None, None, ctypes.c_int64(1), None, ctypes.c_float(1.0), None, ctypes.c_float(0.9),
                               ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(1.0), None) < 0


def test_graphed_step_and_schedule_helpers_on_cpu():
    """Host logic of the CUDA-graph step that needs no GPU: optimizer type check, device-schedule values."""
    import math

    import pytest
    import torch

    from realise_b200.graphed import GraphedTrainStep
    from realise_b200.optim import FusedAdamW
    lin = torch.nn.Linear(4, 4)
    with pytest.raises(TypeError):
        GraphedTrainStep(lin, torch.optim.SGD(lin.parameters(), lr=0.1))
    opt = FusedAdamW(lin.parameters(), lr=5e-5, betas=(0.9, 0.999))
    lr, bc1, bc2 = opt.hyper_values(3)
    assert lr == 5e-5 and math.isclose(bc1, 1 - 0.9 ** 3) and math.isclose(bc2, 1 - 0.999 ** 3)
    opt.param_groups[0]["lr"] = 1e-5          # LambdaLR writes the group's lr: the next replay must read it
    assert opt.hyper_values(4)[0] == 1e-5
