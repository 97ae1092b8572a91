"""GPU dev check for the tcgen05 GEMM / implicit-GEMM conv (run under gpurun)."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(1.0, scale)
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.4e} ref_max={scale:.3f}", flush=True)
    return ok


def t_plain(M, N, K, bias=False, act=0, res=None, out_dtype=torch.bfloat16, scale=False):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    bi = torch.randn(N, device=dev) if bias else None
    sc = (torch.rand(N, device=dev) + 0.5) if scale else None
    r = None
    if res is not None:
        r = torch.randn(M, N, device=dev).to(res)
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    ops.gemm(a, b, out, bias=bi, act=act, res=r, scale=sc)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    if sc is not None:
        ref = ref * sc
    if bi is not None:
        ref = ref + bi
    if r is not None:
        ref = ref + r.float()
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:
        ref = F.relu(ref)
    tol = 2e-2 if out_dtype == torch.bfloat16 else 2e-3
    return report(f"gemm M{M} N{N} K{K} bias={bias} act={act} res={res} out={out_dtype}", out, ref, tol)


def t_trans(M, N, K, a_t, b_t, accumulate=False):
    """MN-major operands (weight-gradient / data-gradient forms), optional in-place f32 accumulation."""
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    a = A.t().contiguous() if a_t else A
    b = B.t().contiguous() if b_t else B
    out = torch.randn(M, N, device=dev) if accumulate else torch.empty(M, N, device=dev)
    ref = A.float() @ B.float().t() + (out if accumulate else 0)
    ops.gemm(a, b, out, a_t=a_t, b_t=b_t, res=out if accumulate else None)
    torch.cuda.synchronize()
    return report(f"gemm-T M{M} N{N} K{K} a_t={a_t} b_t={b_t} acc={accumulate}", out, ref, 2e-3)


def t_conv(nimg, Cin, Cout, S, stride):
    """3x3 pad-1 conv (stride 1 or 2) as implicit GEMM vs F.conv2d.  S = OUTPUT map size."""
    if stride == 1:
        x = (torch.randn(nimg, Cin, S, S, device=dev)).bfloat16()
        xin = x.permute(0, 2, 3, 1).contiguous().view(nimg, 1, S, S, Cin)
        taps = [(kw - 1, kh - 1, 0) for kh in range(3) for kw in range(3)]
        planes = 1
    else:
        Si = 2 * S
        x = (torch.randn(nimg, Cin, Si, Si, device=dev)).bfloat16()
        nhwc = x.permute(0, 2, 3, 1).contiguous()  # [n, Si, Si, C]
        # parity split: [n, ph, pw, S, S, C]
        xin = nhwc.view(nimg, S, 2, S, 2, Cin).permute(0, 2, 4, 1, 3, 5).contiguous().view(nimg, 4, S, S, Cin)
        taps = []
        for kh in range(3):
            for kw in range(3):
                ph, dh = (0, 0) if kh == 1 else (1, -1 if kh == 0 else 0)
                pw, dw = (0, 0) if kw == 1 else (1, -1 if kw == 0 else 0)
                taps.append((dw, dh, ph * 2 + pw))
        planes = 4
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) * 0.1).bfloat16()
    wk = w.permute(0, 2, 3, 1).contiguous().view(Cout, 9 * Cin)  # tap-major (kh, kw, c)
    out = torch.empty(nimg * S * S, Cout, device=dev, dtype=torch.float32)
    ops.conv_gemm(xin, wk, out, nimg=nimg, H=S, W=S, planes=planes, taps=taps)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=1)  # [n, Cout, S, S]
    ref = ref.permute(0, 2, 3, 1).reshape(nimg * S * S, Cout)
    return report(f"conv nimg{nimg} {Cin}->{Cout} out{S}x{S} stride{stride}", out, ref, 5e-3)


def perf(M, N, K, iters=20, bias=False, act=0, res=None, out_dtype=torch.bfloat16):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    bi = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev).to(res) if res is not None else None
    for _ in range(3):
        ops.gemm(a, b, out, bias=bi, act=act, res=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, out, bias=bi, act=act, res=r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    out_b = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b.t(), out=out_b)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"[perf] M{M} N{N} K{K} bias={bias} act={act} res={res} out={out_dtype}: ours {ms*1e3:.1f} us {tf:.0f} TF/s | cublas {ms2*1e3:.1f} us "
          f"{2.0*M*N*K/ms2/1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= t_plain(128, 128, 64)
    ok &= t_plain(128, 256, 64)
    ok &= t_plain(256, 512, 256)
    ok &= t_plain(1000, 768, 768, bias=True)
    ok &= t_plain(8192, 3072, 768, bias=True, act=1)
    ok &= t_plain(4096, 768, 3072, bias=True, res=torch.float32, out_dtype=torch.float32)
    ok &= t_plain(300, 1000, 128, bias=True, out_dtype=torch.float32)
    ok &= t_plain(512, 21128, 768, bias=True, out_dtype=torch.float32)
    ok &= t_plain(512, 328, 192, scale=True, bias=True, act=2, res=torch.bfloat16)
    ok &= t_plain(1000, 64, 576, scale=True, bias=True, act=2, res=torch.bfloat16)
    ok &= t_plain(777, 200, 128, bias=True, res=torch.float32)
    ok &= t_plain(640, 331, 64, bias=True, out_dtype=torch.float32)
    ok &= t_plain(333, 776, 128, bias=True, act=1, res=torch.bfloat16)
    for args in [(128, 64, 64, True, False), (128, 64, 64, False, True), (128, 256, 128, True, True),
                 (768, 768, 8192, True, True), (3072, 768, 4096, True, True), (8192, 768, 3072, False, True),
                 (1000, 21128, 768, False, True), (21128, 768, 1024, True, True, True),
                 (768, 2304, 32, True, True)]:
        try:
            ok &= t_trans(*args)
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[FAIL] gemm-T {args}: {e}", flush=True)
    for args in [(4, 64, 64, 16, 1), (8, 128, 128, 8, 1), (16, 256, 256, 4, 1), (64, 512, 512, 2, 1),
                 (8, 64, 128, 8, 2), (16, 128, 256, 4, 2), (64, 256, 512, 2, 2), (200, 512, 768, 1, 2)]:
        try:
            ok &= t_conv(*args)
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[FAIL] conv {args}: {e}", flush=True)
    if "--pairoff" in sys.argv:
        from realise_b200._lib import lib
        lib().rl_gemm_set_pair_mode(0)
    if "--quick" not in sys.argv:
        for shp in [(8192, 2304, 768), (8192, 768, 768), (8192, 3072, 768), (8192, 768, 3072), (8192, 21128, 768),
                    (16384, 3072, 768)]:
            perf(*shp)
    perf(8192, 2304, 768, bias=True)
    perf(8192, 3072, 768, bias=True, act=1)
    perf(8192, 768, 768, bias=True, res=torch.float32, out_dtype=torch.float32)
    perf(8192, 768, 3072, bias=True, res=torch.float32, out_dtype=torch.float32)
    perf(8192, 21128, 768, bias=True, out_dtype=torch.float32)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
