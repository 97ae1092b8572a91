"""GEMM micro-benchmark at the train-step shapes: generic epilogue (tune_no_pair=5) vs the cost model (=0: compile-time
specialised epilogues where one matches); the two must agree bit for bit.
Checks each result against torch, times with CUDA events (L2 flushed between launches)."""
import sys
import torch
sys.path.insert(0, ".")
from realise_b200 import ops

torch.manual_seed(0)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = [  # name, (M, N, K), a_t, b_t, split_k, out dtype, residual
    ("QKV fwd", (M, 2304, 768), False, False, 0, torch.bfloat16, False),
    ("out-proj fwd (+bias+res f32)", (M, 768, 768), False, False, 0, torch.float32, True),
    ("out-proj fwd (+bias+DROP+res f32)", (M, 768, 768), False, False, 0, torch.float32, True),
    ("dgrad dctx = dy1 Wo (b_t)", (M, 768, 768), False, True, 0, torch.bfloat16, False),
    ("FFN1 fwd", (M, 3072, 768), False, False, 0, torch.bfloat16, False),
    ("FFN2 fwd (+res f32)", (M, 768, 3072), False, False, 0, torch.float32, True),
    ("dgrad du = dy2 W2 (b_t)", (M, 3072, 768), False, True, 0, torch.bfloat16, False),
    ("dgrad dx1 = du W1 (b_t, +res)", (M, 768, 3072), False, True, 0, torch.float32, True),
    ("wgrad dW1 = du^T x1 (split-K)", (3072, 768, M), True, True, -1, torch.float32, False),
    ("wgrad dWqkv (split-K)", (2304, 768, M), True, True, -1, torch.float32, False),
    ("classifier fwd", (M, 21128, 768), False, False, 0, torch.float32, False),
]
for name, (m, n, k), a_t, b_t, sk, odt, res in cases:
    a = (torch.randn(k, m, device=dev) if a_t else torch.randn(m, k, device=dev)).bfloat16()
    b = (torch.randn(k, n, device=dev) if b_t else torch.randn(n, k, device=dev)).bfloat16()
    bias = torch.randn(n, device=dev) if sk == 0 else None
    r = torch.randn(m, n, device=dev) if res else None
    A = a.float().t() if a_t else a.float()
    Bm = b.float() if b_t else b.float().t()
    ref = A @ Bm
    if bias is not None:
        ref = ref + bias
    if r is not None:
        ref = ref + r
    row = [name, f"{m}x{n}x{k}"]
    drop = (0.1, 1234, 7) if "DROP" in name else None
    outs = []
    for mode in (5, 0):
        ops.TUNE_NO_PAIR = mode
        out = torch.zeros(m, n, device=dev, dtype=odt)

        def run():
            if sk:
                out.zero_()
            ops.gemm(a, b, out, bias=bias, res=r, a_t=a_t, b_t=b_t, split_k=sk, drop=drop)
        run()
        torch.cuda.synchronize()
        outs.append(out.clone())
        err = float((out.float() - ref).abs().max() / ref.abs().max()) if drop is None else float("nan")
        if sk:
            def run():  # noqa: F811 — time without the memset
                ops.gemm(a, b, out, a_t=a_t, b_t=b_t, split_k=sk)
        us = timeit(run)
        row.append(f"mode{mode}: {us:7.1f} us {2.0 * m * n * k / us / 1e6:7.1f} TF/s relerr {err:.1e}")
    row.append("bit-identical" if (sk or torch.equal(outs[0], outs[1])) else "MODES DIFFER")
    print(" | ".join(row), flush=True)
ops.TUNE_NO_PAIR = 0
# ---- fused GELU epilogues vs GEMM + element-wise pass ----
m, n, k = M, 3072, 768
a = torch.randn(m, k, device=dev).bfloat16(); w = (torch.randn(n, k, device=dev) * 0.05).bfloat16(); bias = torch.randn(n, device=dev) * 0.1
u = torch.empty(m, n, device=dev, dtype=torch.bfloat16); h = torch.empty_like(u); h2 = torch.empty_like(u); u2 = torch.empty_like(u)
t_plain = timeit(lambda: ops.gemm(a, w, u, bias=bias)); t_gelu = timeit(lambda: ops.gelu(u, h))
t_fused = timeit(lambda: ops.gemm(a, w, h2, bias=bias, act=ops.ACT_GELU_SAVE, out2=u2))
ops.TUNE_NO_PAIR = 5
h3 = torch.empty_like(u); u3 = torch.empty_like(u)
t_fused_generic = timeit(lambda: ops.gemm(a, w, h3, bias=bias, act=ops.ACT_GELU_SAVE, out2=u3))
ops.TUNE_NO_PAIR = 0
print(f"GELU_SAVE epilogue: generic {t_fused_generic:.1f} us, specialised {t_fused:.1f} us, identical {torch.equal(h2, h3) and torch.equal(u2, u3)}")
print(f"FFN1: gemm {t_plain:.1f} + gelu pass {t_gelu:.1f} = {t_plain + t_gelu:.1f} us | fused GELU_SAVE epilogue {t_fused:.1f} us | "
      f"max diff h {float((h.float() - h2.float()).abs().max()):.3e} u {float((u.float() - u2.float()).abs().max()):.3e}", flush=True)
dy = (torch.randn(m, k, device=dev) * 0.1).bfloat16(); w2 = (torch.randn(k, n, device=dev) * 0.05).bfloat16()   # W2 [768, 3072]
du = torch.empty(m, n, device=dev, dtype=torch.bfloat16); du2 = torch.empty_like(du); db = torch.zeros(n, device=dev)
t_plain = timeit(lambda: ops.gemm(dy, w2, du, b_t=True)); t_pass = timeit(lambda: ops.gelu_bwd_colsum(du, u, db))
ops.gemm(dy, w2, du, b_t=True); ops.gelu_bwd_colsum(du, u, db)
t_fused = timeit(lambda: ops.gemm(dy, w2, du2, b_t=True, res=u, act=ops.ACT_GELU_GRAD)); t_cs = timeit(lambda: ops.colsum_bf16(du2, db))
db2 = torch.zeros(n, device=dev)
t_fused2 = timeit(lambda: ops.gemm(dy, w2, du2, b_t=True, res=u, act=ops.ACT_GELU_GRAD, colsum=db2))
db2.zero_(); ops.gemm(dy, w2, du2, b_t=True, res=u, act=ops.ACT_GELU_GRAD, colsum=db2); db.zero_(); ops.colsum_bf16(du2, db)
print(f"du: gemm {t_plain:.1f} + gelu_bwd_colsum {t_pass:.1f} = {t_plain + t_pass:.1f} us | fused GELU_GRAD epilogue {t_fused:.1f} + colsum {t_cs:.1f} us | "
      f"fused incl. colsum {t_fused2:.1f} us | max diff {float((du.float() - du2.float()).abs().max()):.3e} colsum rel diff "
      f"{float((db - db2).abs().max() / db.abs().max()):.2e}", flush=True)
