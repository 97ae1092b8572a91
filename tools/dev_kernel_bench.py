"""Isolated timings of the train-step kernels at the bench shapes (B=128, L=128): median of 10 launches, L2 flushed
(256 MB memset) before each, CUDA events.  Prints achieved GB/s or TFLOP/s next to each."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402

dev = "cuda"
N, H, I, V = 16384, 768, 3072, 21128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = set(sys.argv[1:])


def timeit(name, fn, nbytes=0, flops=0, iters=10):
    if only and not any(o in name for o in only):
        return
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    extra = ""
    if nbytes:
        extra += f"  {nbytes / us / 1e3:8.1f} GB/s"
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    print(f"{name:44s} {us:9.1f} us{extra}", flush=True)


def f32(*s):
    return torch.randn(*s, device=dev)


def bf(*s):
    return torch.randn(*s, device=dev).bfloat16()


# ---- row-wise ----
dy, x, add, g = f32(N, H), f32(N, H), f32(N, H), f32(H)
dx, dxb = f32(N, H), bf(N, H)
dg, db, ds = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
timeit("ln_bwd (no dropout)", lambda: ops.layernorm_bwd(dy, x, g, None, dx, dxb, dg, db, ds, 1e-12), nbytes=N * H * 14)
timeit("ln_bwd (dropout site_out)", lambda: ops.layernorm_bwd(dy, x, g, None, dx, dxb, dg, db, ds, 1e-12, drop_p=0.1,
                                                               drop_seed=1, site_out=1013), nbytes=N * H * 14)
o32, o16 = f32(N, H), bf(N, H)
timeit("layernorm fwd", lambda: ops.layernorm(x, g, g, o32, o16, 1e-12), nbytes=N * H * 10)
for cols in (H, 3 * H, I, V):
    xx = bf(N, cols)
    out = torch.zeros(cols, device=dev)
    timeit(f"colsum_bf16 [{N},{cols}]", lambda: ops.colsum_bf16(xx, out), nbytes=N * cols * 2)
    del xx

# ---- BatchNorm pieces (res_block1: M = N*256, C = 64; res_block2: M = N*64, C = 128) ----
for M, C, S in ((N * 256, 64, 16), (N * 64, 128, 8), (N * 16, 256, 4)):
    x1, x2 = f32(M, C), f32(M, C)
    sums = torch.zeros(2 * C, device=dev)
    timeit(f"bn_stats M={M} C={C}", lambda: ops.bn_stats(x1, sums), nbytes=M * C * 4)
    sc = torch.ones(C, device=dev)
    o = bf(M, C)
    timeit(f"bn_apply 1-in M={M} C={C}", lambda: ops.bn_apply(x1, sc, sc, None, None, None, o, relu=True), nbytes=M * C * 6)
    timeit(f"bn_apply 2-in remap M={M} C={C}", lambda: ops.bn_apply(x1, sc, sc, x2, sc, sc, o, relu=True, remap=True, map_hw=(S, S)),
           nbytes=M * C * 10)
    dyb, act = bf(M, C), bf(M, C)
    d1, dcat = bf(M, C), bf(M, 2 * C)
    z = [torch.zeros(C, device=dev) for _ in range(4)]
    timeit(f"bn_bwd2 (2 branches) M={M} C={C}",
           lambda: ops.bn_bwd2(dyb, act, (x1, sc, sc, sc, z[0], z[1], d1), (x2, sc, sc, sc, z[2], z[3], dcat[:, C:]), M, C, remap=True,
                               map_hw=(S, S)), nbytes=M * C * (2 * (2 + 2 + 8) + 4))
    timeit(f"bn_bwd (1 branch) M={M} C={C}", lambda: ops.bn_bwd(dyb, act, x1, sc, sc, sc, z[0], z[1], dcat[:, :C]),
           nbytes=M * C * (2 * (2 + 2 + 4) + 2))
    x1h, x2h = x1.bfloat16(), x2.bfloat16()
    timeit(f"bn_stats bf16-x M={M} C={C}", lambda: ops.bn_stats(x1h, sums), nbytes=M * C * 2)
    timeit(f"bn_apply 2-in remap bf16-x M={M} C={C}", lambda: ops.bn_apply(x1h, sc, sc, x2h, sc, sc, o, relu=True, remap=True, map_hw=(S, S)),
           nbytes=M * C * 6)
    timeit(f"bn_bwd2 (2 branches) bf16-x M={M} C={C}",
           lambda: ops.bn_bwd2(dyb, act, (x1h, sc, sc, sc, z[0], z[1], d1), (x2h, sc, sc, sc, z[2], z[3], dcat[:, C:]), M, C, remap=True,
                               map_hw=(S, S)), nbytes=M * C * (2 * (2 + 2 + 4) + 4))
    timeit(f"bn_bwd (1 branch) bf16-x M={M} C={C}", lambda: ops.bn_bwd(dyb, act, x1h, sc, sc, sc, z[0], z[1], dcat[:, :C]),
           nbytes=M * C * (2 * (2 + 2 + 2) + 2))
    del x1, x2, o, dyb, act, d1, dcat, x1h, x2h
torch.cuda.empty_cache()

# ---- attention ----
B, L, heads = 128, 128, 12
qkv, ctx, dctx, dqkv = bf(N, 3 * H), bf(N, H), bf(N, H), bf(N, 3 * H)
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
timeit("attention fwd (dropout)", lambda: ops.attention(qkv, mask, ctx, B, L, heads, drop=(0.1, 1, 1011)), flops=4.0 * B * heads * L * L * 64)
timeit("attention bwd (dropout)", lambda: ops.attention_bwd(qkv, mask, ctx, dctx, dqkv, B, L, heads, drop=(0.1, 1, 1011)),
       flops=10.0 * B * heads * L * L * 64)

# ---- GEMMs of one transformer layer (forward and backward) ----
xb, w_qkv, w_o, w_1, w_2 = bf(N, H), bf(3 * H, H), bf(H, H), bf(I, H), bf(H, I)
b3, b1, bI = f32(3 * H), f32(H), f32(I)
res = f32(N, H)
y32, qkvo, h, u = f32(N, H), bf(N, 3 * H), bf(N, I), bf(N, I)
timeit("gemm QKV  [N,768]x[2304,768]^T +bias ->bf16", lambda: ops.gemm(xb, w_qkv, qkvo, bias=b3), flops=2.0 * N * 3 * H * H)
timeit("gemm out  [N,768]x[768,768]^T +b+drop+res ->f32", lambda: ops.gemm(xb, w_o, y32, bias=b1, res=res, drop=(0.1, 1, 1012)),
       flops=2.0 * N * H * H)
timeit("gemm out  (no dropout)", lambda: ops.gemm(xb, w_o, y32, bias=b1, res=res), flops=2.0 * N * H * H)
timeit("gemm out  (bf16 out, no res)", lambda: ops.gemm(xb, w_o, ctx, bias=b1), flops=2.0 * N * H * H)
timeit("gemm FFN1 +bias+GELU_SAVE ->2x bf16", lambda: ops.gemm(xb, w_1, h, bias=bI, act=ops.ACT_GELU_SAVE, out2=u), flops=2.0 * N * I * H)
timeit("gemm FFN1 +bias+GELU ->bf16", lambda: ops.gemm(xb, w_1, h, bias=bI, act=ops.ACT_GELU), flops=2.0 * N * I * H)
timeit("gemm FFN1 +bias ->bf16", lambda: ops.gemm(xb, w_1, h, bias=bI), flops=2.0 * N * I * H)
timeit("gemm FFN2 [N,3072]x[768,3072]^T +b+drop+res ->f32", lambda: ops.gemm(h, w_2, y32, bias=b1, res=res, drop=(0.1, 1, 1013)),
       flops=2.0 * N * H * I)
dyb = bf(N, H)
timeit("gemm du = dy2 W2 * gelu'(u) ->bf16", lambda: ops.gemm(dyb, w_2, h, b_t=True, res=u, act=ops.ACT_GELU_GRAD), flops=2.0 * N * I * H)
timeit("gemm dx1 = du W1 + dy2 ->f32", lambda: ops.gemm(h, w_1, y32, b_t=True, res=res), flops=2.0 * N * I * H)
gw2, gw1, gwo, gwq = f32(H, I), f32(I, H), f32(H, H), f32(3 * H, H)
timeit("gemm dW2 = dy2^T h (split-K)", lambda: ops.gemm(dyb, h, gw2, a_t=True, b_t=True, split_k=-1), flops=2.0 * N * I * H)
timeit("gemm dW1 = du^T x1 (split-K)", lambda: ops.gemm(h, xb, gw1, a_t=True, b_t=True, split_k=-1), flops=2.0 * N * I * H)
timeit("gemm dWo = dy1^T ctx (split-K)", lambda: ops.gemm(dyb, ctx, gwo, a_t=True, b_t=True, split_k=-1), flops=2.0 * N * H * H)
timeit("gemm dctx = dy1 Wo ->bf16", lambda: ops.gemm(dyb, w_o, ctx, b_t=True), flops=2.0 * N * H * H)
timeit("gemm dWqkv = dqkv^T x (split-K)", lambda: ops.gemm(dqkv, xb, gwq, a_t=True, b_t=True, split_k=-1), flops=2.0 * N * 3 * H * H)
timeit("gemm dx = dqkv Wqkv + dy1 ->f32", lambda: ops.gemm(dqkv, w_qkv, y32, b_t=True, res=res), flops=2.0 * N * 3 * H * H)
del qkv, dqkv, h, u
torch.cuda.empty_cache()
seq, E = bf(N, H), bf(V, H)
logits = f32(N, V)
bV = f32(V)
timeit("gemm classifier [N,768]x[21128,768]^T ->f32", lambda: ops.gemm(seq, E, logits, bias=bV), flops=2.0 * N * V * H)
dl = bf(N, V)
gE = f32(V, H)
timeit("gemm dE = dlogits^T seq (split-K)", lambda: ops.gemm(dl, seq, gE, a_t=True, b_t=True, split_k=-1), flops=2.0 * N * V * H)
timeit("gemm dseq = dlogits E ->f32", lambda: ops.gemm(dl, E, y32, b_t=True), flops=2.0 * N * V * H)
uu, hh = bf(N, I), bf(N, I)
dbI = torch.zeros(I, device=dev)
timeit("gelu fwd [N,3072]", lambda: ops.gelu(uu, hh), nbytes=N * I * 4)
timeit("gelu_bwd_colsum [N,3072]", lambda: ops.gelu_bwd_colsum(hh, uu, dbI), nbytes=N * I * 6)
lg = f32(N, V)
tg = torch.randint(0, V, (N,), device=dev)
lm = torch.ones(N, dtype=torch.int64, device=dev)
rw, ls1, lse, cnt = f32(N), f32(1), f32(N), f32(1)
timeit("masked_ce fwd [N,21128]", lambda: ops.masked_ce(lg, tg, lm, rw, ls1, row_lse=lse, count=cnt), nbytes=N * V * 4)
gs = torch.ones(1, device=dev)
timeit("masked_ce bwd [N,21128]", lambda: ops.masked_ce_bwd(lg, tg, lm, lse, cnt, gs, dl), nbytes=N * V * 6)
