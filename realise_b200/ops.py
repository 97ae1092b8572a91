"""Thin Python wrappers (torch tensors in, C-ABI calls out) around librealise_b200.so.

torch is used only for device memory and the current stream; every arithmetic op on the hot path is
a kernel in csrc/.  All wrappers raise if the tensors are not CUDA tensors of the expected dtype.
"""
import ctypes

import torch

from ._lib import GemmDesc, check, lib

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_GELU_GRAD, ACT_GELU_SAVE = 0, 1, 2, 3, 4, 5

LAUNCHES = 0    # kernels launched through this module (each C-ABI call adds its kernel count)
_prof = None    # bench.py sets this to a list to get (kind, work, start_event, end_event) per call


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


class _Timed:
    """CUDA-event bracket around one C-ABI call on the current stream (only when profiling)."""

    def __init__(self, kind, work, detail=None):
        self.kind, self.work, self.detail = kind, work, detail

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None:
            self.e1.record()
            _prof.append((self.kind, self.work, self.e0, self.e1, self.detail))
        return False
_DT = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}   # RL_DT_*: the tensor's dtype IS the per-call format argument
_H16 = (torch.bfloat16, torch.float16)

# tuning knobs of the GEMM (per call through the descriptor; the library keeps no global switches)
TUNE_TILE_N = 0
TUNE_NO_PAIR = 0
SM_RESERVE = 0      # SMs the persistent GEMMs leave free (realise_b200.ddp sets it while a gradient all-reduce is in flight)

# Device-resident dropout step counter (int64 tensor of one element) handed to every dropout-bearing kernel while set:
# the kernels then use seed + *counter, read at run time (realise_b200.graphed bumps it inside the captured graph).
_COUNTER = None


class dropout_counter:
    """with ops.dropout_counter(t): ... — every launch inside passes t as its drop_counter argument."""

    def __init__(self, t):
        self.t = t

    def __enter__(self):
        global _COUNTER
        self.prev, _COUNTER = _COUNTER, self.t
        return self

    def __exit__(self, *exc):
        global _COUNTER
        _COUNTER = self.prev
        return False


def _ctr():
    return ctypes.c_void_p(_COUNTER.data_ptr()) if _COUNTER is not None else None


def _dt(t):
    return ctypes.c_int32(_DT[t.dtype])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (realise_b200 has no CPU path)")
    if dtype is torch.bfloat16 and t.dtype is torch.float16:
        return            # any "bf16" tensor may be IEEE fp16: the call passes its format along
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a, b, out, *, scale=None, bias=None, res=None, act=ACT_NONE, out2=None, a_t=False, b_t=False, drop=None,
         split_k=0, colsum=None, colsumsq=None):
    """out[M,N] = act((A @ B^T) * scale + bias + res) with A = a [M,K] (or a^T when a_t: a is stored [K,M],
    MN-major operand) and B = b [N,K] (or b^T when b_t: b is stored [K,N]).  a, b bf16; out bf16 or f32."""
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    (K, M) = a.shape if a_t else a.shape[::-1]
    (Kb, N) = b.shape if b_t else b.shape[::-1]
    assert Kb == K and tuple(out.shape) == (M, N), (a.shape, b.shape, out.shape, a_t, b_t)
    d = GemmDesc()
    d.a, d.b = a.data_ptr(), b.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb = a.stride(0), b.stride(0)
    d.a_major, d.b_major = int(a_t), int(b_t)
    d.a_dtype, d.b_dtype = _DT[a.dtype], _DT[b.dtype]
    if drop:
        d.drop_p, d.drop_seed, d.drop_site = drop
        d.drop_counter = _COUNTER.data_ptr() if _COUNTER is not None else None
    d.split_k = split_k
    d.a_mode = 0
    _fill_colsums(d, colsum, colsumsq, N)
    _fill_epilogue(d, out, scale, bias, res, act, out2, 0)
    with _Timed("gemm", 2.0 * M * N * K, (M, N, K, int(a_t), int(b_t), split_k, act, str(out.dtype)[6:], res is not None,
                                          drop is not None)):
        check(lib().rl_gemm_bf16(ctypes.byref(d), _stream()), "rl_gemm_bf16")
    _count()
    return out


def _fill_colsums(d, colsum, colsumsq, N):
    for name, t in (("colsum", colsum), ("colsumsq", colsumsq)):
        if t is not None:
            _req(t, torch.float32, name)
            assert t.is_contiguous() and t.numel() == N
            setattr(d, name, t.data_ptr())


def conv_gemm(x, w, out, *, nimg, H, W, planes, taps, scale=None, bias=None, res=None,
              act=ACT_NONE, out_remap=0, remap_plane=0, c_use=0, colsum=None, colsumsq=None):
    """Implicit-GEMM convolution.  x: bf16 activation [nimg, planes, H, W, C] (contiguous);
    w: bf16 [Cout, ntaps*C] tap-major; taps: list of (dw, dh, plane); output rows are (img, oh, ow)
    over an H x W map; out_remap=1 writes rows parity-split for a following stride-2 conv."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    assert x.is_contiguous() and w.stride(1) == 1
    C = x.shape[-1]
    assert x.numel() == nimg * planes * H * W * C
    M, N, K = nimg * H * W, w.shape[0], len(taps) * (c_use or C)
    assert w.shape[1] == K and out.shape[0] == (4 * M if out_remap == 2 else M) and out.shape[1] == N
    d = GemmDesc()
    d.a, d.b = x.data_ptr(), w.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb = C, w.stride(0)
    d.a_mode = 1
    d.a_dtype, d.b_dtype = _DT[x.dtype], _DT[w.dtype]
    d.conv_C, d.conv_W, d.conv_H, d.conv_P, d.conv_NIMG = C, W, H, planes, nimg
    d.ntaps = len(taps)
    d.conv_Cuse = c_use
    d.remap_plane = remap_plane
    for i, (dw, dh, pl) in enumerate(taps):
        d.tap_dw[i], d.tap_dh[i], d.tap_plane[i] = dw, dh, pl
    _fill_colsums(d, colsum, colsumsq, N)
    _fill_epilogue(d, out, scale, bias, res, act, None, out_remap)
    with _Timed("conv_gemm", 2.0 * M * N * K, (M, N, K, "conv", len(taps), out_remap, act, str(out.dtype)[6:], res is not None, False)):
        check(lib().rl_gemm_bf16(ctypes.byref(d), _stream()), "rl_gemm_bf16(conv)")
    _count()
    return out


def conv_wgrad(dy, x, out, *, nimg, H, W, planes, taps, split_k=-1):
    """Conv weight gradient without an im2col matrix: out[co, t*C + ci] += sum over output pixels m = (img, oh, ow) of
    dy[m, co] * x[img, plane_t, oh+dh_t, ow+dw_t, ci].  dy: bf16 [nimg*H*W, Cout]; x: bf16 activation
    [nimg, planes, H, W, C] (C % 64 == 0); out: f32 [Cout, ntaps*C], accumulated into (split-K atomics)."""
    _req(dy, torch.bfloat16, "dy")
    _req(x, torch.bfloat16, "x")
    _req(out, torch.float32, "out")
    assert x.is_contiguous() and dy.stride(1) == 1 and out.stride(1) == 1
    C = x.shape[-1]
    Kpix, M = dy.shape
    N = len(taps) * C
    assert Kpix == nimg * H * W and x.numel() == nimg * planes * H * W * C and tuple(out.shape) == (M, N)
    d = GemmDesc()
    d.a, d.b = dy.data_ptr(), x.data_ptr()
    d.M, d.N, d.K = M, N, Kpix
    d.lda, d.ldb = dy.stride(0), C
    d.a_major, d.b_major, d.b_mode, d.a_mode = 1, 1, 1, 0
    d.a_dtype, d.b_dtype = _DT[dy.dtype], _DT[x.dtype]
    d.conv_C, d.conv_W, d.conv_H, d.conv_P, d.conv_NIMG = C, W, H, planes, nimg
    d.ntaps = len(taps)
    for i, (dw, dh, pl) in enumerate(taps):
        d.tap_dw[i], d.tap_dh[i], d.tap_plane[i] = dw, dh, pl
    d.split_k = split_k
    _fill_epilogue(d, out, None, None, None, ACT_NONE, None, 0)
    with _Timed("gemm", 2.0 * M * N * Kpix, (M, N, Kpix, "wgrad", len(taps), split_k, 0, "float32", False, False)):
        check(lib().rl_gemm_bf16(ctypes.byref(d), _stream()), "rl_gemm_bf16(conv wgrad)")
    _count()
    return out


def _fill_epilogue(d, out, scale, bias, res, act, out2, out_remap):
    if not out.is_cuda or out.dtype not in _DT:
        raise RuntimeError("out must be a CUDA bf16/f32 tensor")
    d.out, d.ldo, d.out_dtype = out.data_ptr(), out.stride(0), _DT[out.dtype]
    d.tune_tile_n, d.tune_no_pair, d.sm_reserve = TUNE_TILE_N, TUNE_NO_PAIR, SM_RESERVE
    if out2 is not None:
        _req(out2, torch.bfloat16, "out2")
        assert out2.dtype == out.dtype, "out2 shares out's 16-bit format"
        d.out2, d.ldo2 = out2.data_ptr(), out2.stride(0)
    if scale is not None:
        _req(scale, torch.float32, "scale")
        d.scale = scale.data_ptr()
    if bias is not None:
        _req(bias, torch.float32, "bias")
        d.bias = bias.data_ptr()
    if res is not None:
        assert res.is_cuda and res.dtype in _DT and res.stride(-1) == 1
        d.res, d.ldr, d.res_dtype = res.data_ptr(), res.stride(0), _DT[res.dtype]
    d.act = act
    d.out_remap = out_remap


_I64P = ctypes.POINTER(ctypes.c_int64)


def _c(v):
    return ctypes.c_int64(int(v))


def _drop(drop):
    """drop = None or (p, seed, site) -> ctypes (float, uint64, uint32)."""
    p, seed, site = drop if drop else (0.0, 0, 0)
    return ctypes.c_float(p), ctypes.c_uint64(seed), ctypes.c_uint32(site)


def attention(qkv, mask, ctx, B, L, heads, drop=None, lse=None):
    """ctx = softmax(QK^T/8 + (1-mask)*-1e4) V per head; qkv bf16 [B*L, 3*heads*64], mask int64 [B, L].
    lse (optional f32 [B*heads*L]) receives the log2-domain logsumexp of every query row for the backward."""
    _req(qkv, torch.bfloat16, "qkv")
    _req(ctx, torch.bfloat16, "ctx")
    _req(mask, torch.int64, "mask")
    assert qkv.is_contiguous() and ctx.is_contiguous() and mask.is_contiguous()
    with _Timed("attention", 4.0 * B * heads * L * L * 64):
        assert ctx.dtype == qkv.dtype
        check(lib().rl_attention_fwd(_ptr(qkv), _ptr(mask), _ptr(ctx), _ptr(lse), _c(B), _c(L), _c(heads), _c(64), _dt(qkv),
                                     *_drop(drop), _ctr(), _stream()), "rl_attention_fwd")
    _count()
    return ctx


def layernorm(x, gamma, beta, out_f32, out_bf16, eps, drop=None, drop_f32=False):
    _req(x, torch.float32, "x")
    rows, H = x.shape
    nbytes = rows * H * (4 + (4 if out_f32 is not None else 0) + (2 if out_bf16 is not None else 0))
    with _Timed("layernorm", nbytes):
        check(lib().rl_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out_f32), _ptr(out_bf16), _c(rows), _c(H),
                                     ctypes.c_float(eps), *_drop(drop), _ctr(), ctypes.c_int32(int(drop_f32)),
                                     _dt(out_bf16) if out_bf16 is not None else ctypes.c_int32(0), _stream()),
              "rl_layernorm_fwd")
    _count()


def embed_ln(ids, word, inputs_embeds, pos, type0, gamma, beta, out_f32, out_bf16, rows, L, H, pos_mode, eps,
             pre_out=None, drop=None):
    if ids is not None:
        _req(ids, torch.int64, "ids")
    with _Timed("embed_ln", rows * H * 10):
        check(lib().rl_embed_ln_fwd(_ptr(ids), _ptr(word), _ptr(inputs_embeds), _ptr(pos), _ptr(type0), _ptr(gamma),
                                    _ptr(beta), _ptr(out_f32), _ptr(out_bf16), _ptr(pre_out), _c(rows), _c(L), _c(H),
                                    ctypes.c_int32(pos_mode), ctypes.c_float(eps), *_drop(drop), _ctr(),
                                    _dt(out_bf16) if out_bf16 is not None else ctypes.c_int32(0), _stream()),
              "rl_embed_ln_fwd")
    _count()


def gate_fuse(mods, sum_mode, mask, gate_w, gate_b, ws, out, gates_out, B, L, H):
    m = list(mods) + [None] * (3 - len(mods))
    for t in mods:
        _req(t, torch.float32, "modality")
    with _Timed("gate_fuse", B * L * H * 4 * (len(mods) + 1)):
        check(lib().rl_gate_fuse_fwd(_ptr(m[0]), _ptr(m[1]), _ptr(m[2]), ctypes.c_int32(len(mods)),
                                     ctypes.c_int32(1 if sum_mode else 0), _ptr(mask), _ptr(gate_w), _ptr(gate_b),
                                     _ptr(ws), _ptr(out), _ptr(gates_out), _c(B), _c(L), _c(H), _stream()),
              "rl_gate_fuse_fwd")
    _count(1 if sum_mode else 2)


def masked_ce(logits, tgt, loss_mask, row_ws, loss, row_lse=None, count=None):
    assert logits.is_cuda and logits.dtype in (torch.float32, torch.float16)
    _req(tgt, torch.int64, "tgt")
    _req(loss_mask, torch.int64, "loss_mask")
    rows, V = logits.shape
    with _Timed("masked_ce", rows * V * logits.element_size()):
        check(lib().rl_masked_ce_fwd(_ptr(logits), _dt(logits), _ptr(tgt), _ptr(loss_mask), _ptr(row_ws), _ptr(loss), _ptr(row_lse),
                                     _ptr(count), _c(rows), _c(V), _c(logits.stride(0)), _stream()), "rl_masked_ce_fwd")
    _count(2)


def gru_input_table(emb, w_ih, b_ih, table):
    V, H = emb.shape
    check(lib().rl_gru_input_table(_ptr(emb), _ptr(w_ih), _ptr(b_ih), _ptr(table), _c(V), _c(H), _stream()),
          "rl_gru_input_table")
    _count()


def gru_step(gh, b_hh, table, pho_idx, lens, h_prev, h_out, h_out_bf16, t):
    _req(pho_idx, torch.int64, "pho_idx")
    _req(lens, torch.int32, "lens")
    rows, T = pho_idx.shape
    H = h_out.shape[1]
    with _Timed("gru_step", rows * H * (12 + 6 + 4 + 6)):
        check(lib().rl_gru_step_fwd(_ptr(gh), _ptr(b_hh), _ptr(table), _ptr(pho_idx), _ptr(lens), _ptr(h_prev),
                                    _ptr(h_out), _ptr(h_out_bf16), _c(rows), _c(H), _c(T), _c(t), _dt(h_out_bf16), _stream()),
              "rl_gru_step_fwd")
    _count()


def glyph_stem(glyphs, ids, w1, wsc, scale1, shift1, scale_sc, shift_sc, y1, ysc, n_img, C):
    _req(glyphs, torch.float32, "glyphs")
    _req(ids, torch.int64, "ids")
    # algorithmic bytes per glyph: C*32*32*4 read + 2 * 16*16*64*2 written (conv1 out + shortcut out)
    with _Timed("glyph_stem", n_img * (C * 4096 + 2 * 32768)):
        check(lib().rl_glyph_stem_fwd(_ptr(glyphs), _ptr(ids), _ptr(w1), _ptr(wsc), _ptr(scale1), _ptr(shift1),
                                      _ptr(scale_sc), _ptr(shift_sc), _ptr(y1), _ptr(ysc), _c(n_img),
                                      ctypes.c_int32(C), _stream()), "rl_glyph_stem_fwd")
    _count()


def argmax_rows(logits, out):
    _req(logits, torch.float32, "logits")
    _req(out, torch.int64, "out")
    rows, V = logits.shape
    with _Timed("argmax", rows * V * 4):
        check(lib().rl_argmax_rows(_ptr(logits), _ptr(out), _c(rows), _c(V), _c(logits.stride(0)), _stream()),
              "rl_argmax_rows")
    _count()
    return out


def glyph_block1(glyphs, ids, w1p, wscp, w2p, t1, t2s, out, n_img, C):
    """Fused res_block1 (eval): glyph gather + conv1/BN/ReLU + conv2/BN + shortcut/BN + ReLU."""
    _req(glyphs, torch.float32, "glyphs")
    _req(ids, torch.int64, "ids")
    for t, n in ((w1p, "w1p"), (wscp, "wscp"), (w2p, "w2p"), (out, "out")):
        _req(t, torch.bfloat16, n)
    # algorithmic bytes per glyph: C*32*32*4 read + 16*16*64*2 written; flops: conv1 + shortcut + conv2
    with _Timed("glyph_block1", n_img * (C * 4096 + 32768)):
        check(lib().rl_glyph_block1_fwd(_ptr(glyphs), _ptr(ids), _ptr(w1p), _ptr(wscp), _ptr(w2p), _ptr(t1),
                                        _ptr(t2s), _ptr(out), _c(n_img), ctypes.c_int32(C), _stream()),
              "rl_glyph_block1_fwd")
    _count()


# ---------------------------------------------------------------------------------------------------------
# training path
# ---------------------------------------------------------------------------------------------------------
def attention_bwd(qkv, mask, ctx, dctx, dqkv, B, L, heads, drop=None, lse=None, dbias=None):
    """dbias (optional f32 [3*heads*64], L <= 128): += column sums of dqkv (the fused q/k/v bias gradient)."""
    for t, n in ((qkv, "qkv"), (ctx, "ctx"), (dctx, "dctx"), (dqkv, "dqkv")):
        _req(t, torch.bfloat16, n)
    assert ctx.dtype == qkv.dtype == dctx.dtype == dqkv.dtype
    if dbias is not None:
        _req(dbias, torch.float32, "dbias")
        assert dbias.numel() == 3 * heads * 64 and dbias.is_contiguous()
    with _Timed("attention_bwd", 14.0 * B * heads * L * L * 64):
        check(lib().rl_attention_bwd(_ptr(qkv), _ptr(mask), _ptr(ctx), _ptr(dctx), _ptr(dqkv), _ptr(dbias), _ptr(lse), _c(B), _c(L),
                                     _c(heads), _c(64), _dt(qkv), *_drop(drop), _ctr(), _stream()), "rl_attention_bwd")
    _count()


def layernorm_bwd(dy, x, gamma, add_in, dx, dx_bf16, dgamma, dbeta, dxsum, eps, drop_p=0.0, drop_seed=0, site_in=0,
                  site_out=0):
    """site_in: dropout site applied to the LN OUTPUT in the forward (dy is masked); site_out: dropout site applied
    to the linear output that fed the LN input (dx_bf16 / dxsum are masked).  0 = no dropout there."""
    _req(dy, torch.float32, "dy")
    _req(x, torch.float32, "x")
    rows, H = x.shape
    with _Timed("layernorm_bwd", rows * H * 18):
        check(lib().rl_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(add_in), _ptr(dx), _ptr(dx_bf16), _ptr(dgamma),
                                     _ptr(dbeta), _ptr(dxsum), _c(rows), _c(H), ctypes.c_float(eps), ctypes.c_float(drop_p),
                                     ctypes.c_uint64(drop_seed), ctypes.c_uint32(site_in), ctypes.c_uint32(site_out),
                                     _ctr(), _stream()), "rl_layernorm_bwd")
    _count()


def colsum_bf16(x, out):
    _req(x, torch.bfloat16, "x")
    _req(out, torch.float32, "out")
    rows, cols = x.shape
    check(lib().rl_colsum_bf16(_ptr(x), _ptr(out), _c(rows), _c(cols), _c(x.stride(0)), _stream()), "rl_colsum_bf16")
    _count()


def gather_rows(table, ids, out):
    """out[r] = table[ids[r]] for f32 rows (glyph cache lookup)."""
    _req(table, torch.float32, "table")
    _req(ids, torch.int64, "ids")
    _req(out, torch.float32, "out")
    rows, H = out.shape
    assert table.is_contiguous() and out.is_contiguous() and ids.numel() == rows and table.shape[1] == H
    check(lib().rl_gather_rows_f32(_ptr(table), _ptr(ids), _ptr(out), _c(rows), _c(H), _stream()), "rl_gather_rows_f32")
    _count()


def split3_bf16(x, out):
    """out[r] = [bf16(x) | bf16(x - bf16(x)) | bf16(x)] (split-precision classifier operand)."""
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    rows, cols = x.shape
    assert x.is_contiguous() and out.is_contiguous() and tuple(out.shape) == (rows, 3 * cols)
    check(lib().rl_split3_bf16(_ptr(x), _ptr(out), _c(rows), _c(cols), _dt(out), _stream()), "rl_split3_bf16")
    _count()


def gelu(u, h):
    """h = gelu_erf(u), bf16 -> bf16 (may be in place)."""
    _req(u, torch.bfloat16, "u")
    _req(h, torch.bfloat16, "h")
    assert u.is_contiguous() and h.is_contiguous() and u.numel() == h.numel() and u.dtype == h.dtype
    with _Timed("gelu", u.numel() * 4):
        check(lib().rl_gelu_fwd(_ptr(u), _ptr(h), _c(u.numel()), _dt(u), _stream()), "rl_gelu_fwd")
    _count()


def gelu_bwd_colsum(t, u, dbias):
    """t <- t * gelu'(u) in place ([rows, cols] bf16); dbias[cols] += column sums of the result."""
    _req(t, torch.bfloat16, "t")
    _req(u, torch.bfloat16, "u")
    rows, cols = t.shape
    assert t.stride(1) == 1 and u.stride() == t.stride()
    with _Timed("gelu_bwd_colsum", t.numel() * 6):
        check(lib().rl_gelu_bwd_colsum(_ptr(t), _ptr(u), _ptr(dbias), _c(rows), _c(cols), _c(t.stride(0)), _dt(u), _stream()),
              "rl_gelu_bwd_colsum")
    _count()


def masked_ce_bwd(logits, tgt, loss_mask, row_lse, count, gscale, dlogits):
    rows, V = logits.shape
    _req(dlogits, torch.bfloat16, "dlogits")
    assert logits.dtype in (torch.float32, torch.float16)
    check(lib().rl_masked_ce_bwd(_ptr(logits), _dt(logits), _ptr(tgt), _ptr(loss_mask), _ptr(row_lse), _ptr(count), _ptr(gscale),
                                 _ptr(dlogits), _c(rows), _c(V), _c(logits.stride(0)), _c(dlogits.stride(0)), _stream()),
          "rl_masked_ce_bwd")
    _count()


def embed_bwd(de, ids, dword, dpos, rows, L, H, pos_mode):
    check(lib().rl_embed_bwd(_ptr(de), _ptr(ids), _ptr(dword), _ptr(dpos), _c(rows), _c(L), _c(H),
                             ctypes.c_int32(pos_mode), _stream()), "rl_embed_bwd")
    _count()


def gate_fuse_bwd(dhid, mods, mask, gates, gate_w, dmods, dgate_w, dgate_b, ws, B, L, H):
    m = list(mods) + [None] * (3 - len(mods))
    dm = list(dmods) + [None] * (3 - len(dmods))
    check(lib().rl_gate_fuse_bwd(_ptr(dhid), _ptr(m[0]), _ptr(m[1]), _ptr(m[2]), ctypes.c_int32(len(mods)), _ptr(mask),
                                 _ptr(gates), _ptr(gate_w), _ptr(dm[0]), _ptr(dm[1]), _ptr(dm[2]), _ptr(dgate_w),
                                 _ptr(dgate_b), _ptr(ws), _c(B), _c(L), _c(H), _stream()), "rl_gate_fuse_bwd")
    _count(4 + len(mods))


def dropout_mask(n, p, seed, site, device="cuda"):
    """uint8 keep mask of a dropout site (what the kernels compute on the fly) — for tests."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    check(lib().rl_dropout_mask(_ptr(out), _c(n), ctypes.c_float(p), ctypes.c_uint64(seed), ctypes.c_uint32(site), _ctr(),
                                _stream()), "rl_dropout_mask")
    return out


def gru_step_bwd(dh, gh, b_hh, table, pho_idx, lens, h_prev, dh_prev, dgi, dgh, onehot, t):
    rows, T = pho_idx.shape
    H = dh.shape[1]
    check(lib().rl_gru_step_bwd(_ptr(dh), _ptr(gh), _ptr(b_hh), _ptr(table), _ptr(pho_idx), _ptr(lens), _ptr(h_prev),
                                _ptr(dh_prev), _ptr(dgi), _ptr(dgh), _ptr(onehot), _c(rows), _c(H), _c(T), _c(t), _stream()),
          "rl_gru_step_bwd")
    _count()


def gru_table_bwd(dtable, emb, w_ih, dw_ih, db_ih, demb):
    V, H = emb.shape
    check(lib().rl_gru_table_bwd(_ptr(dtable), _ptr(emb), _ptr(w_ih), _ptr(dw_ih), _ptr(db_ih), _ptr(demb), _c(V), _c(H),
                                 _stream()), "rl_gru_table_bwd")
    _count(2)


# ---- CharResNet training pieces ----
def bn_stats(x, sums):
    M, C = x.shape
    check(lib().rl_bn_stats(_ptr(x), ctypes.c_int32(_DT[x.dtype]), _ptr(sums), _c(M), _c(C), _c(x.stride(0)), _stream()),
          "rl_bn_stats")
    _count()


def bn_finalize(sums, gamma, beta, running_mean, running_var, nbt, scale, shift, mean, rstd, M, momentum=0.1, eps=1e-5):
    C = gamma.numel()
    check(lib().rl_bn_finalize(_ptr(sums), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(nbt),
                               _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), _c(M), _c(C), ctypes.c_float(momentum),
                               ctypes.c_float(eps), _stream()), "rl_bn_finalize")
    _count()


def bn_apply(x1, sc1, sh1, x2, sc2, sh2, out, relu, remap=False, map_hw=(1, 1)):
    M, C = x1.shape
    check(lib().rl_bn_apply(_ptr(x1), _ptr(sc1), _ptr(sh1), _ptr(x2), _ptr(sc2), _ptr(sh2), ctypes.c_int32(_DT[x1.dtype]),
                            _ptr(out), ctypes.c_int32(_DT[out.dtype]), ctypes.c_int32(int(relu)), _c(M), _c(C), ctypes.c_int32(int(remap)),
                            ctypes.c_int32(map_hw[0]), ctypes.c_int32(map_hw[1]), _stream()), "rl_bn_apply")
    _count()


def bn_bwd(dy, act_out, x, mean, rstd, gamma, dbeta, dgamma, dx, remap=False, map_hw=(1, 1), fwd=None):
    """fwd = (scale, shift) the forward applied (bn_finalize): the ReLU mask is re-derived from x and act_out is not read."""
    M, C = x.shape
    sc, sh = fwd if fwd is not None else (None, None)
    with _Timed("bn_bwd", 0):
        check(lib().rl_bn_bwd(_ptr(dy), ctypes.c_int32(_DT[dy.dtype]), _ptr(act_out),
                              ctypes.c_int32(_DT[act_out.dtype] if act_out is not None else 0), _ptr(x),
                              ctypes.c_int32(_DT[x.dtype]), _ptr(mean), _ptr(rstd),
                              _ptr(gamma), _ptr(dbeta), _ptr(dgamma), _ptr(dx), _c(dx.stride(0)), _c(M), _c(C),
                              ctypes.c_int32(int(remap)), ctypes.c_int32(map_hw[0]), ctypes.c_int32(map_hw[1]), _ptr(sc), _ptr(sh),
                              _stream()), "rl_bn_bwd")
    _count(2)


def bn_bwd2(dy, act_out, br1, br2, M, C, remap=False, map_hw=(1, 1), fwd=None):
    """Backward of out = relu(bn_a(x_a) + bn_b(x_b)) for both branches in one reduce + one apply pass.
    br = (x f32 or bf16 [M,C], mean, rstd, gamma, dbeta, dgamma, dx bf16 view with row stride).
    fwd = ((scale_a, shift_a), (scale_b, shift_b)) of the forward: the ReLU mask is re-derived from x_a, x_b (the same
    fmaf arithmetic as bn_apply) and act_out is not read."""
    args = []
    for br in (br1, br2):
        x, mean, rstd, gamma, dbeta, dgamma, dx = br
        assert x.is_cuda and x.dtype == br1[0].dtype and x.dtype in _DT
        _req(dx, torch.bfloat16, "dx")
        args += [_ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dbeta), _ptr(dgamma), _ptr(dx), _c(dx.stride(0))]
    f = [_ptr(t) for pair in (fwd if fwd is not None else ((None, None), (None, None))) for t in pair]
    with _Timed("bn_bwd", 0):
        check(lib().rl_bn_bwd2(_ptr(dy), ctypes.c_int32(_DT[dy.dtype]), _ptr(act_out),
                               ctypes.c_int32(_DT[act_out.dtype] if act_out is not None else 0),
                               ctypes.c_int32(_DT[br1[0].dtype]), *args, _c(M), _c(C), ctypes.c_int32(int(remap)),
                               ctypes.c_int32(map_hw[0]), ctypes.c_int32(map_hw[1]), *f, _stream()), "rl_bn_bwd2")
    _count(2)


def im2col(x, col, nimg, C, W, H, P, taps):
    n = len(taps)
    arr = ctypes.c_int8 * n
    dw, dh, pl = arr(*[t[0] for t in taps]), arr(*[t[1] for t in taps]), arr(*[t[2] for t in taps])
    check(lib().rl_im2col_bf16(_ptr(x), _ptr(col), _c(nimg), ctypes.c_int32(C), ctypes.c_int32(W), ctypes.c_int32(H),
                               ctypes.c_int32(P), ctypes.c_int32(n), dw, dh, pl, _stream()), "rl_im2col_bf16")
    _count()


def glyph_im2col(glyphs, ids, col1, colsc, n_img, C):
    check(lib().rl_glyph_im2col(_ptr(glyphs), _ptr(ids), _ptr(col1), _ptr(colsc), _c(n_img), ctypes.c_int32(C), _stream()),
          "rl_glyph_im2col")
    _count()


def mt_gather(table, chunks, n_chunks, srcs):
    check(lib().rl_mt_gather(_ptr(table), _ptr(chunks), _c(n_chunks), _ptr(srcs), _stream()), "rl_mt_gather")
    _count()


def _wrap_untimed():
    """Give every wrapper without its own _Timed bracket one (work = 0), so a profiling pass sees the whole step."""
    import functools

    def make(fn, name):
        @functools.wraps(fn)
        def inner(*a, **k):
            if _prof is None:
                return fn(*a, **k)
            with _Timed(name, 0):
                return fn(*a, **k)
        return inner

    for name in ("colsum_bf16", "masked_ce_bwd", "embed_bwd", "gate_fuse_bwd", "gru_step_bwd", "gru_table_bwd",
                 "gru_input_table", "bn_stats", "bn_finalize", "bn_apply", "bn_bwd", "im2col", "glyph_im2col", "mt_gather"):
        globals()[name] = make(globals()[name], name)


_wrap_untimed()
