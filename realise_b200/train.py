"""Training-mode forward + backward of the hot path on librealise_b200.so.

`TrainEngine` runs the train-mode forward (saving the activations the backward needs) and the backward pass
as a fixed sequence of C-ABI kernel calls; `torch.autograd` only sees ONE custom Function whose backward attaches
the parameter gradients, so `loss.backward()` in the reference's loop (src/run.py:191-200) keeps working.

Coverage: every branch of SpellBertPho2ResArch3 / its ablations (src/models_abla.py) — the three BertModel stacks with
counter-based dropout (hidden 0.1 / attention 0.1 as in the reference config), the pinyin GRU (backward through time),
CharResNet with batch-statistics BatchNorm (running statistics updated like nn.BatchNorm2d), gated or sum fusion, the
tied classifier and the masked CE.  Forward tensors are fp16 (or bf16), gradients bf16, accumulation / residual stream /
LayerNorm / softmax / BatchNorm statistics fp32.  Gradients of successive backward() calls ACCUMULATE until the
optimizer step or zero_grad() consumes them (src/run.py:193-205, gradient_accumulation_steps).
"""
import ctypes
import os

import torch

from . import ops

F32, BF16 = torch.float32, torch.bfloat16


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole model: forward -> (loss, logits); backward -> every parameter gradient."""

    @staticmethod
    def forward(ctx, engine, inputs, *params):
        loss, logits = engine.forward(inputs)
        ctx.engine = engine
        ctx.saved = engine.saved          # the activations belong to THIS forward: a second forward before the backward
        engine.saved = None               # (e.g. a train-mode loss probe) neither overwrites nor leaks them
        ctx.mark_non_differentiable(logits)
        ctx.set_materialize_grads(False)  # or autograd zero-fills a [B, L, 21128] gradient for the logits on every backward (0.18 ms)
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        eng = ctx.engine
        if ctx.saved is None:
            raise RuntimeError("realise_b200: backward through the same forward twice (activations already released)")
        eng.saved, ctx.saved = ctx.saved, None
        if gloss is None:                 # only reachable by differentiating a function of the logits alone
            raise RuntimeError("realise_b200: the training graph differentiates the loss; logits carry no gradient")
        eng.backward_and_sync(gloss)
        return (None, None) + (None,) * len(eng.params)


class _GatherEntry(ctypes.Structure):
    _fields_ = [("dst", ctypes.c_void_p), ("map", ctypes.c_void_p), ("n", ctypes.c_int64), ("dst_dtype", ctypes.c_int32),
                ("pad", ctypes.c_int32)]


class MultiGather:
    """One rl_mt_gather launch: dst_k[i] = cast(src[map_k[i] >> 24][map_k[i] & 0xFFFFFF]) (0 where the map is negative)
    for a fixed set of destination tensors.  Maps are built once on the host with ordinary torch indexing (the layout
    code runs on tensors of element CODES instead of values), the launch replaces the per-step cat/transpose/cast chain."""
    CHUNK = 4096

    def __init__(self, device):
        self.dev, self.srcs, self.entries, self._ready = device, [], [], False

    def source(self, t):
        """Register an f32 source tensor; returns an int32 tensor of its element codes, shaped like t."""
        assert t.dtype is F32 and t.is_contiguous() and t.numel() < (1 << 24) and len(self.srcs) < 128
        self.srcs.append(t)
        return (torch.arange(t.numel(), dtype=torch.int32) + ((len(self.srcs) - 1) << 24)).view(t.shape)

    def add(self, dst, codes):
        assert dst.is_contiguous() and dst.numel() == codes.numel()
        self.entries.append((dst, codes.reshape(-1).to(torch.int32).contiguous().to(self.dev)))

    def run(self):
        if not self.entries:
            return
        if not self._ready:
            arr = (_GatherEntry * len(self.entries))()
            chunks = []
            for i, (dst, m) in enumerate(self.entries):
                arr[i] = _GatherEntry(dst.data_ptr(), m.data_ptr(), dst.numel(), ops._DT[dst.dtype], 0)
                chunks += [(i, c) for c in range((dst.numel() + self.CHUNK - 1) // self.CHUNK)]
            self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
            self._chunks = torch.tensor(chunks, dtype=torch.int32).to(self.dev).contiguous()
            self._srcs = torch.tensor([t.data_ptr() for t in self.srcs], dtype=torch.int64).to(self.dev)
            self._ready = True
        ops.mt_gather(self._table, self._chunks, self._chunks.shape[0], self._srcs)


class ZeroPool:
    """Scratch that must start at zero (split-K accumulators, BatchNorm sums, ...) comes out of ONE buffer cleared by ONE
    memset per step instead of a fill kernel per tensor.  The first step of a shape records the sizes (and uses
    torch.zeros); later steps hand out 256-byte aligned slices."""

    def __init__(self, device):
        self.dev, self.sizes, self.buf, self.key, self.off, self.recording = device, {}, None, None, 0, True

    def begin(self, key):
        self.key, self.off = key, 0
        need = self.sizes.get(key)
        self.recording = need is None
        if not self.recording:
            if self.buf is None or self.buf.numel() < need:
                self.buf = torch.empty(need, device=self.dev, dtype=F32)
            self.buf[:need].zero_()

    def get(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        start, self.off = self.off, self.off + (n + 63) // 64 * 64
        if self.recording:
            self.sizes[self.key] = self.off
            return torch.zeros(shape, device=self.dev, dtype=F32)
        if self.off > self.sizes[self.key]:      # a different allocation sequence than the recorded one: re-record
            self.recording = True
            self.sizes[self.key] = self.off
            return torch.zeros(shape, device=self.dev, dtype=F32)
        return self.buf[start:start + n].view(shape)


class TrainEngine:
    def __init__(self, model):
        self.m = model
        c = model.config
        if c.with_res == "yes" and not 1 <= c.num_fonts <= 3:
            raise NotImplementedError("glyph kernels are built for 1..3 fonts (src/run.py --num_fonts; 9 taps x C <= 32)")
        self.saved = None
        self._pending = False         # gradients of an earlier backward() are still waiting for the optimizer
        self._accum = None
        self.debug = None             # tests set a dict to receive intermediate gradients
        self._res = None              # conv operand layouts + their refresh launch, built on the first training forward
        self.multistream = False      # experiment (off): the three encoders (and their backward passes) on three CUDA streams.
        self._streams = None          # Measured on B200: no gain — the persistent GEMMs hold every SM's register file, so the
                                      # other branches' kernels only fill wave tails (42.8 vs 42.0 ms/step within clock noise)
        self.fused_gelu_grad = os.environ.get("RL_FUSED_GELU_GRAD", "0") == "1"   # du and db1 out of the dgrad GEMM's epilogue
        # BatchNorm backward re-derives the ReLU mask from the raw conv outputs instead of reading the activation
        self.bn_mask_recompute = os.environ.get("RL_BN_MASK_RECOMPUTE", "1") == "1"
        self.zpool = ZeroPool(model.classifier.bias.device)
        self.raw_bf16 = True          # res_block1-2 keep raw conv outputs in bf16 (see _resnet_fwd)
        self.seed = 0x5EED            # dropout seed of the next step; advanced every forward (set_seed to pin it)
        # trainable parameters in a fixed order
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self._layout()

    def _layout(self):
        """All gradients live in ONE flat fp32 buffer (stable pointers for the fused optimizer, NCCL all-reduce straight
        out of it), zeroed once per step and written by split-K GEMMs / column-sum kernels.  Blocks are laid out in the
        order in which the BACKWARD pass completes them — output_block + gate, pinyin branch, glyph branch, bert layers
        11..0, bert embeddings (the tied classifier / word-embedding matrix is finished last, by the embedding
        scatter) — so that data parallelism can all-reduce contiguous BUCKETS while the rest of the backward still runs
        (`self.buckets`, realise_b200.ddp.DataParallel(overlap=True)).  The q/k/v weights (biases) of a layer are
        adjacent so that the fused [3H, H] wgrad GEMM ([3H] column sum) writes them in one go.  Parameters that never
        receive a gradient (poolers, the unused word embeddings of pho_model / output_block) are left out, like autograd
        leaves their .grad at None in the reference."""
        m = self.m
        H = m.config.hidden_size
        dev = m.classifier.bias.device
        no_grad = set()
        stacks = [m.bert, m.output_block] + ([m.pho_model] if hasattr(m, "pho_model") else [])
        for stack in stacks:
            no_grad.update(id(p) for p in stack.pooler.parameters())
            if stack is not m.bert:
                no_grad.add(id(stack.embeddings.word_embeddings.weight))
        tied = m.classifier.weight is m.bert.embeddings.word_embeddings.weight
        self._fused = {}
        trainable = {id(p) for p in self.params}
        entries, placed = [], set()          # (kind, attention module, [params])

        def add(ps, kind="p", att=None):
            ps = [p for p in ps if id(p) in trainable and id(p) not in no_grad and id(p) not in placed]
            if ps:
                placed.update(id(p) for p in ps)
                entries.append((kind, att, ps))

        def add_layer(lyr):
            sa = lyr.attention.self
            add([sa.query.weight, sa.key.weight, sa.value.weight], "qkv_w", lyr.attention)
            add([sa.query.bias, sa.key.bias, sa.value.bias], "qkv_b", lyr.attention)
            add(list(lyr.parameters()))

        def add_stack(stack, layers=None):
            for lyr in reversed(list(stack.encoder.layer) if layers is None else layers):
                add_layer(lyr)

        marks = []                           # entry counts at the bucket boundaries
        word = m.bert.embeddings.word_embeddings.weight
        placed.add(id(word))                 # finished last (classifier dE at the start, embedding scatter at the end)
        add([m.classifier.bias] + ([] if tied else [m.classifier.weight]))
        add_stack(m.output_block)
        add(list(m.output_block.embeddings.parameters()))
        if hasattr(m, "gate_net"):
            add(list(m.gate_net.parameters()))
        if hasattr(m, "pho_model"):
            add_stack(m.pho_model)
            add(list(m.pho_model.embeddings.parameters()) + list(m.pho_gru.parameters()) + list(m.pho_embeddings.parameters()))
        if hasattr(m, "resnet"):
            add(list(m.resnet_layernorm.parameters()))
            for b in range(5, 0, -1):
                add(list(getattr(m.resnet, f"res_block{b}").parameters()))
        marks.append(len(entries))
        bert_layers = list(m.bert.encoder.layer)
        half = len(bert_layers) // 2
        add_stack(m.bert, bert_layers[half:])
        marks.append(len(entries))
        add_stack(m.bert, bert_layers[:half])
        placed.discard(id(word))
        add(list(m.bert.embeddings.parameters()))
        add([p for p in self.params])        # anything not named above (none for the shipped classes)
        marks.append(len(entries))

        def padded(n):  # every block starts 256-byte aligned (float4 atomics, TMA stores, vector loads)
            return (n + 63) // 64 * 64

        total = sum(padded(sum(p.numel() for p in ps)) for _, _, ps in entries)
        self.flat = torch.zeros(total, device=dev, dtype=F32)
        self.grads = [None] * len(self.params)
        off, bounds = 0, [0]
        for ei, (kind, att, ps) in enumerate(entries):
            start = off = padded(off)
            for p in ps:
                self.grads[self.index[id(p)]] = self.flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            if kind == "qkv_w":
                self._fused.setdefault(id(att), {})["w"] = self.flat[start:off].view(3 * H, H)
            elif kind == "qkv_b":
                self._fused.setdefault(id(att), {})["b"] = self.flat[start:off]
            if ei + 1 in marks:
                bounds.append(padded(off))
        bounds[-1] = total
        self.buckets = list(zip(bounds[:-1], bounds[1:]))        # [(start, end)] element offsets; may contain empty ranges
        self.bert_split = half               # bucket 1 is complete once bert layer `half` has been differentiated
        self.tied = tied

    def run(self, inputs):
        return _StepFn.apply(self, inputs, *self.params)

    def backward_and_sync(self, gloss):
        """Backward of the saved forward, the data-parallel gradient exchange, and .grad attachment.  Called by
        autograd (loss.backward()) or directly by realise_b200.graphed.GraphedTrainStep."""
        # gradient accumulation (src/run.py:193-205): the kernels write / atomically add into a freshly zeroed flat
        # buffer, so gradients that an earlier backward left for the optimizer are set aside and added back afterwards
        accumulate = self._pending and any(p.grad is not None for p in self.params)
        if accumulate:
            if self._accum is None:
                self._accum = torch.empty_like(self.flat)
            self._accum.copy_(self.flat)
        self.backward(gloss)
        hook = getattr(self.m, "_post_backward", None)   # realise_b200.ddp.DataParallel: all-reduce of self.flat (or the
        if hook is not None:                             # join of the bucketed all-reduces issued during the backward)
            hook(self)
        if accumulate:
            self.flat.add_(self._accum)
        self._pending = True
        # gradients live in persistent buffers (stable pointers for the fused optimizer); they are attached to the
        # parameters directly instead of being handed to autograd, which would copy or alias them
        for p, g in zip(self.params, self.grads):
            if g is not None:
                p.grad = g

    def _side_streams(self):
        if not self.multistream:
            return None, None
        if self._streams is None:
            dev = self.m.classifier.bias.device
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        main = torch.cuda.current_stream()
        for st in self._streams:
            st.wait_stream(main)          # fork: everything issued so far (inputs, zeroed scratch) is visible to the branches
        return self._streams

    def _join(self, *streams):
        main = torch.cuda.current_stream()
        for st in streams:
            if st is not None:
                main.wait_stream(st)

    @staticmethod
    def _on(stream):
        import contextlib
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def _bucket_done(self, k):
        """Every gradient of flat-buffer bucket k has been issued on the current stream: data parallelism may start
        reducing it while the rest of the backward runs (no-op without an overlapping DataParallel)."""
        hook = getattr(self.m, "_bucket_ready", None)
        if hook is not None and k < len(self.buckets):
            hook(self, k)

    def grads_consumed(self):
        """Called by FusedAdamW.step() and model.zero_grad(): the next backward starts from zero again."""
        self._pending = False

    def set_seed(self, seed):
        self.seed = int(seed)

    # dropout sites: stack s in {bert: 1, pho_model: 2, output_block: 3}; the masks are pure functions of
    # (seed, site, element index), regenerated by the backward kernels
    STACK_ID = {"bert": 1, "pho": 2, "out": 3}

    def site(self, stack, layer, k):
        """k: 1 attention probs, 2 attention-output dense, 3 FFN-output dense, 9 embeddings (layer ignored)."""
        return self.STACK_ID[stack] * 1000 + layer * 10 + k

    SITE_FINAL = 9999

    def hdrop(self, site):
        p = self.m.config.hidden_dropout_prob
        return (p, self.step_seed, site) if p > 0 else None

    def adrop(self, site):
        p = self.m.config.attention_probs_dropout_prob
        return (p, self.step_seed, site) if p > 0 else None

    # ---- helpers ---------------------------------------------------------------------------------
    def _new(self, shape, dtype, zero=False):
        if zero:
            assert dtype is F32
            return self.zpool.get(tuple(shape))
        return torch.empty(shape, device=self.m.classifier.bias.device, dtype=dtype)

    def _grad(self, p, zero=False):
        """Persistent fp32 gradient view of parameter p inside the flat buffer."""
        return self.grads[self.index[id(p)]]

    def _qkv_grads(self, att, H):
        """Fused [3H, H] weight / [3H] bias gradient of a layer's query/key/value (views of the flat buffer)."""
        f = self._fused[id(att)]
        return f["w"], f["b"]

    # ---- forward -----------------------------------------------------------------------------------
    def _stack_fwd(self, name, mod, P, mask, B, L, ids=None, inputs_embeds=None, pos_mode=0, last_drop=False):
        c = self.m.config
        N, H, I = B * L, c.hidden_size, c.intermediate_size
        sv = {"layers": [], "ids": ids, "pos_mode": pos_mode, "mod": mod, "P": P, "from_embeds": inputs_embeds is not None,
              "name": name, "last_drop": last_drop}
        A16 = self.act16
        x, xb = self._new((N, H), F32), self._new((N, H), A16)
        sv["e_pre"] = self._new((N, H), F32)
        ops.embed_ln(ids, P["word"], inputs_embeds, P["pos"], P["type0"], P["ln_w"], P["ln_b"], x, xb, N, L, H, pos_mode,
                     c.layer_norm_eps, pre_out=sv["e_pre"], drop=self.hdrop(self.site(name, 0, 9)))
        nl = len(P["layers"])
        for li, lw in enumerate(P["layers"]):
            s = {"xb": xb}
            s["qkv"] = self._new((N, 3 * H), A16)
            ops.gemm(xb, lw["w_qkv"], s["qkv"], bias=lw["b_qkv"])
            s["ctx"] = self._new((N, H), A16)
            s["lse"] = self._new((B * c.num_attention_heads * L,), F32)
            ops.attention(s["qkv"], mask, s["ctx"], B, L, c.num_attention_heads, drop=self.adrop(self.site(name, li, 1)),
                          lse=s["lse"])
            s["y1"] = self._new((N, H), F32)
            ops.gemm(s["ctx"], lw["w_o"], s["y1"], bias=lw["b_o"], res=x, drop=self.hdrop(self.site(name, li, 2)))
            x1, s["x1b"] = self._new((N, H), F32), self._new((N, H), A16)
            ops.layernorm(s["y1"], lw["ln1_w"], lw["ln1_b"], x1, s["x1b"], c.layer_norm_eps)
            s["u"], s["h"] = self._new((N, I), A16), self._new((N, I), A16)
            # h = gelu(u), u kept for the backward: both tiles leave the GEMM epilogue through TMA (packed f32x2 erf
            # polynomial on the 8 epilogue warps, hidden behind the next tile's MMAs)
            ops.gemm(s["x1b"], lw["w_1"], s["h"], bias=lw["b_1"], act=ops.ACT_GELU_SAVE, out2=s["u"])
            s["y2"] = self._new((N, H), F32)
            ops.gemm(s["h"], lw["w_2"], s["y2"], bias=lw["b_2"], res=x1, drop=self.hdrop(self.site(name, li, 3)))
            x, xb = self._new((N, H), F32), self._new((N, H), A16)
            # the classifier consumes dropout(sequence_output) (src/models.py:858): only the bf16 operand copy of
            # the last LayerNorm of output_block is masked
            fin = self.hdrop(self.SITE_FINAL) if (last_drop and li == nl - 1) else None
            ops.layernorm(s["y2"], lw["ln2_w"], lw["ln2_b"], x, xb, c.layer_norm_eps, drop=fin, drop_f32=False)
            sv["layers"].append(s)
        return x, xb, sv

    # ---- CharResNet (train): batch-statistics BatchNorm, every conv as a tcgen05 GEMM -------------------------
    RES_CH = [None, 64, 128, 256, 512, 768]

    @staticmethod
    def _s2_taps(S):
        """taps of a 3x3 stride-2 pad-1 conv over the parity-split input: (dw, dh, plane, kh, kw)."""
        taps = []
        for kh in range(3):
            for kw in range(3):
                ph, dh = (0, 0) if kh == 1 else (1, -1 if kh == 0 else 0)
                pw, dw = (0, 0) if kw == 1 else (1, -1 if kw == 0 else 0)
                if S == 1 and (dh < 0 or dw < 0):
                    continue
                taps.append((dw, dh, ph * 2 + pw, kh, kw))
        return taps

    @staticmethod
    def _s1_taps(S):
        return [(kw - 1, kh - 1, 0, kh, kw) for kh in range(3) for kw in range(3) if not (S == 1 and (kh != 1 or kw != 1))]

    def _conv_layouts(self, w1, w2, wsc, b, cin, cout, S, pad):
        """Operand layouts of one BasicBlock's conv weights as pure indexing of (w1 [cout,cin,3,3], w2 [cout,cout,3,3],
        wsc [cout,cin]) — run once on tensors of element CODES (pad = -1 marks structural zeros) to obtain gather maps."""
        def full(*shape):
            return torch.full(shape, pad, dtype=w1.dtype)
        e = {}
        t2 = self._s1_taps(S)
        e["w2f"] = torch.cat([w2[:, :, kh, kw] for (_, _, _, kh, kw) in t2], 1)
        # data gradient of conv2: da1[h, w] = sum_taps dc2[h - dh, w - dw] W2[:, :, kh, kw]^T
        e["w2t"] = torch.cat([w2[:, :, kh, kw].t() for (_, _, _, kh, kw) in t2], 1)
        if b == 1:
            w1g = full(cout, 32)
            w1g[:, :9 * cin] = w1.reshape(cout, 9 * cin)
            # the 1x1 stride-2 shortcut samples the centre tap of conv1's patch: same im2col matrix, weights on
            # columns c*9 + 4 only (a K = 8 operand would make TMA fetch 16-byte rows)
            wsg = full(cout, 32)
            wsg[:, 4:9 * cin:9] = wsc
            e["w1g"], e["wscg"] = w1g, wsg
        else:
            t1 = self._s2_taps(S)
            e["w1f"] = torch.cat([w1[:, :, kh, kw] for (_, _, _, kh, kw) in t1], 1)
            e["wscf"] = wsc.clone()
            if S == 1:
                # dx_in [N, 4*cin] = [dc1 | dcs] . Wbig^T, Wbig[p*cin+ci] = [W1[:, ci, 1+ph, 1+pw] | (p == 0) Wsc[:, ci]]
                rows = []
                for pl in range(4):
                    a = w1[:, :, 1 + pl // 2, 1 + pl % 2].t()
                    bsc = wsc.t() if pl == 0 else full(cin, cout)
                    rows.append(torch.cat([a, bsc], 1))
                e["wbig"] = torch.cat(rows, 0)                                        # [4*cin, 2*cout]
            else:
                for pl in range(4):
                    ph, pw = pl // 2, pl % 2
                    khs, kws = ([1] if ph == 0 else [0, 2]), ([1] if pw == 0 else [0, 2])
                    cols = [w1[:, :, kh, kw].t() for kh in khs for kw in kws]         # [cin, cout] each
                    if pl == 0:
                        cols.append(wsc.t())                                          # the shortcut rides on tap (0, 0)
                    e[f"dgrad{pl}"] = torch.cat(cols, 1)
        return e

    def _resnet_setup(self):
        """Built once per engine: per block the static geometry (taps), persistent bf16 operand buffers in every layout
        the forward / backward GEMMs read, ONE MultiGather that refreshes them all from the fp32 master weights, the
        persistent tap-major weight-gradient accumulators and ONE MultiGather that scatters them into the parameters'
        [cout, cin, kh, kw] gradients."""
        m, c = self.m, self.m.config
        dev = m.classifier.bias.device
        wg, gg = MultiGather(dev), MultiGather(dev)
        blocks, wtmp_sizes = [], []
        cin = c.num_fonts
        for b in range(1, 6):
            blk = getattr(m.resnet, f"res_block{b}")
            conv1, bn1, _, conv2, bn2 = blk.residual_function
            convs, bns = blk.shortcut
            cout, S = self.RES_CH[b], 32 >> b
            e = {"S": S, "cin": cin, "cout": cout, "conv1": conv1, "conv2": conv2, "convs": convs, "bn1": bn1, "bn2": bn2,
                 "bns": bns, "taps2": self._s1_taps(S)}
            e["taps2t"] = [(-dw, -dh, 0) for (dw, dh, _, _, _) in e["taps2"]]
            if b > 1:
                e["taps1"] = self._s2_taps(S)
                if S > 1:
                    e["dgrad_taps"] = []
                    for pl in range(4):
                        ph, pw = pl // 2, pl % 2
                        khs, kws = ([1] if ph == 0 else [0, 2]), ([1] if pw == 0 else [0, 2])
                        e["dgrad_taps"].append([(-(-1 if kw == 0 else 0), -(-1 if kh == 0 else 0), 0) for kh in khs for kw in kws])
            for conv in (conv1, conv2, convs):
                assert conv.weight.is_contiguous() and conv.weight.dtype is F32
            codes = self._conv_layouts(wg.source(conv1.weight.detach()), wg.source(conv2.weight.detach()),
                                       wg.source(convs.weight.detach()).view(cout, cin), b, cin, cout, S, pad=-1)
            for k, code in codes.items():
                e[k] = torch.empty(code.shape, device=dev, dtype=BF16)
                wg.add(e[k], code)
            if b > 1 and S > 1:
                e["dgrad"] = [(e["dgrad_taps"][pl], e[f"dgrad{pl}"]) for pl in range(4)]
            # tap-major weight-gradient accumulators (f32, split-K adds into them) and their scatter into .grad layout
            def wtmp(rows, cols):
                wtmp_sizes.append((rows, cols))
                return len(wtmp_sizes) - 1
            e["tmp2"] = wtmp(cout, len(e["taps2"]) * cout)
            e["tmp1"] = wtmp(2 * cout, 32) if b == 1 else wtmp(cout, len(e["taps1"]) * cin)
            blocks.append(e)
            cin = cout
        total = sum((r * k + 63) // 64 * 64 for r, k in wtmp_sizes)
        self._wtmp = torch.zeros(total, device=dev, dtype=F32)
        views, off = [], 0
        for r, k in wtmp_sizes:
            views.append(self._wtmp[off:off + r * k].view(r, k))
            off += (r * k + 63) // 64 * 64
        for e in blocks:
            cin, cout = e["cin"], e["cout"]
            t2v, t1v = views[e["tmp2"]], views[e["tmp1"]]
            e["tmp2"], e["tmp1"] = t2v, t1v
            # conv2: tmp [cout, T2*cout] tap-major -> grad [cout, cout, 3, 3]
            code2 = gg.source(t2v).view(cout, len(e["taps2"]), cout).permute(0, 2, 1)            # [co, ci, t]
            m2 = torch.full((cout, cout, 9), -1, dtype=torch.int32)
            m2[:, :, [kh * 3 + kw for (_, _, _, kh, kw) in e["taps2"]]] = code2
            gg.add(self._grad(e["conv2"].weight), m2)
            code1 = gg.source(t1v)
            if e["S"] == 16:      # block 1: rows 0..cout-1 = dW1 over the 32-wide patch, rows cout.. = dWsc on the centre taps
                gg.add(self._grad(e["conv1"].weight), code1[:cout, :9 * cin])
                gg.add(self._grad(e["convs"].weight), code1[cout:, 4:9 * cin:9])
            else:
                t1 = e["taps1"]
                c1 = code1.view(cout, len(t1), cin).permute(0, 2, 1)
                m1 = torch.full((cout, cin, 9), -1, dtype=torch.int32)
                m1[:, :, [kh * 3 + kw for (_, _, _, kh, kw) in t1]] = c1
                gg.add(self._grad(e["conv1"].weight), m1)
        self._res = {"blocks": blocks, "wgather": wg, "ggather": gg}

    def _resnet_weights(self):
        """bf16 operand layouts of the CURRENT conv weights: one gather launch over every layout of every conv."""
        if self._res is None:
            self._resnet_setup()
        self._res["wgather"].run()
        return self._res["blocks"]

    def _bn_train(self, raw, bn, M):
        """Batch statistics of a raw conv output -> (scale, shift, mean, rstd); updates the running stats.
        (The GEMM epilogue can produce the sums itself — rl_gemm_desc.colsum / colsumsq — but on these shapes the
        butterfly reductions cost the 64..256-column conv GEMMs more than this HBM-bound pass: measured 43.3 vs 42.0 ms/step.)"""
        C = raw.shape[1]
        sums = self._new((2 * C,), F32, zero=True)
        ops.bn_stats(raw, sums)
        sc, sh, mu, rs = (self._new((C,), F32) for _ in range(4))
        ops.bn_finalize(sums, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                        bn.num_batches_tracked, sc, sh, mu, rs, M)
        return sc, sh, mu, rs

    def _resnet_fwd(self, P, ids_flat, N, sv):
        c = self.m.config
        Ws = self._resnet_weights()
        glyphs = P["res"]["glyphs"]
        blocks = []
        x = None
        for bi, e in enumerate(Ws):
            S, cin, cout = e["S"], e["cin"], e["cout"]
            M = N * S * S
            s = {"e": e, "x_in": x}
            # raw (pre-BatchNorm) conv outputs: f32, except in res_block1-2 (16x16 / 8x8 maps: 97 % of the BatchNorm
            # bytes), where bf16 halves every BN pass; their statistics average >= 10^4 pixels per channel and the
            # normalised activation is rounded to bf16 right afterwards anyway
            raw_dt = BF16 if (self.raw_bf16 and S >= 8) else F32
            c1, cs = self._new((M, cout), raw_dt), self._new((M, cout), raw_dt)
            if bi == 0:
                s["col1"] = self._new((M, 32), BF16)
                ops.glyph_im2col(glyphs, ids_flat, s["col1"], None, N, c.num_fonts)
                ops.gemm(s["col1"], e["w1g"], c1)
                ops.gemm(s["col1"], e["wscg"], cs)
            elif S == 1:
                xin = x.view(N, 4 * cin)
                ops.gemm(xin, e["w1f"], c1)
                ops.gemm(xin[:, :cin], e["wscf"], cs)
            else:
                xin = x.view(N, 4, S, S, cin)
                ops.conv_gemm(xin, e["w1f"], c1, nimg=N, H=S, W=S, planes=4, taps=[t[:3] for t in e["taps1"]])
                ops.conv_gemm(xin, e["wscf"], cs, nimg=N, H=S, W=S, planes=4, taps=[(0, 0, 0)])
            s["bn1"] = self._bn_train(c1, e["bn1"], M)
            a1 = self._new((M, cout), BF16)
            ops.bn_apply(c1, s["bn1"][0], s["bn1"][1], None, None, None, a1, relu=True)
            c2 = self._new((M, cout), raw_dt)
            if S == 1:
                ops.gemm(a1, e["w2f"], c2)
            else:
                ops.conv_gemm(a1.view(N, 1, S, S, cout), e["w2f"], c2, nimg=N, H=S, W=S, planes=1,
                              taps=[t[:3] for t in e["taps2"]])
            s["bn2"] = self._bn_train(c2, e["bn2"], M)
            s["bns"] = self._bn_train(cs, e["bns"], M)
            last = bi == 4
            out = self._new((M, cout), F32 if last else BF16)
            ops.bn_apply(c2, s["bn2"][0], s["bn2"][1], cs, s["bns"][0], s["bns"][1], out, relu=True, remap=S >= 2, map_hw=(S, S))
            s.update({"c1": c1, "cs": cs, "a1": a1, "c2": c2, "out": out})
            blocks.append(s)
            x = out
        sv["res"] = blocks
        return x   # f32 [N, 768]

    def _conv_wgrad(self, dy, col, tmp, taps, conv=None):
        """tmp[co, t*cin + ci] += dy^T col (split-K GEMM, tap-major accumulator; scattered into the parameter's
        [co, ci, kh, kw] gradient by the engine's gradient MultiGather at the end of the CNN backward).
        conv = (x, nimg, S, planes): the im2col matrix is implicit (gathered by TMA inside the GEMM), col is None."""
        if conv is not None:
            x, nimg, S, planes = conv
            ops.conv_wgrad(dy, x, tmp, nimg=nimg, H=S, W=S, planes=planes, taps=[t[:3] for t in taps])
        else:
            ops.gemm(dy, col, tmp, a_t=True, b_t=True, split_k=-1)

    def _resnet_bwd(self, sv, dout, N):
        self._wtmp.zero_()                                   # every tap-major accumulator at once
        for bi in range(4, -1, -1):
            s = sv["res"][bi]
            if self.debug is not None:
                self.debug[f"dout{bi + 1}"] = dout.clone()
                self.debug[f"out{bi + 1}"] = s["out"].clone()
                self.debug[f"a1_{bi + 1}"] = s["a1"].clone()
            e = s["e"]
            S, cin, cout = e["S"], e["cin"], e["cout"]
            M = N * S * S
            remap = S >= 2
            g = self._grad
            dcat = self._new((M, 2 * cout), BF16)
            dc2 = self._new((M, cout), BF16)
            # out = relu(bn2(c2) + bns(cs))
            ops.bn_bwd2(dout, s["out"],
                        (s["c2"], s["bn2"][2], s["bn2"][3], e["bn2"].weight.detach(), g(e["bn2"].bias), g(e["bn2"].weight), dc2),
                        (s["cs"], s["bns"][2], s["bns"][3], e["bns"].weight.detach(), g(e["bns"].bias), g(e["bns"].weight),
                         dcat[:, cout:]), M, cout, remap=remap, map_hw=(S, S),
                        fwd=(s["bn2"][:2], s["bns"][:2]) if self.bn_mask_recompute else None)
            # conv2 weight gradient (reference layout [cout, cin, kh, kw]) and data gradient
            T2 = len(e["taps2"])
            if T2 == 9:
                self._conv_wgrad(dc2, None, e["tmp2"], e["taps2"], conv=(s["a1"], N, S, 1))
                da1 = self._new((M, cout), BF16)
                ops.conv_gemm(dc2.view(N, 1, S, S, cout), e["w2t"], da1, nimg=N, H=S, W=S, planes=1, taps=e["taps2t"])
            else:  # 1x1 map: only the centre tap touched data
                self._conv_wgrad(dc2, s["a1"], e["tmp2"], e["taps2"])
                da1 = self._new((M, cout), BF16)
                ops.gemm(dc2, e["w2f"], da1, b_t=True)
            # a1 = relu(bn1(c1))
            ops.bn_bwd(da1, s["a1"], s["c1"], s["bn1"][2], s["bn1"][3], e["bn1"].weight.detach(), g(e["bn1"].bias),
                       g(e["bn1"].weight), dcat[:, :cout], fwd=s["bn1"][:2] if self.bn_mask_recompute else None)
            dc1, dcs = dcat[:, :cout], dcat[:, cout:]
            gws = g(e["convs"].weight)
            if bi == 0:
                # [dc1 | dcs]^T col1 in ONE split-K GEMM (a full 128-row tile): rows 0..63 = dW1, rows 64..127 = dWsc
                ops.gemm(dcat, s["col1"], e["tmp1"], a_t=True, b_t=True, split_k=-1)
                break
            x_in = s["x_in"]
            t1 = e["taps1"]
            if S >= 2:   # parity-split block input gathered inside the GEMM (implicit im2col)
                self._conv_wgrad(dc1, None, e["tmp1"], t1, conv=(x_in, N, S, 4))
                ops.conv_wgrad(dcs, x_in, gws.view(cout, cin), nimg=N, H=S, W=S, planes=4, taps=[(0, 0, 0)])
            else:        # 1x1 map: the 4 parity planes of the input are 4 column blocks of one row
                col = self._new((M, cin * len(t1)), BF16)
                ops.im2col(x_in, col, N, cin, S, S, 4, [t[:3] for t in t1])
                self._conv_wgrad(dc1, col, e["tmp1"], t1)
                colsc = self._new((M, cin), BF16)
                ops.im2col(x_in, colsc, N, cin, S, S, 4, [(0, 0, 0)])
                ops.gemm(dcs, colsc, gws.view(cout, cin), a_t=True, b_t=True, split_k=-1)
            # data gradient wrt the block input (parity-split rows = the previous block's output layout)
            dx = self._new((N * 4 * S * S, cin), BF16)
            if S == 1:
                ops.gemm(dcat, e["wbig"], dx.view(N, 4 * cin))
            else:
                dview = dcat.view(N, 1, S, S, 2 * cout)
                for pl, (taps, wp) in enumerate(e["dgrad"]):
                    cu = 2 * cout if pl == 0 else cout
                    ops.conv_gemm(dview, wp, dx, nimg=N, H=S, W=S, planes=1, taps=taps, out_remap=2, remap_plane=pl,
                                  c_use=cu)
            dout = dx
        self._res["ggather"].run()                           # tap-major accumulators -> [cout, cin, kh, kw] gradients

    # ---- pinyin GRU (train): every step's state is kept for the backward-through-time ---------------
    def _gru_fwd(self, P, pho_idx, lens_dev, N, sv):
        H = self.m.config.hidden_size
        T = pho_idx.shape[1]
        G = P["gru"]
        m = self.m
        # the input-projection table depends on trainable weights: rebuild it every training step
        ops.gru_input_table(m.pho_embeddings.weight.detach(), m.pho_gru.weight_ih_l0.detach(), m.pho_gru.bias_ih_l0.detach(),
                            G["table"])
        hs, hbs, ghs = [], [], [None]
        h, hb = self._new((N, H), F32), self._new((N, H), self.act16)
        ops.gru_step(None, G["b_hh"], G["table"], pho_idx, lens_dev, None, h, hb, 0)
        hs.append(h)
        hbs.append(hb)
        for t in range(1, T):
            gh = self._new((N, 3 * H), F32)
            ops.gemm(hbs[-1], G["w_hh"], gh, bias=G["b_hh"])
            h, hb = self._new((N, H), F32), self._new((N, H), self.act16)
            ops.gru_step(gh, G["b_hh"], G["table"], pho_idx, lens_dev, hs[-1], h, hb, t)
            hs.append(h)
            hbs.append(hb)
            ghs.append(gh)
        sv["gru"] = {"hs": hs, "hbs": hbs, "ghs": ghs, "pho_idx": pho_idx, "lens": lens_dev, "T": T}
        return hs[-1]

    def _gru_bwd(self, P, sv, dh, N):
        m = self.m
        H = m.config.hidden_size
        g, G = sv["gru"], P["gru"]
        gru = m.pho_gru
        dW_hh, db_hh = self._grad(gru.weight_hh_l0), self._grad(gru.bias_hh_l0)
        dtable = self._new((64, 3 * H), F32, zero=True)
        for t in range(g["T"] - 1, -1, -1):
            dh_prev = self._new((N, H), F32)
            dgi, dgh = self._new((N, 3 * H), BF16), self._new((N, 3 * H), BF16)
            onehot = self._new((N, 64), BF16)
            ops.gru_step_bwd(dh, g["ghs"][t], G["b_hh"], G["table"], g["pho_idx"], g["lens"],
                             g["hs"][t - 1] if t > 0 else None, dh_prev, dgi, dgh, onehot, t)
            ops.colsum_bf16(dgh, db_hh)
            ops.gemm(onehot, dgi, dtable, a_t=True, b_t=True, split_k=-1)                 # dtable += onehot^T dgi
            if t > 0:
                ops.gemm(dgh, g["hbs"][t - 1], dW_hh, a_t=True, b_t=True, split_k=-1)      # dW_hh += dgh^T h_{t-1}
                dh_next = self._new((N, H), F32)
                ops.gemm(dgh, G["w_hh"], dh_next, b_t=True, res=dh_prev)                 # dh_{t-1} = dh z + dgh W_hh
                dh = dh_next
        ops.gru_table_bwd(dtable, m.pho_embeddings.weight.detach(), gru.weight_ih_l0.detach(),
                          self._grad(gru.weight_ih_l0), self._grad(gru.bias_ih_l0), self._grad(m.pho_embeddings.weight))

    def forward(self, inp):
        m, c = self.m, self.m.config
        if m._prepared is None:
            m.prepare()
        P = m._prepared
        input_ids, mask = inp["src_idx"], inp["masks"]
        B, L = input_ids.shape
        N, H, V = B * L, c.hidden_size, c.vocab_size
        if L > 256:
            raise NotImplementedError("the attention kernels support seq_len <= 256")
        self.act16 = P["half"]       # 16-bit format of activations and gradients (bf16 in training)
        self.zpool.begin((B, L, inp["pho_idx"].shape[1] if "pho_idx" in inp else 0))
        self.step_seed = self.seed
        self.seed = (self.seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        sv = {"B": B, "L": L, "mask": mask, "inp": inp, "seed": self.step_seed}
        # the three encoders are independent until the fusion: issue them on three streams (fork / join; inside a captured
        # graph these become parallel branches)
        s_pho, s_res = self._side_streams() if (c.with_pho == "yes" or c.with_res == "yes") else (None, None)
        if c.with_res == "yes":     # issued first: the longest HBM-bound chain
            with self._on(s_res):
                res_raw = self._resnet_fwd(P, input_ids.view(-1), N, sv)
                sv["res_raw"] = res_raw
                res_h = self._new((N, H), F32)
                ops.layernorm(res_raw, P["res_ln_w"], P["res_ln_b"], res_h, None, c.layer_norm_eps)
        if c.with_pho == "yes":
            with self._on(s_pho):
                pho_gru = self._gru_fwd(P, inp["pho_idx"], inp["pho_lens"], N, sv)
                pho_h, _, sv["pho"] = self._stack_fwd("pho", m.pho_model, P["pho_model"], mask, B, L, inputs_embeds=pho_gru)
        bert_h, _, sv["bert"] = self._stack_fwd("bert", m.bert, P["bert"], mask, B, L, ids=input_ids.view(-1))
        self._join(s_pho, s_res)
        mods = [bert_h]
        if c.with_pho == "yes":
            mods.append(pho_h)
        if c.with_res == "yes":
            mods.append(res_h)
        sv["mods"] = mods
        fused = self._new((N, H), F32)
        if c.fusion == "gate":
            sv["gates"] = self._new((N, 3), F32)
            ops.gate_fuse(mods, False, mask, P["gate_w"], P["gate_b"], self._new((B * 3,), F32), fused, sv["gates"], B, L, H)
        else:
            ops.gate_fuse(mods, True, None, None, None, None, fused, None, B, L, H)
        _, seq_b, sv["out"] = self._stack_fwd("out", m.output_block, P["output_block"], mask, B, L, inputs_embeds=fused,
                                              pos_mode=1, last_drop=True)
        sv["seq_b"] = seq_b
        # train-mode logits live in fp16 (independent of the bf16 operand format: the epilogue packs what it is told):
        # half the bytes written here and re-read by the CE forward and backward; the reference's loop only reads
        # outputs[0] in training (src/run.py:191), callers that want them get an fp16 [B, L, V] tensor
        logits = self._new((N, V), torch.float16)
        ops.gemm(seq_b, P["cls_w"], logits, bias=P["cls_b"])
        loss = self._new((1,), F32)
        sv["lse"], sv["count"] = self._new((N,), F32), self._new((1,), F32)
        ops.masked_ce(logits, inp["tgt_idx"].view(-1), inp["loss_masks"].view(-1), self._new((N,), F32), loss,
                      row_lse=sv["lse"], count=sv["count"])
        sv["logits"] = logits
        self.saved = sv
        return loss[0], logits.view(B, L, V)

    # ---- backward ----------------------------------------------------------------------------------
    def _stack_bwd(self, sv, dx, mask, B, L):
        """dx: f32 [N,H] gradient of the stack output.  Returns the gradient wrt inputs_embeds (or None)."""
        c = self.m.config
        N, H, I = B * L, c.hidden_size, c.intermediate_size
        mod, P, name = sv["mod"], sv["P"], sv["name"]
        hp, seed = c.hidden_dropout_prob, self.step_seed
        nl = len(sv["layers"])
        for li in range(nl - 1, -1, -1):
            s, lw, lyr = sv["layers"][li], P["layers"][li], mod.encoder.layer[li]
            att, out = lyr.attention, lyr.output
            # x = LN2(y2),  y2 = h W2^T + b2 + x1
            dy2, dy2b = self._new((N, H), F32), self._new((N, H), BF16)
            # dx arrives through dropout only for the last LayerNorm of output_block (final dropout, site_in)
            ops.layernorm_bwd(dx, s["y2"], lw["ln2_w"], None, dy2, dy2b, self._grad(out.LayerNorm.weight, True),
                              self._grad(out.LayerNorm.bias, True), self._grad(out.dense.bias, True), c.layer_norm_eps,
                              drop_p=hp, drop_seed=seed,
                              site_in=self.SITE_FINAL if (hp > 0 and sv["last_drop"] and li == nl - 1) else 0,
                              site_out=self.site(name, li, 3) if hp > 0 else 0)
            ops.gemm(dy2b, s["h"], self._grad(out.dense.weight), a_t=True, b_t=True, split_k=-1)              # dW2 = dy2^T h
            du = self._new((N, I), BF16)
            if self.fused_gelu_grad:
                ops.gemm(dy2b, lw["w_2"], du, b_t=True, res=s["u"], act=ops.ACT_GELU_GRAD,
                         colsum=self._grad(lyr.intermediate.dense.bias, True))
            else:
                ops.gemm(dy2b, lw["w_2"], du, b_t=True)
                # du = (dy2 W2) o gelu'(u) and db1 as one element-wise pass
                ops.gelu_bwd_colsum(du, s["u"], self._grad(lyr.intermediate.dense.bias, True))
            ops.gemm(du, s["x1b"], self._grad(lyr.intermediate.dense.weight), a_t=True, b_t=True, split_k=-1)  # dW1 = du^T x1
            dx1 = self._new((N, H), F32)
            ops.gemm(du, lw["w_1"], dx1, b_t=True, res=dy2)                                         # dx1 = du W1 + dy2
            # x1 = LN1(y1),  y1 = ctx Wo^T + bo + x
            dy1, dy1b = self._new((N, H), F32), self._new((N, H), BF16)
            ops.layernorm_bwd(dx1, s["y1"], lw["ln1_w"], None, dy1, dy1b, self._grad(att.output.LayerNorm.weight, True),
                              self._grad(att.output.LayerNorm.bias, True), self._grad(att.output.dense.bias, True),
                              c.layer_norm_eps, drop_p=hp, drop_seed=seed, site_out=self.site(name, li, 2) if hp > 0 else 0)
            ops.gemm(dy1b, s["ctx"], self._grad(att.output.dense.weight), a_t=True, b_t=True, split_k=-1)      # dWo = dy1^T ctx
            dctx = self._new((N, H), BF16)
            ops.gemm(dy1b, lw["w_o"], dctx, b_t=True)
            dqkv = self._new((N, 3 * H), BF16)
            dw, db = self._qkv_grads(att, H)
            fused_db = L <= 128      # the q/k/v bias gradient comes out of the attention backward's epilogue
            ops.attention_bwd(s["qkv"], mask, s["ctx"], dctx, dqkv, B, L, c.num_attention_heads,
                              drop=self.adrop(self.site(name, li, 1)), lse=s["lse"], dbias=db if fused_db else None)
            if not fused_db:
                ops.colsum_bf16(dqkv, db)
            ops.gemm(dqkv, s["xb"], dw, a_t=True, b_t=True, split_k=-1)                                         # dWqkv = dqkv^T x
            dxin = self._new((N, H), F32)
            ops.gemm(dqkv, lw["w_qkv"], dxin, b_t=True, res=dy1)                                    # dx = dqkv Wqkv + dy1
            dx = dxin
            if name == "bert" and li == self.bert_split:
                self._bucket_done(1)                   # the upper half of the bert layers is complete
        # embeddings: x0 = LN(e),  e = word[ids] (or inputs_embeds) + pos + type0
        e = mod.embeddings
        de = self._new((N, H), F32)
        # every token has type 0 (modeling_bert.py:183): the column sum of de IS row 0 of the token-type gradient
        gtype = self._grad(e.token_type_embeddings.weight, True)
        ops.layernorm_bwd(dx, sv["e_pre"], P["ln_w"], None, de, None, self._grad(e.LayerNorm.weight, True),
                          self._grad(e.LayerNorm.bias, True), gtype[0], c.layer_norm_eps, drop_p=hp, drop_seed=seed,
                          site_in=self.site(name, 0, 9) if hp > 0 else 0)
        dpos = self._grad(e.position_embeddings.weight, True)
        if sv["from_embeds"]:
            ops.embed_bwd(de, None, None, dpos, N, L, H, sv["pos_mode"])
            return de
        ops.embed_bwd(de, sv["ids"], self._grad(e.word_embeddings.weight, True), dpos, N, L, H, sv["pos_mode"])
        return None

    def backward(self, gloss):
        m, c, sv = self.m, self.m.config, self.saved
        P = m._prepared
        B, L, mask = sv["B"], sv["L"], sv["mask"]
        N, H, V = B * L, c.hidden_size, c.vocab_size
        self.step_seed = sv["seed"]
        self.flat.zero_()   # every gradient is accumulated into (split-K GEMMs use f32 atomics)
        inp = sv["inp"]
        # classifier + masked CE:  logits = seq E^T + b
        dlogits = self._new((N, V), BF16)
        gscale = gloss.reshape(1).to(F32).contiguous()
        ops.masked_ce_bwd(sv["logits"], inp["tgt_idx"].view(-1), inp["loss_masks"].view(-1), sv["lse"], sv["count"], gscale,
                          dlogits)
        ops.colsum_bf16(dlogits, self._grad(m.classifier.bias, True))
        gE = self._grad(m.classifier.weight, zero=False)
        ops.gemm(dlogits, sv["seq_b"], gE, a_t=True, b_t=True, split_k=-1)                                      # dE = dlogits^T seq
        dseq = self._new((N, H), F32)
        ops.gemm(dlogits, P["cls_w"], dseq, b_t=True)                                               # dseq = dlogits E
        dfused = self._stack_bwd(sv["out"], dseq, mask, B, L)
        # gated fusion
        nm = len(sv["mods"])
        if c.fusion == "gate":
            dms = [self._new((N, H), F32) for _ in range(nm)]
            ws = self._new((N * 3 + 2 * B * H,), F32)
            ops.gate_fuse_bwd(dfused, sv["mods"], mask, sv["gates"], P["gate_w"], dms, self._grad(m.gate_net.weight, True),
                              self._grad(m.gate_net.bias, True), ws, B, L, H)
        else:
            dms = [dfused] * nm
        dm0 = dms[0]
        s_pho, s_res = self._side_streams() if (c.with_pho == "yes" or c.with_res == "yes") else (None, None)
        if c.with_res == "yes":
            with self._on(s_res):
                dres = self._new((N, H), F32)
                ops.layernorm_bwd(dms[-1], sv["res_raw"], P["res_ln_w"], None, dres, None, self._grad(m.resnet_layernorm.weight),
                                  self._grad(m.resnet_layernorm.bias), None, c.layer_norm_eps)
                if self.debug is not None:
                    self.debug["dres"] = dres.clone()
                self._resnet_bwd(sv, dres, N)
        if c.with_pho == "yes":
            with self._on(s_pho):
                dgru = self._stack_bwd(sv["pho"], dms[1], mask, B, L)
                self._gru_bwd(P, sv, dgru, N)
        if getattr(m, "_bucket_ready", None) is not None:
            self._join(s_pho, s_res)
            self._bucket_done(0)                       # output_block, gate, pinyin and glyph gradients are complete
        self._stack_bwd(sv["bert"], dm0, mask, B, L)   # scatter-adds the embedding rows into the (tied) dE buffer
        self._bucket_done(2)
        self._join(s_pho, s_res)
        self.saved = None
        # parameters that never receive a gradient (poolers, unused word embeddings of output_block) -> None
        return self.grads
