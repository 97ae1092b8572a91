"""Training-mode forward + backward of the hot path on librealise_b200.so.

`TrainEngine` runs the train-mode forward (saving the activations the backward needs) and the backward pass
as a fixed sequence of C-ABI kernel calls; `torch.autograd` only sees ONE custom Function whose backward returns
the parameter gradients, so `loss.backward()` in the reference's loop (src/run.py:191-200) keeps working.

Round-1 coverage: the semantic path (BertModel stacks, gated fusion, tied classifier, masked CE).  The pinyin GRU
(BPTT) and CharResNet (conv / batch-stat BatchNorm backward) branches are not differentiated yet, so training is
available for `with_pho='no', with_res='no'` configurations (src/models_abla.py) and raises otherwise.  Dropout must
be 0 in this round (the Philox dropout kernels are not written yet): parity is checked against the oracle's
autograd with dropout off, the protocol of SURVEY.md §8c.
"""
import torch

from . import ops

F32, BF16 = torch.float32, torch.bfloat16


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole model: forward -> (loss, logits); backward -> every parameter gradient."""

    @staticmethod
    def forward(ctx, engine, inputs, *params):
        loss, logits = engine.forward(inputs)
        ctx.engine = engine
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        eng = ctx.engine
        eng.backward(gloss)
        # gradients live in persistent buffers (stable pointers for the fused optimizer); they are attached to the
        # parameters directly instead of being handed to autograd, which would copy or alias them
        for p, g in zip(eng.params, eng.grads):
            if g is not None:
                p.grad = g
        return (None, None) + (None,) * len(eng.params)


class TrainEngine:
    def __init__(self, model):
        self.m = model
        c = model.config
        if c.with_pho == "yes" or c.with_res == "yes":
            raise NotImplementedError("training kernels cover the semantic path only in this round: GRU BPTT and "
                                      "CharResNet/BatchNorm backward are not written yet (use with_pho='no', with_res='no')")
        if c.hidden_dropout_prob != 0.0 or c.attention_probs_dropout_prob != 0.0:
            raise NotImplementedError("train mode needs hidden_dropout_prob = attention_probs_dropout_prob = 0 in this "
                                      "round (Philox dropout kernels not written yet)")
        self.saved = None
        # trainable parameters in a fixed order
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.grads = [None] * len(self.params)   # persistent fp32 gradient buffers (allocated on first use)
        self.zero_list = []                       # buffers that kernels accumulate into: zeroed every step
        self._fused = {}

    def run(self, inputs):
        return _StepFn.apply(self, inputs, *self.params)

    # ---- helpers ---------------------------------------------------------------------------------
    def _new(self, shape, dtype, zero=False):
        dev = self.m.classifier.bias.device
        return torch.zeros(shape, device=dev, dtype=dtype) if zero else torch.empty(shape, device=dev, dtype=dtype)

    def _grad(self, p, zero=False):
        """Persistent fp32 gradient buffer of parameter p (registered for per-step zeroing when accumulated into)."""
        i = self.index[id(p)]
        if self.grads[i] is None:
            self.grads[i] = self._new(tuple(p.shape), F32, zero=zero)
            if zero:
                self.zero_list.append(self.grads[i])
        return self.grads[i]

    def _qkv_grads(self, att, H):
        """Fused [3H, H] weight / [3H] bias gradient of a layer's query/key/value; the parameters get row views."""
        key = id(att)
        if key not in self._fused:
            dw, db = self._new((3 * H, H), F32), self._new((3 * H,), F32, zero=True)
            self.zero_list.append(db)
            for k, lin in enumerate((att.self.query, att.self.key, att.self.value)):
                self.grads[self.index[id(lin.weight)]] = dw[k * H:(k + 1) * H]
                self.grads[self.index[id(lin.bias)]] = db[k * H:(k + 1) * H]
            self._fused[key] = (dw, db)
        return self._fused[key]

    # ---- forward -----------------------------------------------------------------------------------
    def _stack_fwd(self, name, mod, P, mask, B, L, ids=None, inputs_embeds=None, pos_mode=0):
        c = self.m.config
        N, H, I = B * L, c.hidden_size, c.intermediate_size
        sv = {"layers": [], "ids": ids, "pos_mode": pos_mode, "mod": mod, "P": P, "from_embeds": inputs_embeds is not None}
        x, xb = self._new((N, H), F32), self._new((N, H), BF16)
        sv["e_pre"] = self._new((N, H), F32)
        ops.embed_ln(ids, P["word"], inputs_embeds, P["pos"], P["type0"], P["ln_w"], P["ln_b"], x, xb, N, L, H, pos_mode,
                     c.layer_norm_eps, pre_out=sv["e_pre"])
        for lw in P["layers"]:
            s = {"xb": xb}
            s["qkv"] = self._new((N, 3 * H), BF16)
            ops.gemm(xb, lw["w_qkv"], s["qkv"], bias=lw["b_qkv"])
            s["ctx"] = self._new((N, H), BF16)
            ops.attention(s["qkv"], mask, s["ctx"], B, L, c.num_attention_heads)
            s["y1"] = self._new((N, H), F32)
            ops.gemm(s["ctx"], lw["w_o"], s["y1"], bias=lw["b_o"], res=x)
            x1, s["x1b"] = self._new((N, H), F32), self._new((N, H), BF16)
            ops.layernorm(s["y1"], lw["ln1_w"], lw["ln1_b"], x1, s["x1b"], c.layer_norm_eps)
            s["u"], s["h"] = self._new((N, I), BF16), self._new((N, I), BF16)
            ops.gemm(s["x1b"], lw["w_1"], s["h"], bias=lw["b_1"], act=ops.ACT_GELU_SAVE, out2=s["u"])
            s["y2"] = self._new((N, H), F32)
            ops.gemm(s["h"], lw["w_2"], s["y2"], bias=lw["b_2"], res=x1)
            x, xb = self._new((N, H), F32), self._new((N, H), BF16)
            ops.layernorm(s["y2"], lw["ln2_w"], lw["ln2_b"], x, xb, c.layer_norm_eps)
            sv["layers"].append(s)
        return x, xb, sv

    def forward(self, inp):
        m, c = self.m, self.m.config
        if m._prepared is None:
            m.prepare()
        P = m._prepared
        input_ids, mask = inp["src_idx"], inp["masks"]
        B, L = input_ids.shape
        N, H, V = B * L, c.hidden_size, c.vocab_size
        if L > 128:
            raise NotImplementedError("attention backward kernel supports seq_len <= 128")
        sv = {"B": B, "L": L, "mask": mask, "inp": inp}
        bert_h, _, sv["bert"] = self._stack_fwd("bert", m.bert, P["bert"], mask, B, L, ids=input_ids.view(-1))
        mods = [bert_h]
        sv["mods"] = mods
        fused = self._new((N, H), F32)
        if c.fusion == "gate":
            sv["gates"] = self._new((N, 3), F32)
            ops.gate_fuse(mods, False, mask, P["gate_w"], P["gate_b"], self._new((B * 3,), F32), fused, sv["gates"], B, L, H)
        else:
            ops.gate_fuse(mods, True, None, None, None, None, fused, None, B, L, H)
        _, seq_b, sv["out"] = self._stack_fwd("out", m.output_block, P["output_block"], mask, B, L, inputs_embeds=fused,
                                              pos_mode=1)
        sv["seq_b"] = seq_b
        logits = self._new((N, V), F32)
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
        mod, P = sv["mod"], sv["P"]
        for li in range(len(sv["layers"]) - 1, -1, -1):
            s, lw, lyr = sv["layers"][li], P["layers"][li], mod.encoder.layer[li]
            att, out = lyr.attention, lyr.output
            # x = LN2(y2),  y2 = h W2^T + b2 + x1
            dy2, dy2b = self._new((N, H), F32), self._new((N, H), BF16)
            ops.layernorm_bwd(dx, s["y2"], lw["ln2_w"], None, dy2, dy2b, self._grad(out.LayerNorm.weight, True),
                              self._grad(out.LayerNorm.bias, True), self._grad(out.dense.bias, True), c.layer_norm_eps)
            ops.gemm(dy2b, s["h"], self._grad(out.dense.weight), a_t=True, b_t=True)              # dW2 = dy2^T h
            du = self._new((N, I), BF16)
            ops.gemm(dy2b, lw["w_2"], du, b_t=True, res=s["u"], act=ops.ACT_GELU_GRAD)              # du = (dy2 W2) gelu'(u)
            ops.colsum_bf16(du, self._grad(lyr.intermediate.dense.bias, True))
            ops.gemm(du, s["x1b"], self._grad(lyr.intermediate.dense.weight), a_t=True, b_t=True)  # dW1 = du^T x1
            dx1 = self._new((N, H), F32)
            ops.gemm(du, lw["w_1"], dx1, b_t=True, res=dy2)                                         # dx1 = du W1 + dy2
            # x1 = LN1(y1),  y1 = ctx Wo^T + bo + x
            dy1, dy1b = self._new((N, H), F32), self._new((N, H), BF16)
            ops.layernorm_bwd(dx1, s["y1"], lw["ln1_w"], None, dy1, dy1b, self._grad(att.output.LayerNorm.weight, True),
                              self._grad(att.output.LayerNorm.bias, True), self._grad(att.output.dense.bias, True),
                              c.layer_norm_eps)
            ops.gemm(dy1b, s["ctx"], self._grad(att.output.dense.weight), a_t=True, b_t=True)      # dWo = dy1^T ctx
            dctx = self._new((N, H), BF16)
            ops.gemm(dy1b, lw["w_o"], dctx, b_t=True)
            dqkv = self._new((N, 3 * H), BF16)
            ops.attention_bwd(s["qkv"], mask, s["ctx"], dctx, dqkv, B, L, c.num_attention_heads)
            dw, db = self._qkv_grads(att, H)
            ops.colsum_bf16(dqkv, db)
            ops.gemm(dqkv, s["xb"], dw, a_t=True, b_t=True)                                         # dWqkv = dqkv^T x
            dxin = self._new((N, H), F32)
            ops.gemm(dqkv, lw["w_qkv"], dxin, b_t=True, res=dy1)                                    # dx = dqkv Wqkv + dy1
            dx = dxin
        # embeddings: x0 = LN(e),  e = word[ids] (or inputs_embeds) + pos + type0
        e = mod.embeddings
        de = self._new((N, H), F32)
        dtype_sum = self._new((H,), F32, zero=True)
        ops.layernorm_bwd(dx, sv["e_pre"], P["ln_w"], None, de, None, self._grad(e.LayerNorm.weight, True),
                          self._grad(e.LayerNorm.bias, True), dtype_sum, c.layer_norm_eps)
        gtype = self._grad(e.token_type_embeddings.weight, True)
        gtype[0].copy_(dtype_sum)
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
        if self.zero_list:
            torch._foreach_zero_(self.zero_list)
        inp = sv["inp"]
        # classifier + masked CE:  logits = seq E^T + b
        dlogits = self._new((N, V), BF16)
        gscale = gloss.reshape(1).to(F32).contiguous()
        ops.masked_ce_bwd(sv["logits"], inp["tgt_idx"].view(-1), inp["loss_masks"].view(-1), sv["lse"], sv["count"], gscale,
                          dlogits)
        ops.colsum_bf16(dlogits, self._grad(m.classifier.bias, True))
        tied = m.classifier.weight is m.bert.embeddings.word_embeddings.weight
        gE = self._grad(m.classifier.weight, zero=False)
        ops.gemm(dlogits, sv["seq_b"], gE, a_t=True, b_t=True)                                      # dE = dlogits^T seq
        dseq = self._new((N, H), F32)
        ops.gemm(dlogits, P["cls_w"], dseq, b_t=True)                                               # dseq = dlogits E
        dfused = self._stack_bwd(sv["out"], dseq, mask, B, L)
        # gated fusion
        dm0 = self._new((N, H), F32)
        if c.fusion == "gate":
            ws = self._new((N * 3 + 2 * B * H,), F32)
            ops.gate_fuse_bwd(dfused, sv["mods"], mask, sv["gates"], P["gate_w"], [dm0], self._grad(m.gate_net.weight, True),
                              self._grad(m.gate_net.bias, True), ws, B, L, H)
        else:
            dm0 = dfused
        if not tied:
            self._grad(m.bert.embeddings.word_embeddings.weight, True)
        self._stack_bwd(sv["bert"], dm0, mask, B, L)   # scatter-adds the embedding rows into the (tied) dE buffer
        self.saved = None
        # parameters that never receive a gradient (poolers, unused word embeddings of output_block) -> None
        return self.grads
