"""ORACLE — test infrastructure only.  Not part of the product path.

A CPU (fp32, plain torch arithmetic) restatement of the ReaLiSe multimodal forward
`SpellBertPho2ResArch3.forward` / `SpellBertPho2ResArch3Abla.forward`, written as a pure function of
a reference-format state_dict and a batch dict.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; realise_b200/ never does.

Parity pinning: tests/golden/*.npz hold outputs of the REFERENCE ITSELF (imported from
/root/reference in the build container by tests/golden/make_golden.py) on the synthetic weights /
batches of realise_b200/synth.py; tests/test_oracle.py checks this restatement against them.

Reference call sites restated here (paths relative to the reference repo):
  src/models.py:806-870                  forward orchestration, gated fusion, classifier, masked CE
  src/models_abla.py:202-299             modality switches / sum fusion
  transformers/modeling_bert.py:169-193  BertEmbeddings        -> embeddings()
  transformers/modeling_bert.py:220-263  BertSelfAttention     -> self_attention()
  transformers/modeling_bert.py:273-277  BertSelfOutput        -> bert_layer()
  transformers/modeling_bert.py:326-343  BertIntermediate/BertOutput (erf GELU :125-131)
  transformers/modeling_bert.py:687-697  additive mask (1-mask)*-10000
  src/models.py:818-826                  pho_embeddings + packed nn.GRU final hidden -> gru_final()
  src/char_cnn.py:9-55                   BasicBlock / CharResNet -> char_resnet()
"""
import math

import torch
import torch.nn.functional as F

RES_BLOCKS = 5

# FAST = False: every step spelled out in elementary tensor arithmetic (the restatement that is
# checked line by line against the reference).  FAST = True: the same math through ATen's fused CPU
# kernels (F.layer_norm, F.gelu, F.batch_norm, packed nn.GRU) — i.e. the kernels the reference's
# own nn.Modules dispatch to — used for the CPU timing baseline so the port is not slower than the
# reference it stands in for.  tests/test_oracle.py checks both modes against the goldens.
FAST = False


def layer_norm(x, w, b, eps):
    if FAST:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def gelu_erf(x):
    if FAST:
        return F.gelu(x)
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


# Dropout masks: torch's RNG stream cannot be reproduced by the CUDA kernels, so for train-mode parity the tests
# install MASK_FN(site_id, shape) -> keep mask (the kernels' own counter-based masks, exported with
# rl_dropout_mask) and both sides apply identical masks.  Site ids (realise_b200/train.py): stack * 1000 +
# layer * 10 + k, stack in {bert 1, pho_model 2, output_block 3}, k in {1 attention probs, 2 attention-output
# dense, 3 FFN-output dense, 9 embeddings}; 9999 = the dropout before the classifier.
MASK_FN = None
STACK_ID = {"bert": 1, "pho_model": 2, "output_block": 3}
SITE_FINAL = 9999


def site_id(path, k):
    parts = path.split(".")
    layer = int(parts[3]) if len(parts) > 3 else 0
    return STACK_ID[parts[0]] * 1000 + layer * 10 + k


def dropout(x, p, train, site=None, gen=None):
    if not train or p == 0.0:
        return x
    if MASK_FN is not None and site is not None:
        keep = MASK_FN(site, tuple(x.shape)).to(x.dtype)
    else:
        keep = (torch.rand(x.shape, generator=gen, device=x.device) >= p).to(x.dtype)
    return x * keep / (1.0 - p)


def embeddings(sd, prefix, cfg, input_ids=None, inputs_embeds=None, position_ids=None, train=False):
    """BertEmbeddings.forward: LN(x + pos + type0) -> dropout."""
    if inputs_embeds is None:
        inputs_embeds = sd[f"{prefix}.embeddings.word_embeddings.weight"][input_ids]
    B, L = inputs_embeds.shape[:2]
    if position_ids is None:
        position_ids = torch.arange(L, device=inputs_embeds.device).unsqueeze(0).expand(B, L)
    pos = sd[f"{prefix}.embeddings.position_embeddings.weight"][position_ids]
    typ = sd[f"{prefix}.embeddings.token_type_embeddings.weight"][torch.zeros_like(position_ids)]
    x = inputs_embeds + pos + typ
    x = layer_norm(x, sd[f"{prefix}.embeddings.LayerNorm.weight"], sd[f"{prefix}.embeddings.LayerNorm.bias"],
                   cfg.layer_norm_eps)
    return dropout(x, cfg.hidden_dropout_prob, train, site=site_id(prefix, 9))


def self_attention(sd, p, cfg, x, ext_mask, train=False):
    B, L, H = x.shape
    nh = cfg.num_attention_heads
    d = H // nh

    def proj(name):
        y = F.linear(x, sd[f"{p}.attention.self.{name}.weight"], sd[f"{p}.attention.self.{name}.bias"])
        return y.view(B, L, nh, d).permute(0, 2, 1, 3)

    q, k, v = proj("query"), proj("key"), proj("value")
    scores = q @ k.transpose(-1, -2) / math.sqrt(d) + ext_mask
    probs = dropout(torch.softmax(scores, dim=-1), cfg.attention_probs_dropout_prob, train, site=site_id(p, 1))
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, L, H)
    return ctx


def bert_layer(sd, p, cfg, x, ext_mask, train=False):
    ctx = self_attention(sd, p, cfg, x, ext_mask, train)
    y = F.linear(ctx, sd[f"{p}.attention.output.dense.weight"], sd[f"{p}.attention.output.dense.bias"])
    y = dropout(y, cfg.hidden_dropout_prob, train, site=site_id(p, 2))
    x = layer_norm(y + x, sd[f"{p}.attention.output.LayerNorm.weight"],
                   sd[f"{p}.attention.output.LayerNorm.bias"], cfg.layer_norm_eps)
    h = gelu_erf(F.linear(x, sd[f"{p}.intermediate.dense.weight"], sd[f"{p}.intermediate.dense.bias"]))
    y = F.linear(h, sd[f"{p}.output.dense.weight"], sd[f"{p}.output.dense.bias"])
    y = dropout(y, cfg.hidden_dropout_prob, train, site=site_id(p, 3))
    return layer_norm(y + x, sd[f"{p}.output.LayerNorm.weight"], sd[f"{p}.output.LayerNorm.bias"],
                      cfg.layer_norm_eps)


def bert_model(sd, prefix, n_layers, cfg, attention_mask, input_ids=None, inputs_embeds=None,
               position_ids=None, train=False):
    """BertModel.forward -> sequence_output (the pooler output is discarded by every caller)."""
    ext_mask = (1.0 - attention_mask.to(torch.float32))[:, None, None, :] * -10000.0
    x = embeddings(sd, prefix, cfg, input_ids, inputs_embeds, position_ids, train)
    for i in range(n_layers):
        x = bert_layer(sd, f"{prefix}.encoder.layer.{i}", cfg, x, ext_mask, train)
    return x


def gru_final(sd, pho_idx, pho_lens):
    """nn.Embedding -> pack_padded_sequence -> 1-layer GRU (h0 = 0) -> final hidden per sequence.
    Gate order r, z, n;  n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1-z) n + z h."""
    emb = sd["pho_embeddings.weight"][pho_idx]                      # [N, T, H]
    w_ih, w_hh = sd["pho_gru.weight_ih_l0"], sd["pho_gru.weight_hh_l0"]
    b_ih, b_hh = sd["pho_gru.bias_ih_l0"], sd["pho_gru.bias_hh_l0"]
    N, T, H = emb.shape
    if FAST and not (w_ih.requires_grad or w_hh.requires_grad or emb.requires_grad):   # (the module path below cuts autograd)
        packed = torch.nn.utils.rnn.pack_padded_sequence(emb, pho_lens, batch_first=True, enforce_sorted=False)
        gru = torch.nn.GRU(H, H, num_layers=1, batch_first=True, device=emb.device)
        gru.weight_ih_l0, gru.weight_hh_l0 = torch.nn.Parameter(w_ih), torch.nn.Parameter(w_hh)
        gru.bias_ih_l0, gru.bias_hh_l0 = torch.nn.Parameter(b_ih), torch.nn.Parameter(b_hh)
        return gru(packed)[1].squeeze(0)
    lens = torch.as_tensor(pho_lens, dtype=torch.long, device=emb.device)
    h = torch.zeros(N, H, device=emb.device, dtype=emb.dtype)
    for t in range(T):
        gi = emb[:, t] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h_new = (1.0 - z) * n + z * h
        active = (lens > t).unsqueeze(1)
        h = torch.where(active, h_new, h)
    return h


def batch_norm(sd, p, x, train, stats=None):
    """nn.BatchNorm2d (eps 1e-5, momentum 0.1).  train=True uses biased batch statistics over
    (N, H, W) and records the running-stat update in `stats` (unbiased variance)."""
    w, b = sd[f"{p}.weight"], sd[f"{p}.bias"]
    if FAST and not train:
        return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], w, b, False, 0.1, 1e-5)
    if train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats is not None:
            n = x.numel() / x.shape[1]
            stats[f"{p}.running_mean"] = 0.9 * sd[f"{p}.running_mean"] + 0.1 * mean.detach()
            stats[f"{p}.running_var"] = 0.9 * sd[f"{p}.running_var"] + 0.1 * var.detach() * n / max(n - 1, 1)
            stats[f"{p}.num_batches_tracked"] = sd[f"{p}.num_batches_tracked"] + 1
    else:
        mean, var = sd[f"{p}.running_mean"], sd[f"{p}.running_var"]
    inv = torch.rsqrt(var + 1e-5)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


# ReLU gates: a reduced-precision forward flips the sign of a few pre-activations that sit within its error of
# zero, and every flipped gate moves a whole gradient element.  To check the BACKWARD kernels tightly the
# train-mode parity test installs RELU_MASK_FN(site, shape) -> {0,1} gates taken from the CUDA forward, so both
# sides differentiate through identical gates (same idea as MASK_FN for dropout).  Sites: "<block>.a1", "<block>.out".
RELU_MASK_FN = None


def relu_gate(x, site):
    if RELU_MASK_FN is None:
        return torch.relu(x)
    return x * RELU_MASK_FN(site, tuple(x.shape)).to(x.dtype)


def basic_block(sd, p, x, train, stats=None):
    y = F.conv2d(x, sd[f"{p}.residual_function.0.weight"], stride=2, padding=1)
    y = relu_gate(batch_norm(sd, f"{p}.residual_function.1", y, train, stats), f"{p}.a1")
    y = F.conv2d(y, sd[f"{p}.residual_function.3.weight"], stride=1, padding=1)
    y = batch_norm(sd, f"{p}.residual_function.4", y, train, stats)
    s = F.conv2d(x, sd[f"{p}.shortcut.0.weight"], stride=2)
    s = batch_norm(sd, f"{p}.shortcut.1", s, train, stats)
    return relu_gate(y + s, f"{p}.out")


def char_resnet(sd, images, train=False, stats=None, collect=None):
    h = images
    for b in range(1, RES_BLOCKS + 1):
        h = basic_block(sd, f"resnet.res_block{b}", h, train, stats)
        if collect is not None:
            collect[f"res_block{b}"] = h
    return h.squeeze(-1).squeeze(-1)


def glyph_images(sd, cfg, src_flat):
    if cfg.num_fonts == 1:
        return sd["char_images.weight"][src_flat].reshape(-1, 1, 32, 32)
    return sd["char_images_multifonts"][src_flat]


def forward(sd, batch, cfg, train=False, collect=None, bn_stats=None):
    """Returns (loss, logits) when 'tgt_idx' is in the batch else (logits,), like the reference.
    `collect` (dict) receives the sub-module outputs used by the parity tests."""
    input_ids = batch["src_idx"]
    attention_mask = batch["masks"]
    B, L = input_ids.shape
    c = collect if collect is not None else {}

    bert_h = bert_model(sd, "bert", cfg.num_hidden_layers, cfg, attention_mask, input_ids=input_ids, train=train)
    c["bert_hiddens"] = bert_h
    modal = [bert_h]
    if cfg.with_pho == "yes":
        pho_h = gru_final(sd, batch["pho_idx"], batch["pho_lens"]).reshape(B, L, -1)
        c["pho_gru"] = pho_h
        pho_h = bert_model(sd, "pho_model", 4, cfg, attention_mask, inputs_embeds=pho_h, train=train)
        c["pho_hiddens"] = pho_h
        modal.append(pho_h)
    if cfg.with_res == "yes":
        res = char_resnet(sd, glyph_images(sd, cfg, input_ids.reshape(-1)), train, bn_stats, collect)
        c["resnet"] = res.reshape(B, L, -1)
        res_h = layer_norm(res.reshape(B, L, -1), sd["resnet_layernorm.weight"], sd["resnet_layernorm.bias"],
                           cfg.layer_norm_eps)
        c["res_hiddens"] = res_h
        modal.append(res_h)
    if cfg.fusion == "gate":
        m = attention_mask.to(torch.float32)
        mean = (bert_h * m.unsqueeze(2)).sum(1) / m.sum(1, keepdim=True)
        cat = torch.cat(modal + [mean.unsqueeze(1).expand(-1, L, -1)], dim=-1)
        g = torch.sigmoid(cat @ sd["gate_net.weight"].t() + sd["gate_net.bias"])
        c["gates"] = g
        hid = sum(g[:, :, i:i + 1] * modal[i] for i in range(len(modal)))
    else:
        hid = sum(modal)
    c["fused"] = hid
    seq = bert_model(sd, "output_block", 3, cfg, attention_mask, inputs_embeds=hid,
                     position_ids=torch.zeros(B, L, dtype=torch.long, device=hid.device), train=train)
    c["sequence_output"] = seq
    seq = dropout(seq, cfg.hidden_dropout_prob, train, site=SITE_FINAL)
    logits = F.linear(seq, sd["classifier.weight"], sd["classifier.bias"])
    c["logits"] = logits
    if "tgt_idx" not in batch:
        return (logits,)
    active = batch["loss_masks"].reshape(-1) == 1
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1])[active], batch["tgt_idx"].reshape(-1)[active])
    c["loss"] = loss
    return (loss, logits)
