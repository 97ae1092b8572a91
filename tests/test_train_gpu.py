"""GPU: train-mode forward/backward (semantic path) and the fused clip+AdamW against the CPU oracle's autograd.

Protocol: dropout probabilities 0, same seeded weights and batch; compare loss, every parameter
gradient, and one optimizer step.  Tolerances: gradients are bf16-operand GEMM results -> relative L2 error per
tensor <= 2e-2 (measured: median 3e-3, worst 1e-2); gradients that are analytically zero (key bias: softmax is
shift-invariant) are compared absolutely.
"""
import math

import pytest
import torch

from conftest import cached_state_dict
from realise_b200.synth import ArchConfig, synth_batch

pytestmark = pytest.mark.gpu

TRAIN_LOGIT_TOL = 1.5e-2   # TRAIN-mode logits are computed with bf16 operands (one format per MMA, gradients need bf16's
                           # range): the max over ~10^7 logits of a bf16-rounding-sized error lands at 1.2-1.4e-2 (rms 2.4e-3)
                           # against north_star's 1e-2, which the INFERENCE path (fp16 operands, the configuration
                           # BASELINE configs[1] checks logits on) meets with 5.7e-3 at B64 x L128


def _setup(layers=2):
    from realise_b200.model import SpellBertPho2ResArch3Abla
    cfg = ArchConfig(num_hidden_layers=layers, with_pho="no", with_res="no", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    sd = cached_state_dict(cfg, 11)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.train().cuda()


def _oracle_grads(sd, batch, cfg):
    from oracle import realise_oracle as O
    rsd = {k: v.clone() for k, v in sd.items()}
    rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
    leaves = {}
    for k, v in rsd.items():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
            leaves[k] = v
    loss, logits = O.forward(rsd, batch, cfg, train=True)
    loss.backward()
    return loss, logits, leaves


@pytest.mark.parametrize("B,L,seed", [(2, 16, 5), (3, 40, 6), (1, 128, 7), (2, 200, 8), (1, 256, 9)])
def test_backward_matches_oracle_autograd(B, L, seed):
    cfg, sd, model = _setup()
    batch = synth_batch(B, L, seed=seed)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, logits = model(db)
    loss.backward()
    rloss, rlogits, leaves = _oracle_grads(sd, batch, cfg)
    assert abs(loss.item() - rloss.item()) <= 1e-2
    assert (logits.float().cpu() - rlogits).abs().max().item() <= TRAIN_LOGIT_TOL
    gmax = max(v.grad.abs().max().item() for v in leaves.values() if v.grad is not None)
    checked = 0
    for name, p in model.named_parameters():
        if name == "classifier.weight":
            continue
        rg = leaves[name].grad
        if p.grad is None:
            assert rg is None or rg.abs().max().item() == 0.0, name
            continue
        g = p.grad.float().cpu()
        if rg.norm().item() < 1e-6 * gmax:          # analytically zero gradients
            assert g.abs().max().item() <= 1e-4 * gmax, name
        else:
            rel = (g - rg).norm().item() / rg.norm().item()
            assert rel <= 2e-2, (name, rel)
        checked += 1
    assert checked >= 90


def test_pinyin_branch_backward_matches_oracle_autograd():
    """with_pho='yes': GRU backward-through-time (table, W_hh, biases, pho_embeddings) + pho_model stack."""
    from realise_b200.model import SpellBertPho2ResArch3Abla
    cfg = ArchConfig(num_hidden_layers=1, with_pho="yes", with_res="no", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    sd = cached_state_dict(cfg, 12)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.train().cuda()
    batch = synth_batch(3, 24, seed=8)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, _ = model(db)
    loss.backward()
    rloss, _, leaves = _oracle_grads(sd, batch, cfg)
    assert abs(loss.item() - rloss.item()) <= 1e-2
    gmax = max(v.grad.abs().max().item() for v in leaves.values() if v.grad is not None)
    seen = set()
    for name, p in model.named_parameters():
        if name == "classifier.weight" or p.grad is None:
            continue
        rg = leaves[name].grad
        if rg.norm().item() < 1e-6 * gmax:
            continue
        rel = (p.grad.float().cpu() - rg).norm().item() / rg.norm().item()
        assert rel <= 2e-2, (name, rel)
        seen.add(name)
    assert {"pho_gru.weight_ih_l0", "pho_gru.weight_hh_l0", "pho_gru.bias_ih_l0", "pho_gru.bias_hh_l0",
            "pho_embeddings.weight"} <= seen


def _unsplit(x, n, S, C):
    h = S // 2
    return x.view(n, 2, 2, h, h, C).permute(0, 5, 3, 1, 4, 2).reshape(n, C, S, S)


@pytest.mark.parametrize("with_pho,B,L", [("no", 2, 16), ("yes", 3, 24)])
def test_glyph_branch_and_full_arch3_backward(with_pho, B, L):
    """CharResNet with batch-statistics BatchNorm (+ the full three-encoder model).  The bf16-operand forward flips
    the sign of a few ReLU pre-activations that lie within its error of zero; each flip moves a whole gradient
    element, so the oracle differentiates through the CUDA forward's ReLU gates (exported via engine.debug), exactly
    as it is handed the kernels' dropout masks.  Without shared gates the CNN gradients differ by ~20 % in norm while
    every non-CNN gradient still agrees to 1 % — see DESIGN.md."""
    from oracle import realise_oracle as O
    from realise_b200.model import SpellBertPho2ResArch3Abla
    from realise_b200.train import TrainEngine
    cfg = ArchConfig(num_hidden_layers=1, with_pho=with_pho, with_res="yes", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    sd = cached_state_dict(cfg, 13)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.train().cuda()
    model._engine = TrainEngine(model)
    model._engine.debug = dbg = {}
    batch = synth_batch(B, L, seed=9)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, _ = model(db)
    loss.backward()
    n = B * L

    def gates(site, shape):
        b = int(site.split("res_block")[1][0])
        S, C = shape[-1], shape[1]
        if site.endswith(".a1"):
            t = dbg[f"a1_{b}"].float().cpu().view(n, S, S, C).permute(0, 3, 1, 2)
        else:
            t = dbg[f"out{b}"].float().cpu()
            t = _unsplit(t, n, S, C) if S >= 2 else t.view(n, C, 1, 1)
        return (t > 0).float()

    rsd = {k: v.clone() for k, v in sd.items()}
    rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
    leaves = {}
    for k, v in rsd.items():
        if v.dtype.is_floating_point and "running" not in k and not k.startswith("char_images"):
            v.requires_grad_(True)
            leaves[k] = v
    stats = {}
    O.RELU_MASK_FN = gates
    try:
        rloss, _ = O.forward(rsd, batch, cfg, train=True, bn_stats=stats)
        rloss.backward()
    finally:
        O.RELU_MASK_FN = None
    assert abs(loss.item() - rloss.item()) <= 1e-2
    gmax = max(v.grad.abs().max().item() for v in leaves.values() if v.grad is not None)
    n_res = 0
    for name, p in model.named_parameters():
        if name == "classifier.weight" or name.startswith("char_images") or p.grad is None:
            continue
        rg = leaves[name].grad
        if rg.norm().item() < 1e-6 * gmax:
            continue
        rel = (p.grad.float().cpu() - rg).norm().item() / rg.norm().item()
        # the batch-stat CNN forward carries ~2 % error with 32..72 glyphs per BatchNorm batch; it also enters the
        # fused hidden state, so the tolerance of this test is 5e-2 for every tensor (2e-2 in the CNN-free tests)
        assert rel <= 5e-2, (name, rel)
        n_res += name.startswith("resnet")
    assert n_res == 47   # 15 convs + 15 BatchNorm (weight, bias) + resnet_layernorm (weight, bias)
    # BatchNorm running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance, counter)
    for b in (1, 3, 5):
        bn = getattr(model.resnet, f"res_block{b}").residual_function[1]
        key = f"resnet.res_block{b}.residual_function.1"
        assert (bn.running_mean.cpu() - stats[key + ".running_mean"]).abs().max().item() <= 2e-3
        assert (bn.running_var.cpu() - stats[key + ".running_var"]).abs().max().item() <= 2e-3
        assert bn.num_batches_tracked.item() == 1


def test_fused_adamw_step_matches_reference_formula():
    from realise_b200.optim import FusedAdamW
    cfg, sd, model = _setup(layers=1)
    batch = synth_batch(2, 16, seed=5)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, _ = model(db)
    loss.backward()
    named = [(n, p) for n, p in model.named_parameters() if p.grad is not None]
    no_decay = [p for n, p in named if "bias" in n or "LayerNorm.weight" in n]      # src/run.py:146-151
    decay = [p for n, p in named if not ("bias" in n or "LayerNorm.weight" in n)]
    opt = FusedAdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=5e-5,
                     eps=1e-8, max_grad_norm=1.0, model=model)
    before = {id(p): (p.detach().cpu().clone(), p.grad.detach().cpu().clone()) for _, p in named}
    opt.step()
    torch.cuda.synchronize()
    gn = math.sqrt(sum((g.double() ** 2).sum().item() for _, g in before.values()))
    assert abs(opt.grad_norm() - gn) <= 1e-4 * gn
    coef = min(1.0, 1.0 / (gn + 1e-6))                                               # clip_grad_norm_(.., 1.0)
    for n, p in named:
        p0, g0 = before[id(p)]
        wd = 0.0 if ("bias" in n or "LayerNorm.weight" in n) else 0.01
        g = g0 * coef                                                                # optimization.py:113-169, step 1
        m, v = 0.1 * g, 0.001 * g * g
        ref = p0 - 5e-5 * math.sqrt(1 - 0.999) / (1 - 0.9) * m / (v.sqrt() + 1e-8)
        ref = ref - 5e-5 * wd * ref
        assert (p.detach().cpu() - ref).abs().max().item() <= 1e-6, n
    # the bf16 / f32 operand copies used by the kernels were refreshed by the same pass
    P = model._prepared
    lyr = model.bert.encoder.layer[0]
    assert torch.equal(P["bert"]["layers"][0]["w_qkv"][:768], lyr.attention.self.query.weight.detach().to(P["half"]))
    assert torch.equal(P["bert"]["layers"][0]["b_qkv"][768:1536], lyr.attention.self.key.bias.detach())
    loss2, _ = model(db)
    assert loss2.item() < loss.item()


def test_train_mode_guards():
    from realise_b200.model import SpellBertPho2ResArch3
    m = SpellBertPho2ResArch3(ArchConfig(num_hidden_layers=1)).train().cuda()
    batch = synth_batch(2, 16, seed=1, with_labels=False)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    with pytest.raises(RuntimeError, match="tgt_idx"):
        m(db)
    long = synth_batch(1, 300, seed=1)
    with pytest.raises(NotImplementedError, match="seq_len"):
        m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in long.items()})


@pytest.mark.parametrize("B,L", [(3, 40), (2, 192)])
def test_dropout_training_parity_with_exported_masks(B, L):
    """p = 0.1 everywhere (the reference config): the kernels' counter-based masks are exported with
    rl_dropout_mask and installed in the oracle, so both sides drop exactly the same elements.  L = 192 runs the
    128 x 128-blocked attention backward (seq_len > 128)."""
    from oracle import realise_oracle as O
    from realise_b200 import ops
    from realise_b200.model import SpellBertPho2ResArch3Abla
    from realise_b200.train import TrainEngine
    cfg = ArchConfig(num_hidden_layers=2, with_pho="no", with_res="no")
    assert cfg.hidden_dropout_prob == 0.1 and cfg.attention_probs_dropout_prob == 0.1
    sd = cached_state_dict(ArchConfig(num_hidden_layers=2, with_pho="no", with_res="no", hidden_dropout_prob=0.0,
                                      attention_probs_dropout_prob=0.0), 11)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.train().cuda()
    model._engine = TrainEngine(model)
    model._engine.set_seed(4242)
    batch = synth_batch(B, L, seed=6)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, logits = model(db)
    loss.backward()

    def mask_fn(site, shape):
        n = int(torch.tensor(shape).prod().item())
        p = 0.1
        return ops.dropout_mask(n, p, 4242, site).cpu().reshape(shape).float()

    O.MASK_FN = mask_fn
    try:
        rloss, rlogits, leaves = _oracle_grads(sd, batch, cfg)
    finally:
        O.MASK_FN = None
    assert abs(loss.item() - rloss.item()) <= 1e-2
    keep = ops.dropout_mask(1 << 20, 0.1, 4242, 1012).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    gmax = max(v.grad.abs().max().item() for v in leaves.values() if v.grad is not None)
    for name, p in model.named_parameters():
        if name == "classifier.weight" or p.grad is None:
            continue
        rg = leaves[name].grad
        if rg.norm().item() < 1e-6 * gmax:
            continue
        rel = (p.grad.float().cpu() - rg).norm().item() / rg.norm().item()
        assert rel <= 2e-2, (name, rel)
    # a different seed gives different masks (the loss moves), the same seed reproduces the loss bit for bit
    model._engine.set_seed(4242)
    l_same, _ = model(db)
    model._engine.set_seed(7)
    l_other, _ = model(db)
    assert l_same.item() == loss.item() and l_other.item() != loss.item()


def test_flat_gradient_buffer_layout():
    cfg, sd, model = _setup(layers=1)
    batch = synth_batch(2, 16, seed=5)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, _ = model(db)
    loss.backward()
    eng = model._engine
    lo, hi = eng.flat.data_ptr(), eng.flat.data_ptr() + eng.flat.numel() * 4
    n = 0
    for p in model.parameters():
        if p.grad is not None:
            assert lo <= p.grad.data_ptr() < hi          # every gradient is a view of the one flat buffer
            n += p.grad.numel()
    assert n <= eng.flat.numel() <= n + 64 * sum(p.grad is not None for p in model.parameters())   # 256-byte padding
    no_grad = [name for name, p in model.named_parameters() if p.grad is None]
    assert sorted(no_grad) == sorted(["bert.pooler.dense.weight", "bert.pooler.dense.bias",
                                      "output_block.pooler.dense.weight", "output_block.pooler.dense.bias",
                                      "output_block.embeddings.word_embeddings.weight"])


def test_colsum_and_bn_kernels_match_torch():
    """Vectorised bias-gradient column sums, BN batch statistics and the fused two-branch BN backward vs torch."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, cols in [(1000, 768), (4096, 2304), (333, 21128), (77, 40)]:
        x = torch.randn(rows, cols, device="cuda", generator=g).bfloat16()
        out = torch.zeros(cols, device="cuda")
        ops.colsum_bf16(x, out)
        ref = x.float().sum(0)
        assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    for M, C, S, xdt in [(2 * 256, 64, 16, torch.float32), (3 * 64, 128, 8, torch.float32), (5 * 16, 256, 4, torch.float32),
                         (7 * 4, 512, 2, torch.float32), (9, 768, 1, torch.float32), (2 * 256, 64, 16, torch.bfloat16),
                         (3 * 64, 128, 8, torch.bfloat16)]:   # res_block1-2 keep raw conv outputs in bf16
        x1 = (torch.randn(M, C, device="cuda", generator=g) * 2 + 0.5).to(xdt)
        x2 = torch.randn(M, C, device="cuda", generator=g).to(xdt)
        sums = torch.zeros(2 * C, device="cuda")
        ops.bn_stats(x1, sums)
        assert torch.allclose(sums[:C], x1.float().sum(0), rtol=1e-4, atol=1e-2)
        assert torch.allclose(sums[C:], (x1.float() * x1.float()).sum(0), rtol=1e-4, atol=1e-2)
        # out = relu(bn(x1) + bn(x2)) with gamma/beta; dy arrives parity-split when S >= 2 (layout of the next block's input)
        gam = [torch.rand(C, device="cuda", generator=g) + 0.5 for _ in range(2)]
        bet = [torch.randn(C, device="cuda", generator=g) * 0.1 for _ in range(2)]
        xs = [x1.float().clone().requires_grad_(True), x2.float().clone().requires_grad_(True)]
        gl = [t.clone().requires_grad_(True) for t in gam]
        bl = [t.clone().requires_grad_(True) for t in bet]
        stats = []
        y = 0
        for x, ga, be in zip(xs, gl, bl):
            mu, var = x.mean(0), x.var(0, unbiased=False)
            rstd = (var + 1e-5).rsqrt()
            stats.append((mu.detach(), rstd.detach()))
            y = y + (x - mu) * rstd * ga + be
        out = torch.relu(y)
        dy = torch.randn(M, C, device="cuda", generator=g).bfloat16()
        out.backward(dy.float())
        remap = S >= 2
        if remap:
            n = M // (S * S)
            perm = out.detach().view(n, S // 2, 2, S // 2, 2, C).permute(0, 2, 4, 1, 3, 5).reshape(M, C)
            dperm = dy.view(n, S // 2, 2, S // 2, 2, C).permute(0, 2, 4, 1, 3, 5).reshape(M, C).contiguous()
        else:
            perm, dperm = out.detach(), dy
        act = perm.bfloat16().contiguous()
        dcat = torch.zeros(M, 2 * C, device="cuda", dtype=torch.bfloat16)
        dx1 = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16)
        db = [torch.zeros(C, device="cuda") for _ in range(2)]
        dg = [torch.zeros(C, device="cuda") for _ in range(2)]
        ops.bn_bwd2(dperm, act, (x1, stats[0][0], stats[0][1], gam[0], db[0], dg[0], dx1),
                    (x2, stats[1][0], stats[1][1], gam[1], db[1], dg[1], dcat[:, C:]), M, C, remap=remap, map_hw=(S, S))
        for i in range(2):
            assert torch.allclose(db[i], bl[i].grad, rtol=2e-3, atol=2e-2), (M, C, i)
            assert torch.allclose(dg[i], gl[i].grad, rtol=2e-3, atol=2e-2), (M, C, i)
        for got, ref in ((dx1, xs[0].grad), (dcat[:, C:], xs[1].grad)):
            rel = (got.float() - ref).norm().item() / ref.norm().item()
            assert rel <= 1e-2, (M, C, rel)
        # the same with the ReLU mask re-derived from the raw conv outputs (fwd = the forward's scale / shift): act_out unread
        fwd = [(ga * st[1], be - st[0] * ga * st[1]) for ga, be, st in zip(gam, bet, stats)]
        dcat_r = torch.zeros_like(dcat); dx1_r = torch.zeros_like(dx1)
        db_r = [torch.zeros(C, device="cuda") for _ in range(2)]
        dg_r = [torch.zeros(C, device="cuda") for _ in range(2)]
        ops.bn_bwd2(dperm, None, (x1, stats[0][0], stats[0][1], gam[0], db_r[0], dg_r[0], dx1_r),
                    (x2, stats[1][0], stats[1][1], gam[1], db_r[1], dg_r[1], dcat_r[:, C:]), M, C, remap=remap, map_hw=(S, S),
                    fwd=fwd)
        for i in range(2):
            assert torch.allclose(db_r[i], bl[i].grad, rtol=2e-3, atol=2e-2), (M, C, i)
            assert torch.allclose(dg_r[i], gl[i].grad, rtol=2e-3, atol=2e-2), (M, C, i)
        for got, ref in ((dx1_r, xs[0].grad), (dcat_r[:, C:], xs[1].grad)):
            rel = (got.float() - ref).norm().item() / ref.norm().item()
            assert rel <= 1e-2, (M, C, rel)


def test_dropout_counter_offsets_the_seed():
    """drop_counter: kernels use seed + *counter, read at run time (what makes graph replays draw new masks)."""
    from realise_b200 import ops
    ctr = torch.tensor([5], device="cuda", dtype=torch.int64)
    base = ops.dropout_mask(1 << 16, 0.1, 1000, 77)
    plus5 = ops.dropout_mask(1 << 16, 0.1, 1005, 77)
    with ops.dropout_counter(ctr):
        via_ptr = ops.dropout_mask(1 << 16, 0.1, 1000, 77)
    again = ops.dropout_mask(1 << 16, 0.1, 1000, 77)
    assert torch.equal(via_ptr, plus5) and torch.equal(again, base) and not torch.equal(base, plus5)


def test_graphed_train_step_matches_eager_steps():
    """GraphedTrainStep (fwd + bwd + clip + AdamW as one CUDA-graph replay) follows the eager loop: same losses and
    parameters after 4 steps with dropout off (split-K atomics reorder fp32 sums -> small tolerance); with dropout on,
    replays draw different masks."""
    from realise_b200.graphed import GraphedTrainStep
    from realise_b200.model import SpellBertPho2ResArch3Abla
    from realise_b200.optim import FusedAdamW
    cfg = ArchConfig(num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sd = cached_state_dict(cfg, 13)
    batches = [synth_batch(2, 16, seed=20 + i, ragged=False) for i in range(4)]
    T = max(b["pho_idx"].shape[1] for b in batches)
    for b in batches:   # one (B, L, T) shape so that the graph is reused
        pad = torch.zeros(b["pho_idx"].shape[0], T, dtype=torch.int64)
        pad[:, :b["pho_idx"].shape[1]] = b["pho_idx"]
        b["pho_idx"] = pad
    results = []
    for graphed in (False, True):
        model = SpellBertPho2ResArch3Abla(cfg)
        model.tie_cls_weight()
        model.load_state_dict(sd, strict=True)
        model.train().cuda()
        opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=2e-5, max_grad_norm=1.0, model=model)
        step = GraphedTrainStep(model, opt)
        losses = []
        for i, b in enumerate(batches):
            for gr in opt.param_groups:
                gr["lr"] = 2e-5 * (1 + i)                     # a moving schedule must reach the captured update
            if graphed:
                losses.append(step(b).item())
            else:
                db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
                loss = model(db)[0]
                loss.backward()
                opt.step()
                losses.append(loss.item())
        if graphed:
            assert step.replays == 3                           # first step of the shape is eager, the rest replay
        results.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
    (l0, p0), (l1, p1) = results
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (l0, l1)
    # Adam's first steps are sign-like (|update| ~ lr whatever |g|): an element whose tiny gradient flips sign under the
    # split-K atomics' reordering moves the other way, so parameters are compared through the size of their UPDATE
    init = {n: v.cuda() for n, v in sd.items()}
    moved, tot_d, tot_u = 0, 0.0, 0.0
    for n in p0:
        if n == "classifier.weight" or n.endswith("key.bias"):   # key bias: analytically zero gradient (pure noise)
            continue
        upd = (p0[n] - init[n]).norm().item()
        if upd == 0.0:
            assert torch.equal(p0[n], p1[n]), n
            continue
        moved += 1
        dif = (p0[n] - p1[n]).norm().item()
        assert dif <= 1.0 * upd, (n, dif, upd)               # per tensor: never an unrelated update
        tot_d += dif * dif
        tot_u += upd * upd
    assert moved >= 100 and tot_d <= 0.35 ** 2 * tot_u, (moved, tot_d, tot_u)   # overall: the same trajectory
    # a schedule frozen at capture time (the lr of step 2 reused for steps 3 and 4) would move the weights ~30 % less
    w = "bert.encoder.layer.0.output.dense.weight"
    mv = [(pp[w] - init[w]).abs().mean().item() for pp in (p0, p1)]
    assert abs(mv[0] - mv[1]) <= 0.08 * mv[0], mv
    # dropout on: two replays on the same batch and (nearly) the same weights see different masks
    cfg2 = ArchConfig(num_hidden_layers=1, with_pho="no", with_res="no")
    model = SpellBertPho2ResArch3Abla(cfg2)
    model.tie_cls_weight()
    model.train().cuda()
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=0.0, max_grad_norm=1.0, model=model)
    step = GraphedTrainStep(model, opt)
    ls = [step(batches[0]).item() for _ in range(4)]
    assert step.replays == 3 and len({round(x, 6) for x in ls[1:]}) == 3, ls


def test_gemm_operand_formats():
    """bf16 and fp16 operands / outputs per call (no process-wide switch); one MMA cannot mix the two formats (a mixed
    tcgen05 descriptor is an illegal instruction on B200 — tools/probe_mixed_mma.py), so the library rejects it."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 512, 384, 256
    a32 = torch.randn(M, K, device="cuda", generator=g)
    b32 = torch.randn(N, K, device="cuda", generator=g)
    for dt in (torch.float16, torch.bfloat16):
        a, b = a32.to(dt), b32.to(dt)
        ref = a.float() @ b.float().t()
        out = torch.empty(M, N, device="cuda", dtype=torch.float32)
        ops.gemm(a, b, out)
        assert (out - ref).abs().max().item() <= 2e-3, dt
        acc = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        ops.gemm(a.t().contiguous(), b.t().contiguous(), acc, a_t=True, b_t=True, split_k=-1)
        assert (acc - ref).abs().max().item() <= 2e-3, (dt, "mn-major")
        for odt in (torch.float16, torch.bfloat16):          # the output format is independent of the operands'
            o16 = torch.empty(M, N, device="cuda", dtype=odt)
            ops.gemm(a, b, o16)
            assert (o16.float() - ref).abs().max().item() <= (0.05 if odt is torch.float16 else 0.3), (dt, odt)
    with pytest.raises(RuntimeError, match="share one 16-bit format"):
        ops.gemm(a32.bfloat16(), b32.half(), torch.empty(M, N, device="cuda"))


def _small_model(cfg, seed):
    from realise_b200.model import SpellBertPho2ResArch3Abla
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.load_state_dict(cached_state_dict(cfg, seed), strict=True)
    return model.train().cuda()


def _dev(b):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}


def test_graphed_step_survives_an_eval_cycle():
    """model.eval() drops the operand cache a captured train graph points into; the next graphed step
    must re-capture against the new cache instead of replaying reads/writes of freed memory."""
    from realise_b200.graphed import GraphedTrainStep
    from realise_b200.optim import FusedAdamW
    cfg = ArchConfig(num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = _small_model(cfg, 13)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, model=model)
    step = GraphedTrainStep(model, opt)
    b = synth_batch(2, 16, seed=20, ragged=False)
    losses = [step(b).item() for _ in range(3)]
    assert step.replays == 2
    w = model.bert.encoder.layer[0].output.dense.weight
    model.eval()
    with torch.no_grad():
        model(_dev(b))                       # validation pass: builds the eval cache, frees the training one
    torch.cuda.empty_cache()
    model.train()
    junk = torch.full((64 << 20,), float("nan"), device="cuda")   # whatever reuses the freed blocks now holds NaNs
    before = w.detach().clone()
    losses += [step(b).item() for _ in range(3)]
    torch.cuda.synchronize()
    del junk
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]
    assert torch.isfinite(w).all() and not torch.equal(before, w)
    P = model._prepared
    assert torch.equal(P["bert"]["layers"][0]["w_2"], w.detach().to(P["half"]))     # the live cache is the refreshed one


def test_gradient_accumulation_and_probe_forward():
    """successive backward() calls accumulate until the optimizer / zero_grad consumes them
    (src/run.py:193-205); a second forward before the backward does not clobber the first one's activations."""
    cfg = ArchConfig(num_hidden_layers=1, with_res="no", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = _small_model(cfg, 12)
    b1, b2 = _dev(synth_batch(2, 16, seed=31)), _dev(synth_batch(2, 16, seed=32))
    name = "bert.encoder.layer.0.intermediate.dense.weight"
    p = dict(model.named_parameters())[name]
    model(b1)[0].backward()
    g1 = p.grad.clone()
    model.zero_grad()
    assert p.grad is None
    model(b2)[0].backward()
    g2 = p.grad.clone()
    model.zero_grad()
    l1 = model(b1)[0]
    model(b2)                                # probe forward in between: must not disturb l1's saved activations
    l1.backward()
    assert ((p.grad - g1).norm() / g1.norm()).item() <= 1e-3        # (split-K atomics reorder fp32 sums run to run)
    model(b2)[0].backward()                  # no zero_grad: accumulates
    ref = g1 + g2
    assert ((p.grad - ref).norm() / ref.norm()).item() <= 1e-3
    with pytest.raises(RuntimeError, match="twice"):
        l1.backward()


def test_stale_operand_copies_are_refreshed_for_foreign_optimizers():
    """the GEMMs read 16-bit copies of the weights; an optimizer other than FusedAdamW(model=...)
    updates p.data only — the next forward must notice (version counters) and re-cast the copies."""
    cfg = ArchConfig(num_hidden_layers=1, with_res="no", with_pho="no", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    model = _small_model(cfg, 11)
    b = _dev(synth_batch(2, 16, seed=33))
    sgd = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.5)
    l0 = model(b)[0]
    l0.backward()
    sgd.step()
    sgd.zero_grad()
    l1 = model(b)[0].item()
    assert l1 < l0.item() - 1e-3             # the matmuls saw the new weights
    fresh = _small_model(cfg, 11)
    fresh.load_state_dict(model.state_dict())
    l1_fresh = fresh(b)[0].item()
    assert abs(l1 - l1_fresh) <= 1e-4


def test_gemm_many_short_tiles_per_cta():
    """Regression (found by the bench-shape parity tests): a persistent CTA that runs MANY one-chunk tiles back to back
    (N = 64, K = 32: the training stem conv over 4.2 M pixels) reused its TMA-store staging tile while the previous
    tile's store was still reading it."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    for M, N, K, odt in [(1 << 19, 64, 32, torch.bfloat16), (1 << 18, 64, 32, torch.float32), (1 << 17, 128, 64, torch.bfloat16),
                         (1 << 16, 64, 576, torch.bfloat16)]:
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=odt)
        for _ in range(3):
            ops.gemm(a, b, out)
        ref = (a.float() @ b.float().t())
        err = (out.float() - ref).abs().max().item()
        tol = (2.0 ** -8 if odt is torch.bfloat16 else 1e-5) * ref.abs().max().item()      # one output rounding
        assert err <= tol, (M, N, K, err, tol)


def test_gemm_epilogue_column_reductions():
    """rl_gemm_desc.colsum / colsumsq: per-column sum and sum of squares of the fp32 results, accumulated across tiles
    (one N tile: flushed once per CTA; several: per tile) — bias gradients / BatchNorm batch statistics."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(12)
    for M, N, K in [(70000, 64, 32), (5000, 128, 576), (3000, 768, 768), (1000, 200, 64)]:
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        b = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        cs, cq = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
        ops.gemm(a, b, out, bias=bias, colsum=cs, colsumsq=cq)
        ref = a.float() @ b.float().t() + bias
        assert (out.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()
        assert torch.allclose(cs, ref.sum(0), rtol=1e-3, atol=1e-3 * ref.sum(0).abs().max().item()), (M, N, K)
        assert torch.allclose(cq, (ref * ref).sum(0), rtol=1e-3), (M, N, K)


def test_two_font_training_step():
    """num_fonts = 2 through the training stem (glyph_im2col<2>, K = 18 of 32) against the oracle's autograd."""
    from oracle import realise_oracle as O
    cfg = ArchConfig(num_hidden_layers=1, with_pho="no", num_fonts=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = _small_model(cfg, 14)
    sd = cached_state_dict(cfg, 14)
    batch = synth_batch(4, 32, seed=15)
    loss, _ = model(_dev(batch))
    loss.backward()
    rsd = {k: v.clone() for k, v in sd.items()}
    rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
    leaves = {}
    for k, v in rsd.items():
        if v.dtype.is_floating_point and "running" not in k and not k.startswith("char_images"):
            v.requires_grad_(True)
            leaves[k] = v
    rloss, _ = O.forward(rsd, batch, cfg, train=True)
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-2
    for name in ("bert.encoder.layer.0.output.dense.weight", "gate_net.weight", "resnet_layernorm.weight"):
        g, rg = dict(model.named_parameters())[name].grad.float().cpu(), leaves[name].grad
        assert ((g - rg).norm() / rg.norm()).item() <= 5e-2, name
    g = model.resnet.res_block1.residual_function[0].weight.grad
    assert g.shape == (64, 2, 3, 3) and torch.isfinite(g).all() and float(g.abs().max()) > 0


def test_conv3x3_c64_halo_kernel_matches_generic_path_and_torch():
    """res_block1.conv2 shape (64 -> 64 channels, 16x16 maps, 9 taps): the resident-weight / row-halo kernel against the
    generic nine-tap implicit GEMM (tune_no_pair = 5 keeps the generic path) and against torch conv2d, for the forward tap
    order, the mirrored order of the data gradient, a bias vector, and an image count that leaves a partial last wave."""
    import torch.nn.functional as F
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    for n_img, mirrored, with_bias in [(3, False, False), (150, True, True), (37, False, True)]:
        x = torch.randn(n_img, 16, 16, 64, device="cuda", generator=g).bfloat16()
        w4 = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05).bfloat16()       # [co, ci, kh, kw]
        bias = torch.randn(64, device="cuda", generator=g) * 0.1 if with_bias else None
        taps = [(kw - 1, kh - 1, 0) for kh in range(3) for kw in range(3)]
        wk = w4.permute(0, 2, 3, 1).reshape(64, 576).contiguous()                              # tap-major [co, (kh, kw), ci]
        if mirrored:   # any order / sign convention of the nine taps must work: the kernel maps (dw, dh) -> weight block
            order = list(reversed(range(9)))
            taps = [taps[i] for i in order]
            wk = wk.view(64, 9, 64)[:, order].reshape(64, 576).contiguous()
        outs = []
        for mode in (5, 0):
            ops.TUNE_NO_PAIR = mode
            try:
                out = torch.empty(n_img * 256, 64, device="cuda", dtype=torch.bfloat16)
                ops.conv_gemm(x.view(n_img, 1, 16, 16, 64), wk, out, nimg=n_img, H=16, W=16, planes=1, taps=taps, bias=bias)
            finally:
                ops.TUNE_NO_PAIR = 0
            outs.append(out.float())
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias=bias, padding=1).permute(0, 2, 3, 1).reshape(-1, 64)
        scale = ref.abs().max().item()
        assert (outs[1] - ref).abs().max().item() <= 1e-2 * scale, (n_img, mirrored)
        # same products, different summation order of the nine taps: at most a bf16 rounding step apart
        assert (outs[0] - outs[1]).abs().max().item() <= 1.6e-2 * scale, (n_img, mirrored)
        assert ((outs[0] - outs[1]).abs() > 0).float().mean().item() < 0.2


def test_attention_bwd_fused_bias_gradient_equals_column_sums_of_dqkv():
    """rl_attention_bwd's optional dbias output (the fused q/k/v bias gradient) against the column sums of the dqkv it wrote,
    with ragged sentence lengths (rows beyond a sentence must contribute exact zeros) and dropout on the probabilities."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    for B, L, heads in [(3, 23, 2), (5, 128, 12), (2, 100, 4)]:
        H = heads * 64
        qkv = (torch.randn(B * L, 3 * H, device="cuda", generator=g) * 0.5).bfloat16()
        lens = torch.randint(max(1, L // 3), L + 1, (B,), device="cuda", generator=g)
        mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long().contiguous()
        ctx = torch.empty(B * L, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B * heads * L, device="cuda")
        dctx = (torch.randn(B * L, H, device="cuda", generator=g) * 0.1).bfloat16()
        drop = (0.1, 99, 3)
        ops.attention(qkv, mask, ctx, B, L, heads, drop=drop, lse=lse)
        dq_ref = torch.zeros_like(qkv)
        ops.attention_bwd(qkv, mask, ctx, dctx, dq_ref, B, L, heads, drop=drop, lse=lse)
        dq = torch.zeros_like(qkv)
        db = torch.full((3 * H,), 0.25, device="cuda")        # accumulated INTO
        ops.attention_bwd(qkv, mask, ctx, dctx, dq, B, L, heads, drop=drop, lse=lse, dbias=db)
        assert torch.equal(dq, dq_ref)
        want = dq_ref.float().sum(0) + 0.25
        assert torch.allclose(db, want, rtol=2e-2, atol=2e-2 * float(want.abs().max())), (B, L, float((db - want).abs().max()))


def test_specialised_gemm_epilogues_match_the_generic_epilogue_bit_for_bit():
    """Every compile-time specialised epilogue (SPEC_* in gemm.cu) against the generic, runtime-flag epilogue of the same
    kernel family (tune_no_pair = 5): same arithmetic in the same order, so the results must be identical bits (split-K
    sums meet in a different order: close, not identical)."""
    from realise_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)

    def rnd(*s, scale=1.0):
        return torch.randn(*s, device="cuda", generator=g) * scale

    M = 1024
    cases = []
    # (name, N, K, kwargs builder) — K = 768: short-K pair kernel; K = 3072: long-K variant (>= 24 k-blocks)
    for K in (768, 3072):
        cases.append((f"out16 K={K}", 512, K, lambda N: dict(bias=rnd(N)), torch.bfloat16))
        cases.append((f"res32+drop K={K}", 768, K, lambda N: dict(bias=rnd(N), res=rnd(M, N), drop=(0.1, 77, 3)), torch.float32))
        cases.append((f"res32 K={K}", 768, K, lambda N: dict(res=rnd(M, N)), torch.float32))
    cases.append(("gelu_save", 512, 768, lambda N: dict(bias=rnd(N, scale=0.1), act=ops.ACT_GELU_SAVE,
                                                        out2=torch.empty(M, N, device="cuda", dtype=torch.bfloat16)), torch.bfloat16))
    cases.append(("gelu_grad", 512, 768, lambda N: dict(res=rnd(M, N).bfloat16(), act=ops.ACT_GELU_GRAD), torch.bfloat16))
    cases.append(("gelu_grad+colsum", 512, 768, lambda N: dict(res=rnd(M, N).bfloat16(), act=ops.ACT_GELU_GRAD,
                                                               colsum=torch.zeros(N, device="cuda")), torch.bfloat16))
    cases.append(("stem 128x64 tiles", 64, 32, lambda N: dict(), torch.bfloat16))
    # forward-only formats (fp16 operands and results): QKV, FFN1 + GELU
    cases.append(("fp16 out16", 512, 768, lambda N: dict(bias=rnd(N)), torch.float16))
    cases.append(("fp16 gelu", 512, 768, lambda N: dict(bias=rnd(N, scale=0.1), act=ops.ACT_GELU), torch.float16))
    for name, N, K, mk, odt in cases:
        Mi = 4096 if N == 64 else M
        h16 = torch.float16 if odt == torch.float16 else torch.bfloat16
        a = (rnd(Mi, K) * 0.5).to(h16)
        b = (rnd(N, K) * 0.05).to(h16)
        kw = mk(N)
        if "res" in kw and kw["res"].shape[0] != Mi:
            continue
        outs = []
        for mode in (5, 0):
            ops.TUNE_NO_PAIR = mode
            try:
                k2 = dict(kw)
                if "out2" in k2:
                    k2["out2"] = torch.empty_like(kw["out2"])
                if "colsum" in k2:
                    k2["colsum"] = torch.zeros_like(kw["colsum"])
                out = torch.empty(Mi, N, device="cuda", dtype=odt)
                ops.gemm(a, b, out, **k2)
            finally:
                ops.TUNE_NO_PAIR = 0
            outs.append((out, k2.get("out2"), k2.get("colsum")))
        assert torch.equal(outs[0][0], outs[1][0]), name
        if outs[0][1] is not None:
            assert torch.equal(outs[0][1], outs[1][1]), name
        if outs[0][2] is not None:
            assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-3), name
    # split-K weight gradient (TMA reduce-add): specialised vs generic agree to rounding of the different summation order
    a = (rnd(4096, 768) * 0.1).bfloat16()
    b = (rnd(4096, 512) * 0.1).bfloat16()
    outs = []
    for mode in (5, 0):
        ops.TUNE_NO_PAIR = mode
        try:
            out = torch.zeros(768, 512, device="cuda")
            ops.gemm(a, b, out, a_t=True, b_t=True, split_k=-1)
        finally:
            ops.TUNE_NO_PAIR = 0
        outs.append(out)
    ref = a.float().t() @ b.float()
    assert torch.allclose(outs[0], outs[1], rtol=1e-4, atol=1e-3 * float(ref.abs().max()))
    assert torch.allclose(outs[1], ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()))
