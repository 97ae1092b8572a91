"""GPU dev check: train-mode forward/backward + fused AdamW vs the CPU oracle's autograd (semantic-only config)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from oracle import realise_oracle as O  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402
from realise_b200.optim import FusedAdamW  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402


def ref_adamw(p, g, m, v, step, lr, b1, b2, eps, wd, coef):
    g = g * coef
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p = p - step_size * m / (v.sqrt() + eps)
    if wd > 0:
        p = p - lr * wd * p
    return p, m, v


cfg = ArchConfig(num_hidden_layers=2, with_pho="no", with_res="no", hidden_dropout_prob=0.0,
                 attention_probs_dropout_prob=0.0)
sd = synth_state_dict(cfg, 11)
model = SpellBertPho2ResArch3Abla(cfg)
model.tie_cls_weight()
model.load_state_dict(sd, strict=True)
model.train().cuda()
for (B, L, seed) in [(2, 16, 5), (3, 40, 6)]:
    batch = synth_batch(B, L, seed=seed)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    model.zero_grad(set_to_none=True)
    loss, logits = model(db)
    loss.backward()
    torch.cuda.synchronize()
    # oracle autograd
    rsd = {k: v.clone() for k, v in sd.items()}
    rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
    leaves = {}
    for k, v in rsd.items():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
            leaves[k] = v
    rloss, rlogits = O.forward(rsd, batch, cfg, train=True)
    rloss.backward()
    print(f"B{B} L{L}: loss {loss.item():.5f} vs {rloss.item():.5f}; logits err {(logits.float().cpu()-rlogits).abs().max().item():.3e}")
    worst = []
    for name, p in model.named_parameters():
        if name == "classifier.weight":
            continue
        rg = leaves[name].grad
        if p.grad is None:
            if rg is not None and rg.abs().max() > 0:
                print("  MISSING grad", name)
            continue
        g = p.grad.float().cpu()
        rel = (g - rg).norm().item() / (rg.norm().item() + 1e-12)
        worst.append((rel, name, rg.norm().item()))
    worst.sort(reverse=True)
    for rel, name, n in worst[:8]:
        print(f"  rel_err {rel:.3e}  |g|={n:.3e}  {name}")
    print(f"  median rel_err {sorted(w[0] for w in worst)[len(worst)//2]:.3e} over {len(worst)} tensors")

# ---- pinyin branch (GRU BPTT + pho_model) ----
cfg_p = ArchConfig(num_hidden_layers=1, with_pho="yes", with_res="no", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
sd_p = synth_state_dict(cfg_p, 12)
mp_ = SpellBertPho2ResArch3Abla(cfg_p)
mp_.tie_cls_weight()
mp_.load_state_dict(sd_p, strict=True)
mp_.train().cuda()
batch = synth_batch(3, 24, seed=8)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
loss, logits = mp_(db)
loss.backward()
torch.cuda.synchronize()
rsd = {k: v.clone() for k, v in sd_p.items()}
rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
leaves = {}
for k, v in rsd.items():
    if v.dtype.is_floating_point:
        v.requires_grad_(True)
        leaves[k] = v
rloss, rlogits = O.forward(rsd, batch, cfg_p, train=True)
rloss.backward()
print(f"with_pho: loss {loss.item():.5f} vs {rloss.item():.5f}")
worst = []
for name, p in mp_.named_parameters():
    if name == "classifier.weight":
        continue
    rg = leaves[name].grad
    if p.grad is None:
        if rg is not None and rg.abs().max() > 0:
            print("  MISSING grad", name, rg.norm().item())
        continue
    worst.append(((p.grad.float().cpu() - rg).norm().item() / (rg.norm().item() + 1e-12), name, rg.norm().item()))
worst.sort(reverse=True)
for rel, name, n in worst[:14]:
    print(f"  rel_err {rel:.3e}  |g|={n:.3e}  {name}")
for name in ["pho_gru.weight_ih_l0", "pho_gru.weight_hh_l0", "pho_gru.bias_ih_l0", "pho_gru.bias_hh_l0", "pho_embeddings.weight"]:
    r = [w for w in worst if w[1] == name]
    print("  ", name, r[0][0] if r else "n/a")

# ---- glyph branch (CharResNet with batch-stat BatchNorm) and the full Arch3 ----
for (wp, tag, B_, L_) in [("no", "with_res", 2, 16), ("yes", "full arch3", 3, 24)]:
    cfg_r = ArchConfig(num_hidden_layers=1, with_pho=wp, with_res="yes", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sd_r = synth_state_dict(cfg_r, 13)
    mr = SpellBertPho2ResArch3Abla(cfg_r)
    mr.tie_cls_weight()
    mr.load_state_dict(sd_r, strict=True)
    mr.train().cuda()
    batch = synth_batch(B_, L_, seed=9)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, logits = mr(db)
    loss.backward()
    torch.cuda.synchronize()
    rsd = {k: v.clone() for k, v in sd_r.items()}
    rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
    leaves = {}
    for k, v in rsd.items():
        if v.dtype.is_floating_point and "running" not in k and not k.startswith("char_images"):
            v.requires_grad_(True)
            leaves[k] = v
    stats = {}
    rloss, rlogits = O.forward(rsd, batch, cfg_r, train=True, bn_stats=stats)
    rloss.backward()
    print(f"{tag}: loss {loss.item():.5f} vs {rloss.item():.5f}")
    worst = []
    for name, p in mr.named_parameters():
        if name == "classifier.weight" or name.startswith("char_images"):
            continue
        rg = leaves[name].grad
        if p.grad is None:
            if rg is not None and rg.abs().max() > 0:
                print("  MISSING grad", name, rg.norm().item())
            continue
        worst.append(((p.grad.float().cpu() - rg).norm().item() / (rg.norm().item() + 1e-12), name, rg.norm().item()))
    worst.sort(reverse=True)
    for rel, name, n in [w for w in worst if "key.bias" not in w[1]][:10]:
        print(f"  rel_err {rel:.3e}  |g|={n:.3e}  {name}")
    resn = [w for w in worst if w[1].startswith("resnet")]
    print(f"  resnet tensors: {len(resn)}, max rel {max(w[0] for w in resn):.3e}, median {sorted(w[0] for w in resn)[len(resn)//2]:.3e}")
    bn = mr.resnet.res_block2.residual_function[1]
    print("  running_mean err", (bn.running_mean.cpu() - stats["resnet.res_block2.residual_function.1.running_mean"]).abs().max().item(),
          "running_var err", (bn.running_var.cpu() - stats["resnet.res_block2.residual_function.1.running_var"]).abs().max().item(),
          "nbt", bn.num_batches_tracked.item())

# ---- dropout ON: feed the kernels' own masks to the oracle ----
from realise_b200 import ops  # noqa: E402
cfg_d = ArchConfig(num_hidden_layers=2, with_pho="no", with_res="no")   # p = 0.1 / 0.1 as in the reference config
md = SpellBertPho2ResArch3Abla(cfg_d)
md.tie_cls_weight()
md.load_state_dict(sd, strict=True)
md.train().cuda()
B, L = 3, 40
batch = synth_batch(B, L, seed=6)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
from realise_b200.train import TrainEngine  # noqa: E402
md._engine = TrainEngine(md)
md._engine.set_seed(4242)
loss, logits = md(db)
loss.backward()
torch.cuda.synchronize()


def mask_fn(site, shape):
    n = 1
    for d in shape:
        n *= d
    p = cfg_d.attention_probs_dropout_prob if site % 10 == 1 and site != 9999 else cfg_d.hidden_dropout_prob
    return ops.dropout_mask(n, p, 4242, site).cpu().reshape(shape).float()


O.MASK_FN = mask_fn
rsd = {k: v.clone() for k, v in sd.items()}
rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
leaves = {}
for k, v in rsd.items():
    if v.dtype.is_floating_point:
        v.requires_grad_(True)
        leaves[k] = v
rloss, rlogits = O.forward(rsd, batch, cfg_d, train=True)
rloss.backward()
O.MASK_FN = None
print(f"dropout 0.1: loss {loss.item():.5f} vs {rloss.item():.5f}; logits err {(logits.float().cpu()-rlogits).abs().max().item():.3e}")
keep = ops.dropout_mask(1 << 20, 0.1, 4242, 1012).float().mean().item()
print(f"  keep rate {keep:.4f}")
worst = []
for name, p in md.named_parameters():
    if name == "classifier.weight" or p.grad is None:
        continue
    rg = leaves[name].grad
    worst.append(((p.grad.float().cpu() - rg).norm().item() / (rg.norm().item() + 1e-12), name, rg.norm().item()))
worst.sort(reverse=True)
for rel, name, n in worst[:8]:
    print(f"  rel_err {rel:.3e}  |g|={n:.3e}  {name}")

# optimizer parity on the last gradients
params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
no_decay = [p for n, p in model.named_parameters() if p.grad is not None and ("bias" in n or "LayerNorm.weight" in n)]
decay = [p for p in params if all(p is not q for q in no_decay)]
opt = FusedAdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=5e-5,
                 max_grad_norm=1.0, model=model)
names = {id(p): n for n, p in model.named_parameters()}
before = {id(p): (p.detach().cpu().clone(), p.grad.detach().cpu().clone()) for p in params}
opt.step()
torch.cuda.synchronize()
gn = math.sqrt(sum((g.double() ** 2).sum().item() for _, g in before.values()))
coef = min(1.0, 1.0 / (gn + 1e-6))
print(f"grad norm {gn:.4f} (kernel {opt.grad_norm():.4f}), clip coef {coef:.4f}")
mx = 0.0
for p in params:
    p0, g0 = before[id(p)]
    wd = 0.01 if any(p is q for q in decay) else 0.0
    ref, _, _ = ref_adamw(p0, g0, torch.zeros_like(p0), torch.zeros_like(p0), 1, 5e-5, 0.9, 0.999, 1e-8, wd, coef)
    mx = max(mx, (p.detach().cpu() - ref).abs().max().item())
print(f"AdamW step max |p - ref| = {mx:.3e}")
P = model._prepared
lyr = model.bert.encoder.layer[0]
print("shadow qkv err", (P["bert"]["layers"][0]["w_qkv"][:768].float() - lyr.attention.self.query.weight.detach().bfloat16().float()).abs().max().item(),
      "bias shadow err", (P["bert"]["layers"][0]["b_qkv"][768:1536] - lyr.attention.self.key.bias.detach()).abs().max().item())
# second step runs with refreshed operands
loss2, _ = model(db)
print("loss after one step", loss2.item(), "(before", loss.item(), ")")
