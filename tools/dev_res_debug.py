import sys
import torch
sys.path.insert(0, ".")
from oracle import realise_oracle as O
from realise_b200.model import SpellBertPho2ResArch3Abla
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
from realise_b200.train import TrainEngine


def unsplit(x, n, S, C):
    h = S // 2
    return x.view(n, 2, 2, h, h, C).permute(0, 5, 3, 1, 4, 2).reshape(n, C, S, S)


cfg = ArchConfig(num_hidden_layers=1, with_pho="no", with_res="yes", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
sd = synth_state_dict(cfg, 13)
m = SpellBertPho2ResArch3Abla(cfg)
m.tie_cls_weight()
m.load_state_dict(sd, strict=True)
m.train().cuda()
m._engine = TrainEngine(m)
m._engine.debug = {}
batch = synth_batch(2, 16, seed=9)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
loss, _ = m(db)
loss.backward()
torch.cuda.synchronize()
dbg = m._engine.debug
rsd = {k: v.clone() for k, v in sd.items()}
rsd["classifier.weight"] = rsd["bert.embeddings.word_embeddings.weight"]
for k, v in rsd.items():
    if v.dtype.is_floating_point and "running" not in k and not k.startswith("char_images"):
        v.requires_grad_(True)
col = {}
# patch char_resnet to retain grads of block outputs
orig = O.basic_block
outs = {}


def bb(sd_, p, x, train, stats=None):
    y = orig(sd_, p, x, train, stats)
    y.retain_grad()
    outs[p] = y
    return y


O.basic_block = bb
n = 32


def gates(site, shape):
    b = int(site.split("res_block")[1][0])
    S, C = shape[-1], shape[1]
    if site.endswith(".a1"):
        t = dbg[f"a1_{b}"].float().cpu().view(n, S, S, C).permute(0, 3, 1, 2)
    else:
        t = dbg[f"out{b}"].float().cpu()
        t = unsplit(t, n, S, C) if S >= 2 else t.view(n, C, 1, 1)
    return (t > 0).float()


if "--gates" in sys.argv:
    O.RELU_MASK_FN = gates
rloss, _ = O.forward(rsd, batch, cfg, train=True, collect=col)
rloss.backward()
for b in range(5, 0, -1):
    y = outs[f"resnet.res_block{b}"]
    S, C = y.shape[-1], y.shape[1]
    g_ref, o_ref = y.grad, y.detach()
    g, o = dbg[f"dout{b}"].float().cpu(), dbg[f"out{b}"].float().cpu()
    if S >= 2:
        g, o = unsplit(g, n, S, C), unsplit(o, n, S, C)
    else:
        g, o = g.view(n, C, 1, 1), o.view(n, C, 1, 1)
    print(f"block{b}: out rel err {(o - o_ref).norm() / o_ref.norm():.3e}; dout rel err {(g - g_ref).norm() / g_ref.norm():.3e}; "
          f"|dout| {g_ref.norm():.3e}; relu agree {((o > 0) == (o_ref > 0)).float().mean():.4f}")

leaves = {k: v for k, v in rsd.items() if v.requires_grad}
worst = []
for name, p in m.named_parameters():
    if name == "classifier.weight" or p.grad is None or not name.startswith("resnet"):
        continue
    rg = leaves[name].grad
    worst.append(((p.grad.float().cpu() - rg).norm().item() / (rg.norm().item() + 1e-12), name))
worst.sort(reverse=True)
print("resnet param grads: max rel", worst[0], "median", worst[len(worst) // 2][0])
