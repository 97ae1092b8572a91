"""Step-by-step check of the TRAIN-mode CharResNet forward (each stage against torch on the CUDA path's own inputs)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from realise_b200.model import SpellBertPho2ResArch3Abla
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
from realise_b200.train import TrainEngine

torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False


def unsplit(x, n, S, C):   # parity-split rows [img][h&1][w&1][h/2][w/2][C] -> [n, C, S, S]
    h = S // 2
    return x.view(n, 2, 2, h, h, C).permute(0, 5, 3, 1, 4, 2).reshape(n, C, S, S)


def plain(x, n, S, C):     # rows (img, h, w) -> [n, C, S, S]
    return x.view(n, S, S, C).permute(0, 3, 1, 2)


def run(B, L):
    cfg = ArchConfig(num_hidden_layers=1, with_pho="no", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sd = synth_state_dict(cfg, 0)
    m = SpellBertPho2ResArch3Abla(cfg); m.tie_cls_weight(); m.load_state_dict(sd, strict=True); m.train().cuda()
    eng = m._engine = TrainEngine(m)
    batch = synth_batch(B, L, seed=99)
    db = {k: (v.cuda().contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}
    m.prepare()
    eng.forward(db)
    sv = eng.saved
    torch.cuda.synchronize()
    n = B * L
    ids = db["src_idx"].view(-1)
    x = m.char_images_multifonts.detach()[ids].float()          # [n, 3, 32, 32]
    print(f"--- N images = {n}")
    for bi, s in enumerate(sv["res"]):
        b = bi + 1
        blk = getattr(m.resnet, f"res_block{b}")
        conv1, bn1, _, conv2, bn2 = blk.residual_function
        convs, bns = blk.shortcut
        S, C = 32 >> b, [None, 64, 128, 256, 512, 768][b]
        xin = x if bi == 0 else (unsplit(sv["res"][bi - 1]["out"].float(), n, 2 * S, sv["res"][bi - 1]["out"].shape[1]))
        xin16 = xin.bfloat16().float()
        c1 = plain(s["c1"].float(), n, S, C); cs = plain(s["cs"].float(), n, S, C)
        c1r = F.conv2d(xin16, conv1.weight.detach().bfloat16().float(), stride=2, padding=1)
        csr = F.conv2d(xin16, convs.weight.detach().bfloat16().float(), stride=2)
        e = {"c1": float((c1 - c1r).abs().max()), "cs": float((cs - csr).abs().max())}
        mu, rstd = s["bn1"][2], s["bn1"][3]
        e["mean1"] = float((mu - c1.mean((0, 2, 3))).abs().max()); e["rstd1"] = float((rstd - (c1.var((0, 2, 3), unbiased=False) + 1e-5).rsqrt()).abs().max() / rstd.abs().max())
        a1r = torch.relu((c1 - mu[None, :, None, None]) * (rstd * bn1.weight.detach())[None, :, None, None] + bn1.bias.detach()[None, :, None, None])
        a1 = plain(s["a1"].float(), n, S, C)
        e["a1"] = float((a1 - a1r).abs().max())
        c2 = plain(s["c2"].float(), n, S, C)
        c2r = F.conv2d(a1, conv2.weight.detach().bfloat16().float(), padding=1)
        e["c2"] = float((c2 - c2r).abs().max())
        mu2, r2 = s["bn2"][2], s["bn2"][3]; mus, rs = s["bns"][2], s["bns"][3]
        e["mean2"] = float((mu2 - c2.mean((0, 2, 3))).abs().max()); e["means"] = float((mus - cs.mean((0, 2, 3))).abs().max())
        outr = torch.relu((c2 - mu2[None, :, None, None]) * (r2 * bn2.weight.detach())[None, :, None, None] + bn2.bias.detach()[None, :, None, None]
                          + (cs - mus[None, :, None, None]) * (rs * bns.weight.detach())[None, :, None, None] + bns.bias.detach()[None, :, None, None])
        out = s["out"].float()
        out = unsplit(out, n, S, C) if S >= 2 else out.view(n, C, 1, 1)
        e["out"] = float((out - outr).abs().max())
        e["|c1|"] = float(c1r.abs().max()); e["|out|"] = float(outr.abs().max())
        print(f"block{b} S={S} C={C}:", {k: round(v, 4) for k, v in e.items()}, flush=True)


run(4, 32)
run(16, 128)
run(128, 128)
