"""Localise a train-mode forward discrepancy: per-branch error of the CUDA train forward vs the oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import realise_oracle as O
from realise_b200 import ops
from realise_b200.model import SpellBertPho2ResArch3Abla
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
from realise_b200.train import TrainEngine


def run(B, L, layers, p, ragged=True, seed=4242):
    cfg = ArchConfig(num_hidden_layers=layers, hidden_dropout_prob=p, attention_probs_dropout_prob=p)
    sd = synth_state_dict(cfg, 0)
    m = SpellBertPho2ResArch3Abla(cfg); m.tie_cls_weight(); m.load_state_dict(sd, strict=True); m.train().cuda()
    eng = m._engine = TrainEngine(m); eng.set_seed(seed)
    batch = synth_batch(B, L, seed=99, ragged=ragged)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    db["pho_lens"] = torch.tensor(batch["pho_lens"], dtype=torch.int32, device="cuda")
    m.prepare()
    loss, logits = eng.forward({k: v.contiguous() for k, v in db.items()})
    sv = eng.saved
    torch.cuda.synchronize()
    def mask_fn(site, shape):
        return ops.dropout_mask(int(np.prod(shape)), p, seed, site).reshape(shape).float()
    rsd = {k: v.cuda() for k, v in sd.items()}
    ob = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    col = {}
    O.MASK_FN = mask_fn if p > 0 else None
    O.FAST = True
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        rloss, rlogits = O.forward(rsd, ob, cfg, train=True, collect=col)
    O.MASK_FN = None
    names = ["bert_hiddens", "pho_hiddens", "res_hiddens"]
    out = {}
    for n, t in zip(names, sv["mods"]):
        out[n] = float((t.view(B, L, -1) - col[n]).abs().max())
    out["pho_gru"] = float((sv["gru"]["hs"][-1].view(B, L, -1) - col["pho_gru"]).abs().max())
    out["resnet_raw"] = float((sv["res_raw"].view(B, L, -1) - col["resnet"]).abs().max())
    out["logits"] = float((logits - rlogits).abs().max())
    out["loss"] = (float(loss), float(rloss))
    print(f"B={B} L={L} layers={layers} p={p} ragged={ragged}:", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in out.items()}, flush=True)


run(4, 32, 2, 0.0)
run(16, 128, 2, 0.0)
run(16, 128, 2, 0.1)
run(16, 128, 12, 0.0)
run(16, 128, 12, 0.1)
run(16, 128, 12, 0.1, ragged=False)
