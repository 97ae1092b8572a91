"""BASELINE configs[4]: ablation sweep on 1 GPU — semantic-only / +glyph / +pinyin / full (src/models_abla.py) at
seq_len 64 / 128 / 256.  Forward-only (eval, CUDA graph, fp16 operands) for every cell; the training step (fwd + bwd +
clip + AdamW as one CUDA-graph replay) for every cell too (batch 64 at seq_len 256).
Prints one JSON object per cell and a summary table; random-init weights, synthetic batches (realise_b200.synth)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from realise_b200.graphed import GraphedTrainStep  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402
from realise_b200.optim import FusedAdamW  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch  # noqa: E402

dev = torch.device("cuda", 0)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rows = []
for name, pho, res in (("semantic-only", "no", "no"), ("+glyph", "no", "yes"), ("+pinyin", "yes", "no"), ("full", "yes", "yes")):
    cfg = ArchConfig(with_pho=pho, with_res=res)
    torch.manual_seed(0)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.to(dev)
    opt = None
    for L in (64, 128, 256):
        B = 128 if L <= 128 else 64
        host = synth_batch(B, L, seed=7, ragged=False, with_labels=True)
        db = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
        db["pho_lens"] = torch.tensor(host["pho_lens"], dtype=torch.int32, device=dev)
        cell = {"variant": name, "with_pho": pho, "with_res": res, "seq_len": L, "batch": B}
        # ---- forward-only ----
        model.eval()
        fb = {k: v for k, v in db.items() if k not in ("tgt_idx", "loss_masks")}
        with torch.no_grad():
            for _ in range(3):
                model(fb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(STEPS):
                model(fb)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / STEPS
        cell["fwd_ms"] = round(ms, 3)
        cell["fwd_sentences_per_s"] = round(B / ms * 1e3, 1)
        # ---- train step ----
        if True:
            model.train()
            if opt is None:
                opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=5e-5, max_grad_norm=1.0, model=model)
            step = GraphedTrainStep(model, opt)
            for _ in range(3):
                step(db)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(STEPS):
                step(db)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / STEPS
            cell["train_ms"] = round(ms, 3)
            cell["train_sentences_per_s"] = round(B / ms * 1e3, 1)
            del step
            model._engine = None
        torch.cuda.empty_cache()
        rows.append(cell)
        print(json.dumps(cell), flush=True)
    del model, opt
    torch.cuda.empty_cache()
print("| variant | seq_len | batch | fwd ms | fwd sent/s | train ms | train sent/s |")
print("|---|---|---|---|---|---|---|")
for c in rows:
    print(f"| {c['variant']} | {c['seq_len']} | {c['batch']} | {c['fwd_ms']} | {c['fwd_sentences_per_s']} | {c.get('train_ms')} | "
          f"{c.get('train_sentences_per_s')} |")
