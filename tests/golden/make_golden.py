"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (imported from /root/reference in
the build container; it does not exist on the GPU box) on realise_b200.synth weights/batches.

    python tests/golden/make_golden.py            # writes arch3_B2_L16.npz, abla_*.npz, keys.json

Import recipe: SURVEY.md Appendix B (stub the non-arithmetic third-party modules, put the vendored
transformers 2.2.2 ahead of site-packages).  Outputs are stored as float32; the big [B,L,V] logits
are kept for a few tokens only, plus per-token argmax / max / logsumexp over the full vocabulary.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402

REF = "/root/reference"


def import_reference():
    for name in ["torchcrf", "pypinyin", "opencc", "boto3", "botocore", "botocore.exceptions", "botocore.config",
                 "sacremoses"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchcrf"].CRF = object
    sys.modules["pypinyin"].Style = type("S", (), {"TONE3": 8})
    sys.modules["pypinyin"].pinyin = lambda *a, **k: [["U"]]
    sys.modules["opencc"].OpenCC = lambda *a, **k: None
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore.config"].Config = object
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.modules["botocore"].config = sys.modules["botocore.config"]
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.path.insert(0, REF)
    for m in [k for k in sys.modules if k == "transformers" or k.startswith("transformers.")]:
        del sys.modules[m]
    from transformers import BertConfig  # vendored 2.2.2
    import models
    import models_abla
    return BertConfig, models, models_abla


def ref_config(BertConfig, cfg: ArchConfig):
    rc = BertConfig(vocab_size_or_config_json_file=cfg.vocab_size, num_hidden_layers=cfg.num_hidden_layers)
    rc.image_model_type = cfg.image_model_type
    rc.num_fonts = cfg.num_fonts
    rc.with_pho, rc.with_res, rc.fusion = cfg.with_pho, cfg.with_res, cfg.fusion
    return rc


def capture(model, names):
    store, hooks = {}, []
    for name, mod in names.items():
        def hook(m, inp, out, name=name):
            o = out[0] if isinstance(out, tuple) else out
            if name == "pho_gru":
                o = out[1].squeeze(0)
            if name.startswith("res_block"):
                o = o[:8]                       # keep the fixture small: first 8 glyph images only
            store[name] = o.detach().float().clone()
        hooks.append(mod.register_forward_hook(hook))
    return store, hooks


def run_case(BertConfig, model_cls, cfg, B, L, wseed, bseed, train, out_path, keep_tokens=6):
    torch.manual_seed(0)
    model = model_cls(ref_config(BertConfig, cfg))
    model.tie_cls_weight()
    sd = synth_state_dict(cfg, seed=wseed)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if train:
        model.train()
        for m in model.modules():              # parity protocol: dropout off, BN in batch-stat mode
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    else:
        model.eval()
    batch = synth_batch(B, L, seed=bseed, ragged=True)
    names = {"bert_hiddens": model.bert, "output_block": model.output_block}
    if cfg.with_pho == "yes":
        names.update({"pho_gru": model.pho_gru, "pho_hiddens": model.pho_model})
    if cfg.with_res == "yes":
        names.update({"resnet": model.resnet, "res_hiddens": model.resnet_layernorm,
                      "res_block1": model.resnet.res_block1, "res_block2": model.resnet.res_block2})
    store, hooks = capture(model, names)
    with torch.set_grad_enabled(train):
        loss, logits = model(batch)
    out = {k: v.numpy() for k, v in store.items()}
    flat = logits.detach().reshape(B * L, -1)
    out["loss"] = np.float32(loss.item())
    out["logits_argmax"] = flat.argmax(-1).numpy().astype(np.int64)
    out["logits_max"] = flat.max(-1).values.numpy()
    out["logits_lse"] = torch.logsumexp(flat, -1).numpy()
    top2 = flat.topk(2, dim=-1).values
    out["logits_top2_gap"] = (top2[:, 0] - top2[:, 1]).numpy()
    keep = np.linspace(0, B * L - 1, keep_tokens).astype(np.int64)
    out["logits_rows"] = keep
    out["logits_kept"] = flat[keep].numpy()
    if train:
        loss.backward()
        grads = {}
        for n, p in model.named_parameters():
            if p.grad is not None:
                g = p.grad.detach().float()
                grads[n] = np.array([g.norm().item(), g.abs().max().item(), g.flatten()[:: max(1, g.numel() // 64)][:64].sum().item()],
                                    dtype=np.float64)
        out["grad_names"] = np.array(sorted(grads))
        out["grad_stats"] = np.stack([grads[n] for n in sorted(grads)])
        bn = {n: b.detach().clone() for n, b in model.named_buffers() if "running" in n}
        out["bn_names"] = np.array(sorted(bn))
        out["bn_sums"] = np.array([bn[n].double().sum().item() for n in sorted(bn)])
    out["meta"] = np.array(json.dumps({"B": B, "L": L, "wseed": wseed, "bseed": bseed, "train": train,
                                       "cfg": cfg.__dict__}))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: getattr(v, "shape", None) for k, v in out.items() if k != "meta"}, "loss", out["loss"])
    return model


if __name__ == "__main__":
    BertConfig, models, models_abla = import_reference()
    cfg = ArchConfig()
    m = run_case(BertConfig, models.SpellBertPho2ResArch3, cfg, 2, 16, 0, 1234, False,
                 os.path.join(HERE, "arch3_eval_B2_L16.npz"))
    keys = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "arch3_state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0)
    del m
    run_case(BertConfig, models.SpellBertPho2ResArch3, cfg, 3, 40, 1, 77, False,
             os.path.join(HERE, "arch3_eval_B3_L40.npz"))
    run_case(BertConfig, models.SpellBertPho2ResArch3, cfg, 2, 16, 0, 1234, True,
             os.path.join(HERE, "arch3_train_B2_L16.npz"))
    for wp, wr, fu in [("no", "no", "gate"), ("yes", "no", "gate"), ("no", "yes", "gate"), ("yes", "yes", "sum")]:
        c = ArchConfig(with_pho=wp, with_res=wr, fusion=fu)
        run_case(BertConfig, models_abla.SpellBertPho2ResArch3Abla, c, 2, 16, 0, 1234, False,
                 os.path.join(HERE, f"abla_pho-{wp}_res-{wr}_{fu}_B2_L16.npz"))
