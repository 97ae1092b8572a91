"""Loads the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py; falls back to /root/reference in
the build container) and builds its `SpellBertPho2ResArch3` through the reference's own public API.  Used ONLY by
bench.py's `--impl reference` arm and `incumbent` leg — never by realise_b200/ (the product path) or the tests.

Recipe: the third-party modules the reference imports but never uses on this path are stubbed,
and the vendored transformers 2.2.2 is put ahead of site-packages (the image also holds HF transformers 5.x).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = None


def locate():
    for root, src in ((os.path.join(HERE, "_ref"), os.path.join(HERE, "_ref", "src")), ("/root/reference", "/root/reference/src")):
        if os.path.isfile(os.path.join(src, "models.py")) and os.path.isdir(os.path.join(root, "transformers")):
            return root, src
    return None, None


def available():
    return locate()[0] is not None


def import_reference():
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    root, src = locate()
    if root is None:
        raise RuntimeError("reference not installed: run baseline/install_ref.py in the build container")
    for name in ["torchcrf", "pypinyin", "opencc", "boto3", "botocore", "botocore.exceptions", "botocore.config", "sacremoses"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchcrf"].CRF = object
    sys.modules["pypinyin"].Style = type("S", (), {"TONE3": 8})
    sys.modules["pypinyin"].pinyin = lambda *a, **k: [["U"]]
    sys.modules["opencc"].OpenCC = lambda *a, **k: None
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore.config"].Config = object
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.modules["botocore"].config = sys.modules["botocore.config"]
    for m in [k for k in sys.modules if k == "transformers" or k.startswith("transformers.")]:
        del sys.modules[m]
    sys.path.insert(0, src)
    sys.path.insert(0, root)
    import transformers
    assert transformers.__version__ == "2.2.2", transformers.__version__
    from transformers import AdamW, BertConfig, get_linear_schedule_with_warmup
    import models
    _CACHE = {"root": root, "BertConfig": BertConfig, "AdamW": AdamW, "schedule": get_linear_schedule_with_warmup,
              "models": models}
    return _CACHE


def build(state_dict, num_fonts=3, device="cpu"):
    """The reference model with the given (synthetic) weights, constructed exactly like src/run.py:417-431 does:
    config -> model class -> tie_cls_weight()."""
    R = import_reference()
    cfg = R["BertConfig"](vocab_size_or_config_json_file=21128)
    cfg.image_model_type, cfg.num_fonts = 0, num_fonts
    model = R["models"].SpellBertPho2ResArch3(cfg)
    model.tie_cls_weight()
    res = model.load_state_dict(state_dict, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.to(device)


class TrainStep:
    """The body of the reference's training loop (src/run.py:186-212) on one fixed batch: forward, backward,
    clip_grad_norm_(1.0), the vendored AdamW (grouped as src/run.py:146-152) + linear warm-up schedule, zero_grad."""

    def __init__(self, model, batch, autocast_dtype=None, lr=5e-5):
        import torch
        R = import_reference()
        self.torch, self.model, self.batch, self.autocast_dtype = torch, model.train(), batch, autocast_dtype
        no_decay = ["bias", "LayerNorm.weight"]
        groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 0.0},
                  {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
        self.opt = R["AdamW"](groups, lr=lr, eps=1e-8)
        self.sched = R["schedule"](self.opt, num_warmup_steps=10000, num_training_steps=1000000)

    def __call__(self):
        torch = self.torch
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self.model(self.batch)[0]
        else:
            loss = self.model(self.batch)[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
        self.opt.step()
        self.sched.step()
        self.model.zero_grad()
        return loss
