"""CPU: the drop-in boundary around the model class against the REFERENCE ITSELF where it
can be imported (/root/reference exists in the build container only; those tests skip elsewhere):

  * build_glyce_embed / build_glyce_embed_multifonts vs the reference's rasteriser on simhei.ttf / xiaozhuan.ttf;
  * checkpoints in the reference's file layout (config.json + pytorch_model.bin) cross-load in both directions;
  * the src/ shim: the reference's unchanged src/run.py and src/test.py resolve MODEL_CLASSES to realise_b200 classes,
    `--local-rank` is accepted, `from_pretrained(dir, config=<reference BertConfig>)` builds our model.
"""
import json
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from realise_b200.synth import ArchConfig, synth_state_dict

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")


_REF_MODULES = None


def import_reference():
    """(BertConfig, models, models_abla) of the reference — imported ONCE per process (a second import would re-create
    the vendored transformers classes under the already-imported `models`)."""
    global _REF_MODULES
    if _REF_MODULES is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        try:
            from make_golden import import_reference as imp
        finally:
            sys.path.pop(0)
        _REF_MODULES = imp()
    return _REF_MODULES


def write_vocab(d, n=21128, seed=0):
    """A vocab.txt of n entries shaped like bert-base-chinese's: specials, [unusedN], ASCII, punctuation, ##pieces and a
    few hundred CJK ideographs (the only entries the single-font builder draws)."""
    rng = np.random.default_rng(seed)
    vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(1, 100)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    vocab += list("!\"#$%&'()*+,-./0123456789:;<=>?@abcdefghijklmnopqrstuvwxyz~，。！？、")
    cjk = [chr(c) for c in rng.choice(np.arange(0x4E00, 0x9FA5), size=400, replace=False)] + list("的一是不了人我在有他这为之大来以个中上们")
    vocab += cjk + ["㐀", "豈", "龘"]
    vocab += [f"##p{i}" for i in range(n - len(vocab))]
    assert len(vocab) == n
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    return vocab


@needs_ref
def test_glyph_table_builders_match_the_reference(tmp_path):
    _, models, _ = import_reference()
    from realise_b200 import glyphs
    write_vocab(str(tmp_path))
    simhei, xiaozhuan = os.path.join(REF, "simhei.ttf"), os.path.join(REF, "xiaozhuan.ttf")
    # single font (src/models.py:703-733): only one-ideograph entries are drawn
    ref_self = types.SimpleNamespace(char_images=torch.nn.Embedding(21128, 1024))
    models.SpellBertPho2ResArch3.build_glyce_embed(ref_self, str(tmp_path), simhei)
    ours = types.SimpleNamespace(char_images=torch.nn.Embedding(21128, 1024), _invalidate=lambda: None)
    glyphs.build_glyce_embed(ours, str(tmp_path), simhei)
    assert torch.equal(ours.char_images.weight.data, ref_self.char_images.weight.data)
    assert float(ours.char_images.weight.data[0].abs().max()) > 0      # standardised zeros are not zero
    # one font of the multi-font builder (src/models.py:763-795): every single-character entry is drawn
    for font in (simhei, xiaozhuan):
        ref = models.SpellBertPho2ResArch3.build_glyce_embed_onefont(types.SimpleNamespace(), vocab_dir=str(tmp_path),
                                                                      font_path=font, font_size=32, use_traditional=False)
        got = torch.from_numpy(glyphs.rasterize(glyphs.read_vocab(str(tmp_path)), font, 32))
        assert torch.equal(got, ref)
    # the three-plane table, traditional plane through a stand-in converter on both sides
    s2t = {"这": "這", "为": "為", "来": "來", "个": "個", "们": "們"}
    conv = types.SimpleNamespace(convert=lambda c: s2t.get(c, c))
    cwd = os.getcwd()
    os.chdir(REF)                                   # the reference resolves 'simhei.ttf' against the working directory
    try:
        ref_self = types.SimpleNamespace(char_images_multifonts=torch.nn.Parameter(torch.zeros(21128, 3, 32, 32)), converter=conv)
        ref_self.build_glyce_embed_onefont = lambda **k: models.SpellBertPho2ResArch3.build_glyce_embed_onefont(ref_self, **k)
        sys.modules["opencc"].OpenCC = lambda *a, **k: conv
        models.SpellBertPho2ResArch3.build_glyce_embed_multifonts(ref_self, str(tmp_path), 3, True)
    finally:
        os.chdir(cwd)
    ours = types.SimpleNamespace(char_images_multifonts=torch.nn.Parameter(torch.zeros(21128, 3, 32, 32)),
                                 _invalidate=lambda: None)
    glyphs.build_glyce_embed_multifonts(ours, str(tmp_path), 3, True, font_dir=REF, s2t=conv.convert)
    assert torch.equal(ours.char_images_multifonts.data, ref_self.char_images_multifonts.data)
    assert glyphs.multifont_plan(2, True) == [("simhei.ttf", False), ("simhei.ttf", True)]
    assert glyphs.multifont_plan(3, False)[-1] == ("simhei.ttf", True)


def small_cfg():
    return ArchConfig(num_hidden_layers=1)


def test_save_and_from_pretrained_round_trip(tmp_path):
    """config.json + pytorch_model.bin (modeling_utils.py:236-251): what save_pretrained writes, from_pretrained reads
    back bit for bit, in eval mode, with the extra config attributes (src/run.py:421-425) preserved."""
    from realise_b200.model import SpellBertPho2ResArch3Abla
    cfg = small_cfg()
    cfg.with_pho = "no"
    m = SpellBertPho2ResArch3Abla(cfg)
    m.tie_cls_weight()
    m.load_state_dict(synth_state_dict(cfg, seed=4), strict=True)
    m.save_pretrained(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["config.json", "pytorch_model.bin"]
    m2 = SpellBertPho2ResArch3Abla.from_pretrained(str(tmp_path))
    assert not m2.training and m2.config.with_pho == "no" and m2.config.num_hidden_layers == 1
    assert m2.missing_keys == [] and m2.unexpected_keys == []
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    # a semantic-only checkpoint (bert-base-chinese style: `bert.*` + unrelated `cls.*` heads, old LayerNorm names)
    part = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v
            for k, v in sd.items() if k.startswith("bert.")}
    part["cls.predictions.bias"] = torch.zeros(3)
    d2 = tmp_path / "bert_only"
    os.makedirs(d2)
    torch.save(part, d2 / "pytorch_model.bin")
    shutil.copy(tmp_path / "config.json", d2 / "config.json")
    m3 = SpellBertPho2ResArch3Abla.from_pretrained(str(d2))
    assert m3.unexpected_keys == ["cls.predictions.bias"] and all(not k.startswith("bert.") for k in m3.missing_keys)
    assert torch.equal(m3.bert.encoder.layer[0].output.LayerNorm.weight, sd["bert.encoder.layer.0.output.LayerNorm.weight"])


@needs_ref
def test_checkpoints_cross_load_with_the_reference(tmp_path):
    BertConfig, models, _ = import_reference()
    from realise_b200.model import SpellBertPho2ResArch3
    cfg = small_cfg()
    ours = SpellBertPho2ResArch3(cfg)
    ours.tie_cls_weight()
    ours.load_state_dict(synth_state_dict(cfg, seed=5), strict=True)
    d1 = str(tmp_path / "ours")
    ours.save_pretrained(d1)
    rc = BertConfig.from_pretrained(d1)                                   # the reference parses OUR config.json
    assert rc.num_fonts == 3 and rc.image_model_type == 0 and rc.num_hidden_layers == 1 and rc.vocab_size == 21128
    ref = models.SpellBertPho2ResArch3.from_pretrained(d1, config=rc)     # ... and loads OUR pytorch_model.bin
    rsd, osd = ref.state_dict(), ours.state_dict()
    assert set(rsd) == set(osd) and all(torch.equal(rsd[k], osd[k]) for k in rsd)
    d2 = str(tmp_path / "ref")
    os.makedirs(d2)
    ref.save_pretrained(d2)                                               # the reference's own files ...
    back = SpellBertPho2ResArch3.from_pretrained(d2, config=BertConfig.from_pretrained(d2), cache_dir=None)
    bsd = back.state_dict()                                               # ... load into ours given a REFERENCE config object
    assert back.missing_keys == [] and back.unexpected_keys == []
    assert all(torch.equal(bsd[k], osd[k]) for k in osd)
    back2 = SpellBertPho2ResArch3.from_pretrained(d2)                     # and from the reference's config.json alone
    assert back2.config.num_hidden_layers == 1 and back2.config.hidden_dropout_prob == 0.1


def test_argv_and_torch_load_compat():
    from realise_b200 import compat
    assert compat.fix_argv(["src/run.py", "--local-rank=1", "--seed", "17"]) == ["src/run.py", "--local_rank=1", "--seed", "17"]
    assert compat.fix_argv(["src/run.py", "--local-rank", "1"]) == ["src/run.py", "--local_rank", "1"]
    assert compat.fix_argv(["src/run.py", "--seed", "1"], {"LOCAL_RANK": "3", "WORLD_SIZE": "8"})[-1] == "--local_rank=3"
    assert compat.fix_argv(["src/run.py", "--local_rank=0"], {"LOCAL_RANK": "3", "WORLD_SIZE": "8"}) == ["src/run.py", "--local_rank=0"]
    assert compat.fix_argv(["src/test.py", "--x"], {"LOCAL_RANK": "0", "WORLD_SIZE": "2"}) == ["src/test.py", "--x"]


@needs_ref
def test_reference_drivers_resolve_to_realise_b200_through_the_shim(tmp_path):
    """A reference checkout whose src/models.py and src/models_abla.py were replaced by shim/src/*: the UNCHANGED
    src/run.py (train.sh) and src/test.py (test.sh) import, their MODEL_CLASSES point at realise_b200 classes, run.py's
    argparse takes torchrun's `--local-rank`, training_args.bin unpickles, and `from_pretrained(dir, config=<reference
    BertConfig>)` + tie_cls_weight + build_glyce_embed_multifonts (src/run.py:417-440) produce our model."""
    src = tmp_path / "src"
    os.makedirs(src)
    for f in ("run.py", "test.py", "utils.py", "metric.py", "metric_core.py", "remove_de.py"):
        shutil.copy(os.path.join(REF, "src", f), src / f)
    for f in ("models.py", "models_abla.py"):
        shutil.copy(os.path.join(ROOT, "shim", "src", f), src / f)
    os.symlink(os.path.join(REF, "transformers"), tmp_path / "transformers")
    for f in ("simhei.ttf", "xiaozhuan.ttf"):
        os.symlink(os.path.join(REF, f), tmp_path / f)
    (tmp_path / "pypinyin.py").write_text("class Style:\n    TONE3 = 8\n\ndef pinyin(c, **k):\n    return [['U']]\n")
    (tmp_path / "opencc.py").write_text("class OpenCC:\n    def __init__(self, *a):\n        pass\n    def convert(self, c):\n        return c\n")
    # third-party packages of the reference's requirements.txt that this container lacks (none is on the arithmetic path)
    (tmp_path / "boto3.py").write_text("")
    (tmp_path / "sacremoses.py").write_text("")
    (tmp_path / "torchcrf.py").write_text("CRF = object\n")
    os.makedirs(tmp_path / "botocore")
    (tmp_path / "botocore" / "__init__.py").write_text("")
    (tmp_path / "botocore" / "exceptions.py").write_text("ClientError = Exception\n")
    (tmp_path / "botocore" / "config.py").write_text("Config = object\n")
    pre = tmp_path / "pretrained"
    write_vocab(str(pre))
    cfg = small_cfg()
    with open(pre / "config.json", "w") as f:
        json.dump({k: v for k, v in cfg.__dict__.items() if k not in ("with_pho", "with_res", "fusion", "num_fonts",
                                                                      "image_model_type")}, f)
    torch.save({k: v for k, v in synth_state_dict(cfg, seed=6).items() if k.startswith("bert.")}, pre / "pytorch_model.bin")
    code = r'''
import sys, os, argparse, torch
sys.argv = ["src/run.py", "--local-rank=0", "--model_type", "bert-pho2-res-arch3"]
import run, test as ref_test
cls = run.MODEL_CLASSES["bert-pho2-res-arch3"][1]
assert cls.__module__ == "realise_b200.model" and cls.__name__ == "SpellBertPho2ResArch3", cls
assert run.MODEL_CLASSES["bert-pho2-res-arch3-abla"][1].__module__ == "realise_b200.model"
assert ref_test.MODEL_CLASSES["bert-pho2-res-arch3"] is cls
assert sys.argv[1] == "--local_rank=0"
torch.save(argparse.Namespace(model_type="bert-pho2-res-arch3"), "training_args.bin")
assert torch.load("training_args.bin").model_type == "bert-pho2-res-arch3"          # src/test.py:105 as written
config_class, model_class, tokenizer_class = run.MODEL_CLASSES["bert-pho2-res-arch3"]
config = config_class.from_pretrained("pretrained", image_model_type=0, cache_dir=None)   # src/run.py:418-425
config.image_model_type, config.num_fonts, config.with_pho, config.with_res, config.fusion = 0, 3, "yes", "yes", "gate"
tokenizer = tokenizer_class.from_pretrained("pretrained", do_lower_case=False, cache_dir=None)
model = model_class.from_pretrained("pretrained", config=config, cache_dir=None)
model.tie_cls_weight()
model.build_glyce_embed_multifonts("pretrained", 3, True)
assert model.classifier.weight is model.bert.embeddings.word_embeddings.weight
assert float(model.char_images_multifonts.std()) > 0.5 and not model.char_images_multifonts.requires_grad
batch = {"src_idx": torch.tensor([[101, 250, 251, 102]])}
batch = model_class.build_batch(batch, tokenizer)                                          # src/run.py:443 -> :100
assert batch["pho_idx"].shape[0] == 4 and len(batch["pho_lens"]) == 4
try:
    run.MODEL_CLASSES["bert"][1].from_pretrained("pretrained")
    raise SystemExit("out-of-scope class did not raise")
except NotImplementedError:
    pass
named = [n for n, p in model.named_parameters() if not any(nd in n for nd in ["bias", "LayerNorm.weight"])]   # :146-151
assert "gate_net.weight" in named and "bert.embeddings.LayerNorm.weight" not in named
print("SHIM_OK")
'''
    env = dict(os.environ, REALISE_B200_ROOT=ROOT, PYTHONPATH=f"{tmp_path}:{tmp_path / 'src'}")
    p = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert p.returncode == 0 and "SHIM_OK" in p.stdout, p.stdout[-3000:]
