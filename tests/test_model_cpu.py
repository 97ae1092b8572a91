"""CPU: host-side contract of the drop-in model class (no compute)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from realise_b200.model import MODEL_CLASSES, SpellBertPho2ResArch3, SpellBertPho2ResArch3Abla
from realise_b200.synth import ArchConfig, state_dict_spec, synth_batch


@pytest.fixture(scope="module")
def model():
    m = SpellBertPho2ResArch3(ArchConfig())
    m.tie_cls_weight()
    return m


def test_state_dict_keys_and_shapes_match_reference(model):
    ref = json.load(open(os.path.join(GOLDEN, "arch3_state_dict_keys.json")))
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _ in ref]
    for k, shp in ref:
        assert list(sd[k].shape) == shp, k
    assert sorted((k, list(s)) for k, s in state_dict_spec(ArchConfig())) == sorted((k, s) for k, s in ref)


def test_tied_classifier_and_frozen_glyphs(model):
    assert model.classifier.weight is model.bert.embeddings.word_embeddings.weight
    assert not model.char_images_multifonts.requires_grad
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert n_train == 204016523  # SURVEY.md §2.2 (probed on the reference)


def test_no_cpu_fallback(model):
    batch = synth_batch(2, 16)
    model.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(batch)


def test_ablation_configs_build():
    for wp, wr, fu, g in [("no", "no", "gate", 1), ("yes", "no", "gate", 2), ("no", "yes", "gate", 2)]:
        m = SpellBertPho2ResArch3Abla(ArchConfig(with_pho=wp, with_res=wr, fusion=fu, num_hidden_layers=1))
        assert tuple(m.gate_net.weight.shape) == (g, (g + 1) * 768)
    assert set(MODEL_CLASSES) >= {"bert-pho2-res-arch3"}


def test_synth_batch_shape_contract():
    b = synth_batch(4, 32, seed=3)
    assert b["src_idx"].shape == (4, 32) and b["pho_idx"].shape[0] == 128 and len(b["pho_lens"]) == 128
    assert (b["src_idx"][:, 0] == 101).all()
    lens = torch.tensor(b["pho_lens"])
    assert lens.min() >= 1 and lens.max() <= 7
    special = (b["src_idx"].view(-1) == 0) | (b["src_idx"].view(-1) == 101) | (b["src_idx"].view(-1) == 102)
    assert (lens[special] == 1).all() and (b["pho_idx"][special, 0] == 32).all()
    assert ((b["pho_idx"] != 0).sum(1) == lens).all()


def test_flat_gradient_layout_on_cpu():
    """TrainEngine's flat gradient buffer (host logic, no kernels): every trainable parameter that the reference gives
    a gradient owns a 256-byte-aligned, non-overlapping slice; a layer's q/k/v weights (biases) are adjacent so the
    fused [3H, H] weight gradient is one view; poolers and unused word embeddings own none (grad stays None)."""
    from realise_b200.train import TrainEngine
    m = SpellBertPho2ResArch3Abla(ArchConfig(num_hidden_layers=1))
    m.tie_cls_weight()
    eng = TrainEngine(m)
    base = eng.flat.data_ptr()
    spans = []
    for p, g in zip(eng.params, eng.grads):
        if g is None:
            continue
        assert g.shape == p.shape and g.dtype == torch.float32
        off = g.data_ptr() - base
        assert 0 <= off and off + g.numel() * 4 <= eng.flat.numel() * 4
        spans.append((off, off + g.numel() * 4))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))              # disjoint
    names = {id(p): n for n, p in m.named_parameters()}
    no_grad = sorted(names[id(p)] for p, g in zip(eng.params, eng.grads) if g is None)
    assert no_grad == sorted(["bert.pooler.dense.weight", "bert.pooler.dense.bias", "pho_model.pooler.dense.weight",
                              "pho_model.pooler.dense.bias", "pho_model.embeddings.word_embeddings.weight",
                              "output_block.pooler.dense.weight", "output_block.pooler.dense.bias",
                              "output_block.embeddings.word_embeddings.weight"])
    att = m.bert.encoder.layer[0].attention
    w, b = eng._qkv_grads(att, 768)
    assert w.shape == (2304, 768) and w.data_ptr() == eng._grad(att.self.query.weight).data_ptr()
    assert eng._grad(att.self.key.weight).data_ptr() == w.data_ptr() + 768 * 768 * 4
    assert eng._grad(att.self.value.bias).data_ptr() == b.data_ptr() + 2 * 768 * 4
    assert (w.data_ptr() - base) % 256 == 0 and (b.data_ptr() - base) % 256 == 0


def test_flat_buffer_buckets_follow_backward_completion_order():
    """The flat gradient buffer is laid out in the order the backward completes it, in three contiguous buckets that data
    parallelism all-reduces behind the remaining backward: [output_block, gate, pinyin, glyph | upper bert layers |
    lower bert layers + embeddings (the tied classifier matrix is finished last, by the embedding scatter)]."""
    from realise_b200.train import TrainEngine
    m = SpellBertPho2ResArch3Abla(ArchConfig(num_hidden_layers=4))
    m.tie_cls_weight()
    eng = TrainEngine(m)
    assert len(eng.buckets) == 3 and eng.buckets[0][0] == 0 and eng.buckets[-1][1] == eng.flat.numel()
    assert all(a[1] == b[0] for a, b in zip(eng.buckets, eng.buckets[1:])) and eng.bert_split == 2
    base = eng.flat.data_ptr()

    def bucket_of(p):
        off = (eng._grad(p).data_ptr() - base) // 4
        return next(k for k, (a, b) in enumerate(eng.buckets) if a <= off < b)

    named = dict(m.named_parameters())
    for n, p in named.items():
        if not p.requires_grad or eng.grads[eng.index[id(p)]] is None:
            continue
        k = bucket_of(p)
        if n.startswith(("output_block.", "gate_net.", "pho_", "resnet", "classifier.bias")):
            assert k == 0, n
        elif n.startswith(("bert.encoder.layer.2.", "bert.encoder.layer.3.")):
            assert k == 1, n
        elif n.startswith("bert."):
            assert k == 2, n
    assert bucket_of(m.classifier.weight) == 2              # tied to bert.embeddings.word_embeddings
    # within bucket 1 the later layer comes first (it is differentiated first)
    l3 = eng._grad(named["bert.encoder.layer.3.output.dense.weight"]).data_ptr()
    l2 = eng._grad(named["bert.encoder.layer.2.output.dense.weight"]).data_ptr()
    assert l3 < l2
