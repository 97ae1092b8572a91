"""CPU: the oracle restatement against outputs of the reference itself (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from conftest import cached_state_dict, load_golden
from oracle import realise_oracle as O
from realise_b200.synth import ArchConfig, synth_batch

CASES = [
    "arch3_eval_B2_L16.npz",
    "arch3_eval_B3_L40.npz",
    "abla_pho-no_res-no_gate_B2_L16.npz",
    "abla_pho-yes_res-no_gate_B2_L16.npz",
    "abla_pho-no_res-yes_gate_B2_L16.npz",
    "abla_pho-yes_res-yes_sum_B2_L16.npz",
]
TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, fast, monkeypatch):
    monkeypatch.setattr(O, "FAST", fast)
    g, meta = load_golden(name)
    cfg = ArchConfig(**meta["cfg"])
    sd = cached_state_dict(cfg, meta["wseed"])
    batch = synth_batch(meta["B"], meta["L"], seed=meta["bseed"])
    col = {}
    with torch.no_grad():
        loss, logits = O.forward(sd, batch, cfg, collect=col)
    n = meta["B"] * meta["L"]
    pairs = {"bert_hiddens": "bert_hiddens", "pho_gru": "pho_gru", "pho_hiddens": "pho_hiddens", "resnet": "resnet",
             "res_hiddens": "res_hiddens", "output_block": "sequence_output"}
    for gk, ck in pairs.items():
        if gk in g.files:
            a = col[ck].reshape(g[gk].shape).numpy()
            assert np.abs(a - g[gk]).max() <= TOL, gk
    for blk in ("res_block1", "res_block2"):
        if blk in g.files:
            assert np.abs(col[blk][:8].numpy() - g[blk]).max() <= TOL, blk
    flat = logits.reshape(n, -1)
    assert np.abs(flat[torch.from_numpy(g["logits_rows"])].numpy() - g["logits_kept"]).max() <= TOL
    assert np.abs(torch.logsumexp(flat, -1).numpy() - g["logits_lse"]).max() <= 1e-4
    am = flat.argmax(-1).numpy()
    safe = g["logits_top2_gap"] > 10 * TOL
    assert (am[safe] == g["logits_argmax"][safe]).all()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5


def test_oracle_train_mode_matches_reference_golden():
    """dropout p=0, BatchNorm in batch-statistics mode: loss, running-stat updates, gradients."""
    g, meta = load_golden("arch3_train_B2_L16.npz")
    cfg = ArchConfig(**meta["cfg"])
    cfg.hidden_dropout_prob = 0.0
    cfg.attention_probs_dropout_prob = 0.0
    sd = {k: v.clone() for k, v in cached_state_dict(ArchConfig(**meta["cfg"]), meta["wseed"]).items()}
    sd["classifier.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    params = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k and not k.startswith("char_images"):
            v.requires_grad_(True)
            params[k] = v
    batch = synth_batch(meta["B"], meta["L"], seed=meta["bseed"])
    stats = {}
    loss, logits = O.forward(sd, batch, cfg, train=True, bn_stats=stats)
    assert abs(loss.item() - float(g["loss"])) <= 2e-5
    for name, ref in zip(g["bn_names"], g["bn_sums"]):
        key = "resnet." + str(name).split("resnet.", 1)[1]
        assert abs(stats[key].double().sum().item() - ref) <= 1e-3 * max(1.0, abs(ref)), key
    loss.backward()
    names = [str(x) for x in g["grad_names"]]
    for name, ref in zip(names, g["grad_stats"]):
        if name == "classifier.weight":
            continue  # tied: the gradient lives on bert.embeddings.word_embeddings.weight
        gr = params[name].grad
        assert gr is not None, name
        assert abs(gr.norm().item() - ref[0]) <= 2e-4 * max(1.0, ref[0]), name
