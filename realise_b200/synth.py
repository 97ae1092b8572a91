"""Deterministic synthetic weights and batches for SpellBertPho2ResArch3.

Weights are produced per state_dict key from numpy PCG64 streams seeded by crc32(key) ^ seed, so the
same tensors can be regenerated bit-for-bit wherever they are needed (golden generation against the
reference in the build container, the oracle, the CUDA model on the GPU box) without shipping 1 GB
of parameters.  Key names / shapes follow the reference state_dict (src/models.py:654-698,
src/char_cnn.py:9-45, transformers/modeling_bert.py:155-415).
"""
import math
import zlib
from dataclasses import dataclass

import numpy as np
import torch

CLS, SEP, PAD = 101, 102, 0
PHO_VOCAB = 33  # src/utils.py:61-67  ('P', '1'..'5', 'a'..'z', 'U')


@dataclass
class ArchConfig:
    vocab_size: int = 21128
    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    num_hidden_layers: int = 12      # `bert`; pho_model is always 4, output_block always 3
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    num_fonts: int = 3
    image_model_type: int = 0
    with_pho: str = "yes"
    with_res: str = "yes"
    fusion: str = "gate"

    @property
    def num_gates(self):
        return 1 + (self.with_pho == "yes") + (self.with_res == "yes")


RES_CHANNELS = [None, 64, 128, 256, 512, 768]


def _bert_keys(prefix, n_layers, cfg):
    H, I = cfg.hidden_size, cfg.intermediate_size
    out = [
        (f"{prefix}.embeddings.word_embeddings.weight", (cfg.vocab_size, H)),
        (f"{prefix}.embeddings.position_embeddings.weight", (cfg.max_position_embeddings, H)),
        (f"{prefix}.embeddings.token_type_embeddings.weight", (cfg.type_vocab_size, H)),
        (f"{prefix}.embeddings.LayerNorm.weight", (H,)),
        (f"{prefix}.embeddings.LayerNorm.bias", (H,)),
    ]
    for i in range(n_layers):
        p = f"{prefix}.encoder.layer.{i}"
        out += [
            (f"{p}.attention.self.query.weight", (H, H)), (f"{p}.attention.self.query.bias", (H,)),
            (f"{p}.attention.self.key.weight", (H, H)), (f"{p}.attention.self.key.bias", (H,)),
            (f"{p}.attention.self.value.weight", (H, H)), (f"{p}.attention.self.value.bias", (H,)),
            (f"{p}.attention.output.dense.weight", (H, H)), (f"{p}.attention.output.dense.bias", (H,)),
            (f"{p}.attention.output.LayerNorm.weight", (H,)), (f"{p}.attention.output.LayerNorm.bias", (H,)),
            (f"{p}.intermediate.dense.weight", (I, H)), (f"{p}.intermediate.dense.bias", (I,)),
            (f"{p}.output.dense.weight", (H, I)), (f"{p}.output.dense.bias", (H,)),
            (f"{p}.output.LayerNorm.weight", (H,)), (f"{p}.output.LayerNorm.bias", (H,)),
        ]
    out += [(f"{prefix}.pooler.dense.weight", (H, H)), (f"{prefix}.pooler.dense.bias", (H,))]
    return out


def _resnet_keys(cfg):
    out = []
    cin = cfg.num_fonts
    for b in range(1, 6):
        cout = RES_CHANNELS[b]
        p = f"resnet.res_block{b}"
        out.append((f"{p}.residual_function.0.weight", (cout, cin, 3, 3)))
        out += _bn_keys(f"{p}.residual_function.1", cout)
        out.append((f"{p}.residual_function.3.weight", (cout, cout, 3, 3)))
        out += _bn_keys(f"{p}.residual_function.4", cout)
        out.append((f"{p}.shortcut.0.weight", (cout, cin, 1, 1)))
        out += _bn_keys(f"{p}.shortcut.1", cout)
        cin = cout
    return out


def _bn_keys(p, c):
    return [(f"{p}.weight", (c,)), (f"{p}.bias", (c,)), (f"{p}.running_mean", (c,)),
            (f"{p}.running_var", (c,)), (f"{p}.num_batches_tracked", ())]


def state_dict_spec(cfg: ArchConfig):
    """[(key, shape)] in the reference's registration order (src/models.py:654-698 or
    src/models_abla.py:35-95 when a modality is switched off)."""
    H = cfg.hidden_size
    keys = _bert_keys("bert", cfg.num_hidden_layers, cfg)
    if cfg.with_pho == "yes":
        keys += [("pho_embeddings.weight", (PHO_VOCAB, H)),
                 ("pho_gru.weight_ih_l0", (3 * H, H)), ("pho_gru.weight_hh_l0", (3 * H, H)),
                 ("pho_gru.bias_ih_l0", (3 * H,)), ("pho_gru.bias_hh_l0", (3 * H,))]
        keys += _bert_keys("pho_model", 4, cfg)
    if cfg.with_res == "yes":
        if cfg.num_fonts == 1:
            keys += [("char_images.weight", (cfg.vocab_size, 1024))]
        else:
            keys += [("char_images_multifonts", (21128, cfg.num_fonts, 32, 32))]
        keys += _resnet_keys(cfg)
        keys += [("resnet_layernorm.weight", (H,)), ("resnet_layernorm.bias", (H,))]
    if cfg.fusion == "gate":
        g = cfg.num_gates
        keys += [("gate_net.weight", (g, (g + 1) * H)), ("gate_net.bias", (g,))]
    keys += _bert_keys("output_block", 3, cfg)
    keys += [("classifier.weight", (cfg.vocab_size, H)), ("classifier.bias", (cfg.vocab_size,))]
    return keys


def _gen(key, shape, seed):
    rng = np.random.Generator(np.random.PCG64((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF))
    n = int(np.prod(shape)) if len(shape) else 1

    def normal(std):
        return (rng.standard_normal(n, dtype=np.float32) * np.float32(std)).reshape(shape)

    def uniform(lo, hi):
        return (rng.random(n, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).reshape(shape)

    if key.endswith("num_batches_tracked"):
        return np.zeros((), dtype=np.int64)
    if key.startswith("char_images"):
        return uniform(0.0, 1.0)
    if key.startswith("pho_gru."):
        k = 1.0 / math.sqrt(shape[-1] if len(shape) == 2 else 768)
        return uniform(-k, k)
    if key.startswith("resnet.res_block"):
        if len(shape) == 4:
            b = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            return uniform(-b, b)
        if key.endswith("running_var") or key.endswith(".weight"):
            return uniform(0.5, 1.5)
        return normal(0.1)  # BN bias, running_mean
    if key.endswith("LayerNorm.weight") or key == "resnet_layernorm.weight":
        return (1.0 + normal(0.1)).astype(np.float32)
    if key.endswith("LayerNorm.bias") or key == "resnet_layernorm.bias":
        return normal(0.05)
    if key.endswith(".bias"):
        return normal(0.02)
    return normal(0.02)


def synth_state_dict(cfg: ArchConfig, seed: int = 0, tie_cls: bool = True):
    """fp32 CPU tensors keyed like the reference state_dict.  With tie_cls (the shipped setup,
    src/models.py:700-701) classifier.weight aliases bert.embeddings.word_embeddings.weight."""
    sd = {}
    for key, shape in state_dict_spec(cfg):
        if tie_cls and key == "classifier.weight":
            sd[key] = sd["bert.embeddings.word_embeddings.weight"]
            continue
        sd[key] = torch.from_numpy(np.ascontiguousarray(_gen(key, shape, seed)))
    return sd


def synth_batch(B: int, L: int, seed: int = 1234, ragged: bool = True, with_labels: bool = True,
                vocab_size: int = 21128):
    """SURVEY.md §8d synthetic batch: [CLS] ids [SEP] pad..., SIGHAN-like 5 % edits in tgt_idx,
    pinyin = tone digit then 1..6 letters (src/utils.py:72-98), 'U' (=32) for special tokens."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if ragged:
        lens = rng.integers(max(1, L // 2), L - 1, size=B)  # in [L/2, L-2]
    else:
        lens = np.full(B, L - 2)
    src = np.zeros((B, L), dtype=np.int64)
    masks = np.zeros((B, L), dtype=np.int64)
    loss_masks = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        n = int(lens[b])
        src[b, 0] = CLS
        src[b, 1:n + 1] = rng.integers(1, vocab_size, size=n)
        src[b, n + 1] = SEP
        masks[b, :n + 2] = 1
        loss_masks[b, 1:n + 1] = 1
    tgt = src.copy()
    edit = (rng.random((B, L)) < 0.05) & (loss_masks == 1)
    tgt[edit] = rng.integers(1, vocab_size, size=int(edit.sum()))
    flat = src.reshape(-1)
    special = (flat == PAD) | (flat == CLS) | (flat == SEP)
    pho_lens = np.where(special, 1, rng.integers(2, 8, size=flat.shape[0]))
    T = int(pho_lens.max())
    pho = np.zeros((flat.shape[0], T), dtype=np.int64)
    tone = rng.integers(1, 6, size=flat.shape[0])
    letters = rng.integers(6, 32, size=(flat.shape[0], T))
    for t in range(T):
        col = np.where(t == 0, tone, letters[:, t])
        pho[:, t] = np.where(t < pho_lens, col, 0)
    pho[special, 0] = 32
    batch = {
        "src_idx": torch.from_numpy(src), "masks": torch.from_numpy(masks),
        "loss_masks": torch.from_numpy(loss_masks),
        "pho_idx": torch.from_numpy(pho), "pho_lens": [int(x) for x in pho_lens],
    }
    if with_labels:
        batch["tgt_idx"] = torch.from_numpy(tgt)
    return batch
