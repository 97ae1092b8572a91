"""Device-side batch builder: the pinyin part of `build_batch` as a table lookup.

The reference converts every token of every batch on the host (src/models.py:797-804 -> src/utils.py:72-98:
`Pinyin2.convert` calls pypinyin per character, then `pad_sequence`) and ships `pho_idx` to the GPU.  The pinyin
string is a pure function of the vocabulary entry, so it is tabulated once — `table[v] = ids of the TONE3 string with
the tone digit first` (0-padded to 7 symbols), `lens[v]` — and a batch needs only `table[src_idx]` on the device: no
per-batch Python loop, no H2D copy of `pho_idx`, and a fixed T = 7 (one CUDA-graph shape for every batch).
Symbols beyond a token's length are never read by the GRU (length masking), so padding to 7 instead of to the batch
maximum changes nothing downstream.
"""
import torch

MAX_PHO_LEN = 7   # tone digit + at most 6 letters ('zhuang1' -> '1zhuang')


def pho_vocab():
    """Symbol ids of src/utils.py:61-67: 'P' (pad) 0, '1'..'5' 1-5, 'a'..'z' 6-31, 'U' (no pinyin) 32."""
    syms = ["P"] + [chr(x) for x in range(ord("1"), ord("5") + 1)] + [chr(x) for x in range(ord("a"), ord("z") + 1)] + ["U"]
    return {c: i for i, c in enumerate(syms)}


class PinyinTable:
    def __init__(self, table, lens):
        assert table.dtype == torch.int64 and lens.dtype == torch.int32 and table.shape == (lens.numel(), MAX_PHO_LEN)
        self.table, self.lens = table.contiguous(), lens.contiguous()

    @classmethod
    def from_tokens(cls, tokens, get_pinyin):
        """tokens[v] = vocabulary string of id v (tokenizer.convert_ids_to_tokens(range(V))); get_pinyin(token) -> the
        reference's `Pinyin2.get_pinyin` string ('1zhuang', 'U', ...)."""
        vocab = pho_vocab()
        V = len(tokens)
        table = torch.zeros(V, MAX_PHO_LEN, dtype=torch.int64)
        lens = torch.zeros(V, dtype=torch.int32)
        for v, tok in enumerate(tokens):
            s = get_pinyin(tok)
            if not 1 <= len(s) <= MAX_PHO_LEN:
                raise ValueError(f"pinyin {s!r} of token {tok!r} (id {v}) has {len(s)} symbols, expected 1..{MAX_PHO_LEN}")
            ids = [vocab[ch] for ch in s]
            table[v, :len(ids)] = torch.tensor(ids, dtype=torch.int64)
            lens[v] = len(ids)
        return cls(table, lens)

    def to(self, device):
        return PinyinTable(self.table.to(device), self.lens.to(device))

    def lookup(self, src_idx):
        """src_idx int64 [B, L] (any device matching the table) -> pho_idx int64 [B*L, 7], pho_lens int32 [B*L]."""
        flat = src_idx.reshape(-1)
        return self.table.index_select(0, flat), self.lens.index_select(0, flat)

    def build_batch(self, batch):
        """Drop-in for `model_class.build_batch(batch, tokenizer)` once the batch ids are on the device."""
        batch["pho_idx"], batch["pho_lens"] = self.lookup(batch["src_idx"])
        return batch
