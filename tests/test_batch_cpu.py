"""CPU: the tabulated pinyin lookup reproduces the reference's per-batch conversion (src/utils.py:72-98)."""
import torch

from realise_b200.batch import MAX_PHO_LEN, PinyinTable, pho_vocab

FAKE = {"中": "1zhong", "国": "2guo", "装": "1zhuang", "的": "5de", "a": "U", "[CLS]": "U", "[PAD]": "U", "##ing": "U"}


def get_pinyin(tok):                       # stand-in for Pinyin2.get_pinyin (pypinyin is not in this image)
    return "U" if len(tok) > 1 else FAKE.get(tok, "U")


def reference_convert(chars):
    """src/utils.py:86-98 restated: ids per symbol, pad_sequence(padding_value=0), lengths."""
    vocab = pho_vocab()
    pinyins = [get_pinyin(c) for c in chars]
    ids = [torch.tensor([vocab[ch] for ch in p]) for p in pinyins]
    return torch.nn.utils.rnn.pad_sequence(ids, batch_first=True, padding_value=0), [len(p) for p in pinyins]


def test_table_lookup_matches_per_batch_conversion():
    tokens = list(FAKE.keys())
    tab = PinyinTable.from_tokens(tokens, get_pinyin)
    assert tab.table.shape == (len(tokens), MAX_PHO_LEN)
    g = torch.Generator().manual_seed(0)
    src = torch.randint(0, len(tokens), (3, 9), generator=g)
    pho_idx, pho_lens = tab.lookup(src)
    ref_idx, ref_lens = reference_convert([tokens[i] for i in src.flatten().tolist()])
    T = ref_idx.shape[1]
    assert torch.equal(pho_idx[:, :T], ref_idx) and int(pho_idx[:, T:].abs().sum()) == 0
    assert pho_lens.tolist() == ref_lens
    batch = tab.build_batch({"src_idx": src})
    assert batch["pho_idx"].shape == (27, MAX_PHO_LEN) and batch["pho_lens"].dtype == torch.int32


def test_pho_vocab_is_the_reference_symbol_table():
    v = pho_vocab()
    assert len(v) == 33 and v["P"] == 0 and v["1"] == 1 and v["5"] == 5 and v["a"] == 6 and v["z"] == 31 and v["U"] == 32


def test_pho_lens_host_path_accepts_lists_tensors_and_empty_input():
    """realise_b200.hostio.lens_to_device: batch['pho_lens'] arrives as a Python list (src/utils.py:92-98); the result is
    an int32 tensor with the same values whichever container it came in (the CUDA branch adds a pinned staging ring)."""
    import torch
    from realise_b200 import hostio
    lens = [1, 7, 3, 2, 6, 1]
    a = hostio.lens_to_device(lens, "cpu")
    assert a.dtype == torch.int32 and a.tolist() == lens
    b = hostio.lens_to_device(torch.tensor(lens, dtype=torch.int64), "cpu")
    assert b.dtype == torch.int32 and b.tolist() == lens
    assert hostio.lens_to_device([], "cpu").numel() == 0
    # the returned tensor owns its memory (the list's temporary buffer is gone)
    lens2 = list(range(1, 2001))
    c = hostio.lens_to_device(lens2, "cpu")
    del lens2
    assert int(c.sum()) == 2001 * 2000 // 2
