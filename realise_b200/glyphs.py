"""Glyph table builder: the host-side, one-off rasterisation behind
`model.build_glyce_embed(vocab_dir, font_path)` / `model.build_glyce_embed_multifonts(vocab_dir, num_fonts,
use_traditional_font)` (reference: src/models.py:703-795, called from src/run.py:433-440).

Behaviour restated from the reference, quirks included:
  * every vocabulary entry is drawn with `ImageFont.truetype(font, 32).getmask(char)`; the mask is cropped to its
    top-left 32x32 and then ALWAYS pasted centred into a zero 32x32 canvas — the reference's test
    `image.size != (font_size, font_size)` compares numpy's element COUNT with a tuple, so it is true for every glyph
    (src/models.py:722, :784);
  * single-font builder (`build_glyce_embed`): entries that are not exactly one CJK ideograph (`_is_chinese_char`,
    src/models.py:20-30) stay all-zero; the multi-font builder only skips multi-character entries ("[CLS]", "##ing"),
    so letters, digits and punctuation ARE drawn there (src/models.py:773-775);
  * each font's table is standardised with the mean / std over the WHOLE table (zeros included) (:731, :793);
  * multi-font order: simhei, xiaozhuan, simhei on the traditional form (opencc s2t) — `font_paths[:num_fonts]`, the
    last one replaced by traditional simhei when `use_traditional_font` (:742-750).

Pure numpy + PIL; nothing here runs per batch.  The result is written into the model's frozen `char_images.weight`
([vocab, 1024]) or `char_images_multifonts` ([vocab, C, 32, 32]) parameter, which the CUDA stem kernels gather from.
"""
import os

import numpy as np
import torch


def is_chinese_char(cp):
    """src/models.py:20-30 (the CJK Unified Ideographs blocks BERT's tokenizer treats as Chinese characters)."""
    return ((0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x20000 <= cp <= 0x2A6DF) or (0x2A700 <= cp <= 0x2B73F)
            or (0x2B740 <= cp <= 0x2B81F) or (0x2B820 <= cp <= 0x2CEAF) or (0xF900 <= cp <= 0xFAFF)
            or (0x2F800 <= cp <= 0x2FA1F))


def read_vocab(vocab_dir):
    with open(os.path.join(vocab_dir, "vocab.txt"), "r", encoding="utf-8") as f:
        return [s.strip() for s in f]


def _draw(font, char, size):
    mask = font.getmask(char)
    img = np.asarray(mask).astype(np.float32).reshape(mask.size[::-1])     # PIL size is (w, h): rows = size[1]
    img = img[:size, :size]
    canvas = np.zeros((size, size), dtype=np.float32)
    o0, o1 = (size - img.shape[0]) // 2, (size - img.shape[1]) // 2
    canvas[o0:o0 + img.shape[0], o1:o1 + img.shape[1]] = img
    return canvas


def rasterize(vocab, font_path, font_size=32, chinese_only=False, convert=None):
    """[len(vocab), size, size] float32, standardised over the whole table.  chinese_only: the single-font rule (only
    one-ideograph entries are drawn); convert: optional char -> char map applied to one-character entries (opencc s2t)."""
    from PIL import ImageFont
    font = ImageFont.truetype(font_path, size=font_size)
    out = np.zeros((len(vocab), font_size, font_size), dtype=np.float32)
    for i, char in enumerate(vocab):
        if convert is not None and len(char) == 1:
            char = convert(char)
        if chinese_only:
            if len(char) != 1 or not is_chinese_char(ord(char)):
                continue
        elif len(char) > 1:
            continue
        out[i] = _draw(font, char, font_size)
    return (out - np.mean(out)) / np.std(out)


def build_glyce_embed(model, vocab_dir, font_path, font_size=32):
    """src/models.py:703-733 — fills model.char_images.weight ([vocab, font_size^2], frozen)."""
    table = rasterize(read_vocab(vocab_dir), font_path, font_size, chinese_only=True)
    flat = torch.from_numpy(table).reshape(table.shape[0], -1)
    if tuple(flat.shape) != tuple(model.char_images.weight.shape):
        raise ValueError(f"vocab.txt gives a {tuple(flat.shape)} glyph table, the model holds {tuple(model.char_images.weight.shape)}")
    model.char_images.weight.data.copy_(flat)
    model._invalidate()


def multifont_plan(num_fonts, use_traditional_font):
    """[(font file, draw the traditional form?)] — src/models.py:737-747."""
    fonts = [("simhei.ttf", False), ("xiaozhuan.ttf", False), ("simhei.ttf", True)][:num_fonts]
    if use_traditional_font:
        fonts = fonts[:-1] + [("simhei.ttf", True)]
    return fonts


def build_glyce_embed_multifonts(model, vocab_dir, num_fonts, use_traditional_font, font_size=32, font_dir=".",
                                 s2t=None):
    """src/models.py:735-761 — fills model.char_images_multifonts ([vocab, num_fonts, 32, 32], frozen).  Fonts are looked
    up in font_dir (the reference resolves 'simhei.ttf' / 'xiaozhuan.ttf' against the working directory).  s2t: the
    simplified -> traditional converter, default opencc.OpenCC('s2t.json').convert (imported only when needed)."""
    vocab = read_vocab(vocab_dir)
    planes = []
    for font, traditional in multifont_plan(num_fonts, use_traditional_font):
        conv = None
        if traditional:
            if s2t is None:
                import opencc
                s2t = opencc.OpenCC("s2t.json").convert
            conv = s2t
        planes.append(torch.from_numpy(rasterize(vocab, os.path.join(font_dir, font), font_size, convert=conv)))
    table = torch.stack(planes, dim=1).contiguous()
    if tuple(table.shape) != tuple(model.char_images_multifonts.shape):
        raise ValueError(f"glyph table {tuple(table.shape)} does not match the model's {tuple(model.char_images_multifonts.shape)}")
    model.char_images_multifonts.data.copy_(table)
    model._invalidate()
