"""GPU dev check: fused res_block1 kernel vs the CPU oracle (and vs the unfused stem + conv path)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import realise_oracle as O  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402


def unsplit(x, n, S, C):
    h = S // 2
    return x.view(n, 2, 2, h, h, C).permute(0, 5, 3, 1, 4, 2).reshape(n, C, S, S)


for fonts in (3, 1):
    cfg = ArchConfig(num_hidden_layers=1, num_fonts=fonts)
    sd = synth_state_dict(cfg, 7)
    model = SpellBertPho2ResArch3(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.eval().cuda()
    model.use_cuda_graph = False
    P = model.prepare()
    n = 300
    ids = torch.randint(0, 21128, (n,), device="cuda")
    col = {}
    with torch.no_grad():
        ref = O.char_resnet(sd, O.glyph_images(sd, cfg, ids.cpu()), collect=col)
    for fuse in (False, True):
        model.fuse_block1 = fuse
        model.collect = {}
        with torch.no_grad():
            out = model._resnet(P, ids, n)
        torch.cuda.synchronize()
        b1 = unsplit(model.collect["res_block1_split"], n, 16, 64).cpu()
        print(f"fonts={fonts} fuse={fuse}: block1 err {(b1 - col['res_block1']).abs().max().item():.4e} "
              f"(ref max {col['res_block1'].abs().max().item():.3f}); resnet err {(out.cpu() - ref).abs().max().item():.4e}",
              flush=True)
    model.collect = None
    if fonts == 3:
        n = 8192
        ids = torch.randint(0, 21128, (n,), device="cuda")
        for fuse in (False, True):
            model.fuse_block1 = fuse
            with torch.no_grad():
                for _ in range(2):
                    model._resnet(P, ids, n)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    model._resnet(P, ids, n)
                e1.record()
                torch.cuda.synchronize()
            print(f"resnet n={n} fuse={fuse}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
