"""GPU dev check: CUDA model vs committed goldens (reference outputs) and the CPU oracle."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from realise_b200.model import SpellBertPho2ResArch3  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402


def unsplit(x, n, S, C):
    """parity-split rows [n][ph][pw][S/2][S/2][C] -> NCHW [n, C, S, S]"""
    h = S // 2
    t = x.view(n, 2, 2, h, h, C).permute(0, 5, 3, 1, 4, 2).reshape(n, C, S, S)
    return t


def main():
    g = np.load("tests/golden/arch3_eval_B2_L16.npz")
    cfg = ArchConfig()
    t0 = time.time()
    sd = synth_state_dict(cfg, 0)
    model = SpellBertPho2ResArch3(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.eval().cuda()
    print("model ready", time.time() - t0, flush=True)
    batch = synth_batch(2, 16, seed=1234)
    dbatch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    model.collect = {}
    with torch.no_grad():
        loss, logits = model(dbatch)
    torch.cuda.synchronize()
    c = model.collect
    print("loss", loss.item(), float(g["loss"]), flush=True)
    for k in ["bert_hiddens", "pho_gru", "pho_hiddens", "resnet", "res_hiddens"]:
        a = c[k].cpu().numpy().reshape(g[k].shape)
        print(f"{k}: max_abs_err {np.abs(a - g[k]).max():.4e} (ref absmax {np.abs(g[k]).max():.3f})", flush=True)
    n = 32
    b1 = unsplit(c["res_block1_split"], n, 16, 64).cpu().numpy()[:8]
    print(f"res_block1: {np.abs(b1 - g['res_block1']).max():.4e} (ref absmax {np.abs(g['res_block1']).max():.3f})")
    b2 = unsplit(c["res_block2_split"], n, 8, 128).cpu().numpy()[:8]
    print(f"res_block2: {np.abs(b2 - g['res_block2']).max():.4e} (ref absmax {np.abs(g['res_block2']).max():.3f})")
    a = c["sequence_output"].cpu().numpy().reshape(g["output_block"].shape)
    print(f"output_block: {np.abs(a - g['output_block']).max():.4e}")
    flat = logits.reshape(32, -1).float().cpu()
    err = np.abs(flat[torch.from_numpy(g["logits_rows"])].numpy() - g["logits_kept"]).max()
    am = flat.argmax(-1).numpy()
    gap = g["logits_top2_gap"]
    mism = am != g["logits_argmax"]
    print(f"logits max_abs_err {err:.4e}; argmax mismatches {mism.sum()} / 32; gaps at mismatches {gap[mism]}")
    # timing at the bench shape
    B, L = 64, 128
    batch = synth_batch(B, L, seed=1, ragged=False, with_labels=False)
    dbatch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    model.collect = None
    with torch.no_grad():
        for _ in range(3):
            model(dbatch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(dbatch)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"forward B{B} L{L}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} sentences/s", flush=True)


if __name__ == "__main__":
    main()
