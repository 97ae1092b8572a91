"""max |logit - oracle logit| of the eval forward on the FULL architecture (12+4+3 layers), with and without the
split-precision classifier.  Oracle = CPU fp32 restatement on the same seeded weights / batch."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import realise_oracle as O  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402

cfg = ArchConfig()
sd = synth_state_dict(cfg, seed=5)
O.FAST = True
for B, L, seed in ((4, 128, 3), (8, 64, 4)):
    batch = synth_batch(B, L, seed=seed)
    with torch.no_grad():
        rloss, rlogits = O.forward(sd, batch, cfg)
    for precise in (False, True):
        m = SpellBertPho2ResArch3(cfg)
        m.tie_cls_weight()
        m.load_state_dict(sd, strict=True)
        m.precise_classifier = precise
        m.eval().cuda()
        m.collect = {}
        with torch.no_grad():
            loss, logits = m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()})
        d = (logits.float().cpu() - rlogits).abs()
        act = batch["masks"].bool()
        amax_ok = (logits.float().cpu().argmax(-1) == rlogits.argmax(-1))[act].float().mean().item()
        print(f"B{B} L{L} precise={precise}: max|dlogit| {d.max().item():.3e} (active tokens {d[act].max().item():.3e}), "
              f"rms {d.pow(2).mean().sqrt().item():.3e}, |logit|max {rlogits.abs().max().item():.2f}, loss {loss.item():.5f} vs "
              f"{rloss.item():.5f}, argmax agreement {amax_ok:.4f}", flush=True)
        del m
        torch.cuda.empty_cache()
