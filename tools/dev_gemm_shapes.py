"""Per-shape GEMM time inside one eager train step (B=128, L=128): which shapes carry the GEMM milliseconds."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402
from realise_b200.optim import FusedAdamW  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch  # noqa: E402

B, L = 128, 128
cfg = ArchConfig()
torch.manual_seed(0)
model = SpellBertPho2ResArch3Abla(cfg)
model.tie_cls_weight()
model.train().cuda()
batch = synth_batch(B, L, seed=1, ragged=False)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=5e-5, max_grad_norm=1.0, model=model)


def step():
    loss = model(db)[0]
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
ops._prof = []
step()
torch.cuda.synchronize()
prof, ops._prof = ops._prof, None
agg = {}
for kind, work, e0, e1, detail in prof:
    if kind not in ("gemm", "conv_gemm"):
        continue
    a = agg.setdefault(detail, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
    a[2] += work
tot = sum(a[1] for a in agg.values())
print(f"total GEMM time {tot / 1e3:.2f} ms, {sum(a[0] for a in agg.values())} launches")
print("   us_total   n   us_each  TFLOP/s  (M, N, K, a_t|conv, b_t|taps, split_k|remap, act, out, res, drop)")
for d, (n, us, w) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} {n:3d} {us / n:9.1f} {w / us / 1e6:8.1f}  {d}")
