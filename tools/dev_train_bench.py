"""Timing of one training step (fwd + bwd + clip + AdamW) of the semantic-only configuration on 1 GPU."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402
from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402
from realise_b200.optim import FusedAdamW  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch  # noqa: E402

B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 128)
FULL = "--full" in sys.argv
cfg = ArchConfig(with_pho="yes" if FULL else "no", with_res="yes" if FULL else "no")
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
    return loss


for _ in range(3):
    l0 = step()
torch.cuda.synchronize()
print("loss", l0.item(), "mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    l1 = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"train step B{B} L{L} (sem-only, 12+3 layers, dropout 0.1): {ms:.2f} ms -> {B/ms*1e3:.0f} sentences/s; loss {l1.item():.4f}")
# time every ops.* wrapper (including the ones without a built-in _Timed bracket)
import types  # noqa: E402
names = [n for n, f in vars(ops).items() if isinstance(f, types.FunctionType) and not n.startswith("_")
         and n not in ("gemm", "conv_gemm", "attention", "layernorm", "embed_ln", "gate_fuse", "masked_ce", "gru_step",
                       "glyph_stem", "glyph_block1", "argmax_rows", "attention_bwd", "layernorm_bwd", "dropout_mask")]
import realise_b200.train as T  # noqa: E402


def wrap(name, fn):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        if ops._prof is not None:
            ops._prof.append((name, 0, e0, e1))
        return r
    return inner


for n in names:
    setattr(ops, n, wrap(n, getattr(ops, n)))
ops._prof = []
step()
torch.cuda.synchronize()
agg = {}
for kind, work, a, b in ops._prof:
    t = agg.setdefault(kind, [0, 0.0])
    t[0] += 1
    t[1] += a.elapsed_time(b)
ops._prof = None
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:16s} n={n:4d} {t:8.3f} ms")
