"""One training step (fwd + bwd + clip + AdamW) of the full model at the bench shape between
cudaProfilerStart/Stop (for ncu --profile-from-start off).  Also prints host-issue time vs device time per step."""
import sys
import time

import torch

sys.path.insert(0, ".")
from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402
from realise_b200.optim import FusedAdamW  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch  # noqa: E402

B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 128)
cfg = ArchConfig(with_pho="yes", with_res="yes")
torch.manual_seed(0)
model = SpellBertPho2ResArch3Abla(cfg)
model.tie_cls_weight()
model.train().cuda()
batch = synth_batch(B, L, seed=1, ragged=False)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
db["pho_lens"] = torch.tensor(batch["pho_lens"], dtype=torch.int32, device="cuda")
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=5e-5, max_grad_norm=1.0, model=model)


def step():
    loss = model(db)[0]
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print(f"host issue {t_issue * 1e3:.2f} ms, issue+drain {t_total * 1e3:.2f} ms", flush=True)
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
