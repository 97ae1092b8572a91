"""One eager forward at the bench shape between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200.model import SpellBertPho2ResArch3  # noqa: E402
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict  # noqa: E402

cfg = ArchConfig()
model = SpellBertPho2ResArch3(cfg)
model.tie_cls_weight()
model.load_state_dict(synth_state_dict(cfg, 0), strict=True)
model.eval().cuda()
model.use_cuda_graph = False
B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 128)
batch = synth_batch(B, L, seed=1234, ragged=False, with_labels=False)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
with torch.no_grad():
    model(db)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model(db)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
