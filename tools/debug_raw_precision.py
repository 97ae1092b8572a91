"""Experiment: unshared-gate gradient / logit error of the train step with bf16 vs f32 raw conv outputs in res_block1-2,
next to torch autocast(bf16) of the oracle (an independent 16-bit implementation)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_bench_shape_parity_gpu as T
from oracle import realise_oracle as O
from realise_b200 import ops
from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
from realise_b200.train import TrainEngine

torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
B, L, seed = int(sys.argv[1]), 128, 4242
cfg = ArchConfig()
sd = synth_state_dict(cfg, 0)
batch = synth_batch(B, L, seed=99, ragged=True)
ob = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
mask_fn = lambda site, shape: ops.dropout_mask(int(np.prod(shape)), 0.1, seed, site).reshape(shape).float()
rsd, leaves = T.oracle_leaves(sd, "cuda")
O.MASK_FN, O.FAST = mask_fn, True
rloss, rlogits = O.forward(rsd, ob, cfg, train=True); rloss.backward()
rsd2, leaves2 = T.oracle_leaves(sd, "cuda")
with torch.autocast("cuda", dtype=torch.bfloat16):
    l2, lg2 = O.forward(rsd2, ob, cfg, train=True)
l2.backward()
ac = {k: float((leaves2[k].grad.float() - v.grad).norm() / v.grad.norm()) for k, v in leaves.items()
      if v.grad is not None and leaves2[k].grad is not None and float(v.grad.norm()) > 1e-12 and not k.endswith("key.bias")}
cnn = [v for k, v in ac.items() if k.startswith("resnet.")]; rest = [v for k, v in ac.items() if not k.startswith("resnet.")]
print("autocast-bf16 oracle vs fp32 oracle: logits max", float((lg2.float() - rlogits).abs().max()), "rms", float((lg2.float() - rlogits).pow(2).mean().sqrt()),
      "| cnn grads max/median", max(cnn), float(np.median(cnn)), "| rest max/median", max(rest), float(np.median(rest)), flush=True)
del rsd2, leaves2, l2, lg2
for raw_bf16 in (True, False):
    m = T.build_model(cfg, 0, train=True)
    eng = m._engine = TrainEngine(m); eng.set_seed(seed); eng.raw_bf16 = raw_bf16
    loss, logits = m(T.to_dev(batch)); loss.backward(); torch.cuda.synchronize()
    errs = T.grad_errors(m, leaves)
    print("ours raw_bf16 =", raw_bf16, ": logits max", float((logits.float() - rlogits).abs().max()), "rms", float((logits.float() - rlogits).pow(2).mean().sqrt()),
          T.summarize(errs), flush=True)
    del m, eng
O.MASK_FN, O.FAST = None, False
