"""Probe: does tcgen05.mma kind::f16 accept different 16-bit formats for A and B?  One process per combination
(a faulting kernel poisons the CUDA context)."""
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from realise_b200 import ops
adt, bdt, mode = sys.argv[1], sys.argv[2], sys.argv[3]
D = {"bf16": torch.bfloat16, "f16": torch.float16}
g = torch.Generator(device="cuda").manual_seed(9)
M, N, K = 512, 384, 256
a = torch.randn(M, K, device="cuda", generator=g).to(D[adt]); b = torch.randn(N, K, device="cuda", generator=g).to(D[bdt])
ref = a.float() @ b.float().t()
out = torch.zeros(M, N, device="cuda")
if mode == "k":
    ops.gemm(a, b, out)
elif mode == "nopair":
    ops.TUNE_NO_PAIR = 1
    ops.gemm(a, b, out)
else:
    ops.gemm(a.t().contiguous(), b.t().contiguous(), out, a_t=True, b_t=True, split_k=-1)
torch.cuda.synchronize()
print("OK", adt, bdt, mode, float((out - ref).abs().max()))
'''
for adt in ("bf16", "f16"):
    for bdt in ("bf16", "f16"):
        for mode in ("k", "nopair", "mn"):
            p = subprocess.run([sys.executable, "-c", CODE, adt, bdt, mode], capture_output=True, text=True)
            print(adt, bdt, mode, "->", (p.stdout.strip() or p.stderr.strip().splitlines()[-1])[:200], flush=True)
