"""Two representative GEMM launches for an `ncu --set full` capture (bf16-out and f32-out+residual)."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402

dev = "cuda"
M = 8192
a = torch.randn(M, 768, device=dev).bfloat16()
w1 = torch.randn(2304, 768, device=dev).bfloat16()
w2 = torch.randn(768, 768, device=dev).bfloat16()
b1 = torch.randn(2304, device=dev)
b2 = torch.randn(768, device=dev)
o1 = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
o2 = torch.empty(M, 768, device=dev, dtype=torch.float32)
r2 = torch.randn(M, 768, device=dev)
for _ in range(2):
    ops.gemm(a, w1, o1, bias=b1)
    ops.gemm(a, w2, o2, bias=b2, res=r2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.gemm(a, w1, o1, bias=b1)
ops.gemm(a, w2, o2, bias=b2, res=r2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
