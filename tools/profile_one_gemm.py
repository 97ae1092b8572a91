"""One GEMM launch bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full --import-source on --warp-sampling-interval 0`.
usage: profile_one_gemm.py M N K [f32]"""
import sys
import torch
sys.path.insert(0, ".")
from realise_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16384, 2304, 768)
odt = torch.float32 if "f32" in sys.argv else torch.bfloat16
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16(); bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=odt)
for _ in range(3):
    ops.gemm(a, b, out, bias=bias)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.gemm(a, b, out, bias=bias)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
