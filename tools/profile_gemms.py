"""Plain transformer GEMMs at the train-bench shapes, ONE launch each between cudaProfilerStart/Stop
(for `ncu --profile-from-start off --set full`): where does the persistent tcgen05 GEMM wait?"""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402

dev = "cuda"
N, H, I = 16384, 768, 3072


def f32(*s):
    return torch.randn(*s, device=dev)


def bf(*s):
    return torch.randn(*s, device=dev).bfloat16()


xb, w_qkv, w_1, w_2 = bf(N, H), bf(3 * H, H), bf(I, H), bf(H, I)
b3, bI, b1 = f32(3 * H), f32(I), f32(H)
qkvo, h, y32, res = bf(N, 3 * H), bf(N, I), f32(N, H), f32(N, H)
gw1 = f32(I, H)


def run():
    ops.gemm(xb, w_qkv, qkvo, bias=b3)                              # QKV: bf16 out
    ops.gemm(xb, w_1, h, bias=bI)                                   # FFN1 plain: bf16 out, K = 768
    ops.gemm(h, w_2, y32, bias=b1, res=res)                         # FFN2: f32 out + residual, K = 3072
    ops.gemm(h, w_1, y32, b_t=True, res=res)                        # dx1 = du W1 + dy2 (B MN-major)
    ops.gemm(h, xb, gw1, a_t=True, b_t=True, split_k=-1)            # dW1 = du^T x1 (both MN-major, split-K)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
