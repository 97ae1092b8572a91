"""Representative GEMM launches for an `ncu --set full` capture: pair (cta_group::2) and single-CTA kernels."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402
from realise_b200._lib import lib  # noqa: E402

dev = "cuda"
M = 8192
a = torch.randn(M, 768, device=dev).bfloat16()
a2 = torch.randn(M, 3072, device=dev).bfloat16()
w1 = torch.randn(3072, 768, device=dev).bfloat16()
w2 = torch.randn(768, 3072, device=dev).bfloat16()
b1 = torch.randn(3072, device=dev)
b2 = torch.randn(768, device=dev)
o1 = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
o2 = torch.empty(M, 768, device=dev, dtype=torch.float32)
r2 = torch.randn(M, 768, device=dev)


def run():
    ops.gemm(a, w1, o1, bias=b1)                 # big-N, bf16 out
    ops.gemm(a2, w2, o2, bias=b2, res=r2)        # deep-K, f32 out + residual


for mode in (1, 0):
    lib().rl_gemm_set_pair_mode(mode)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, w1, o1, bias=b1)
    e1.record()
    torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10):
        ops.gemm(a2, w2, o2, bias=b2, res=r2)
    e1.record()
    torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / 10
    print(f"pair={mode}: FFN1-like {t1*1e3:.1f} us ({2*M*3072*768/t1/1e9:.0f} TF/s), FFN2-like {t2*1e3:.1f} us "
          f"({2*M*3072*768/t2/1e9:.0f} TF/s)", flush=True)
    torch.cuda.cudart().cudaProfilerStart()
    run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
