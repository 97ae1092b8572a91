"""Bottleneck decomposition of the GEMM kernels with the debug knobs (results are garbage, only time matters)."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402
from realise_b200._lib import lib  # noqa: E402

dev = "cuda"
M = 8192
a = torch.randn(M, 768, device=dev).bfloat16()
w1 = torch.randn(3072, 768, device=dev).bfloat16()
b1 = torch.randn(3072, device=dev)
o1 = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
names = {0: "full", 1: "no-epilogue", 2: "no-TMA", 4: "no-MMA", 3: "MMA only", 5: "TMA only", 6: "epilogue only"}
for pair in (0, 1):
    lib().rl_gemm_set_pair_mode(pair)
    for bn in (256, 128):
        lib().rl_gemm_set_tile_n(bn)
        for dbg in (0, 1, 2, 4, 3, 5, 6):
            lib().rl_gemm_set_debug_mode(dbg)
            for _ in range(2):
                ops.gemm(a, w1, o1, bias=b1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.gemm(a, w1, o1, bias=b1)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 10
            print(f"pair={pair} bn={bn} {names[dbg]:14s}: {t*1e3:6.1f} us", flush=True)
lib().rl_gemm_set_debug_mode(0)
