"""Attention kernel alone at the bench shape (for timing + ncu)."""
import sys

import torch

sys.path.insert(0, ".")
from realise_b200 import ops  # noqa: E402

B, L, H = 64, 128, 768
qkv = (torch.randn(B * L, 3 * H, device="cuda") * 0.5).bfloat16()
mask = torch.ones(B, L, dtype=torch.int64, device="cuda")
ctx = torch.empty(B * L, H, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, mask, ctx, B, L, 12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(qkv, mask, ctx, B, L, 12)
e1.record()
torch.cuda.synchronize()
print(f"attention B{B} L{L}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
# correctness vs torch
q, k, v = qkv.float().view(B, L, 3, 12, 64).permute(2, 0, 3, 1, 4)
ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
ref = ref.permute(0, 2, 1, 3).reshape(B * L, H)
print("max err", (ctx.float() - ref).abs().max().item())
torch.cuda.cudart().cudaProfilerStart()
ops.attention(qkv, mask, ctx, B, L, 12)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
