"""One attention forward + backward launch (B=128, L=128, 12 heads, dropout 0.1, saved lse) bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full --import-source on --warp-sampling-interval 0`."""
import sys
import torch
sys.path.insert(0, ".")
from realise_b200 import ops
B, L, heads = 128, 128, 12
H = heads * 64
torch.manual_seed(0)
qkv = (torch.randn(B * L, 3 * H, device="cuda") * 0.5).bfloat16()
lens = torch.randint(40, L + 1, (B,), device="cuda")
mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long().contiguous()
ctx = torch.empty(B * L, H, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * heads * L, device="cuda")
dctx = (torch.randn(B * L, H, device="cuda") * 0.1).bfloat16()
dqkv = torch.zeros_like(qkv)
drop = (0.1, 1234, 5)
for _ in range(3):
    ops.attention(qkv, mask, ctx, B, L, heads, drop=drop, lse=lse)
    ops.attention_bwd(qkv, mask, ctx, dctx, dqkv, B, L, heads, drop=drop, lse=lse)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.attention(qkv, mask, ctx, B, L, heads, drop=drop, lse=lse)
ops.attention_bwd(qkv, mask, ctx, dctx, dqkv, B, L, heads, drop=drop, lse=lse)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
