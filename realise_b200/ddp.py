"""Data parallelism for the training step: batch sharding + ONE gradient all-reduce over the flat buffer.

Reference behaviour (src/run.py:131-137, :165-167, :200): every rank keeps examples r, r+W, ... (tail dropped so all
ranks see the same count), the model is wrapped in DistributedDataParallel, and `loss.backward()` all-reduces
(sum / W) the gradients in 25 MB buckets.  Here the gradients of a rank already live in one contiguous fp32 buffer
(realise_b200.train.TrainEngine.flat), so the exchange is a single NCCL all-reduce(sum) on it — over NVLink 5 /
NVSwitch on the 8xB200 box — and the division by W is folded into the fused clip+AdamW kernel (grad_div).
BatchNorm statistics stay per-rank (plain BatchNorm2d in the reference, no SyncBN).
"""
import os

import torch
import torch.distributed as dist


def shard_examples(examples, rank, world_size):
    """src/run.py:131-137: rank-strided shard, tail dropped so every rank has len(examples) // world_size items."""
    n = len(examples) // world_size
    return examples[rank::world_size][:n]


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of a flat gradient buffer (NCCL for CUDA tensors, gloo for the CPU tests)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 1
    if flat.is_cuda and dist.get_backend(group) != "nccl":
        # gloo stages device tensors through host memory on its own streams: fence both sides so that the kernels that
        # produced `flat` are done before it is read and the reduced values are in place before the optimizer runs
        # (test-only path: two ranks sharing one GPU; NCCL is stream-ordered and needs none of this)
        torch.cuda.synchronize()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.synchronize()
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return dist.get_world_size(group)


class DataParallel:
    """Attach to a realise_b200 model: after every backward the flat gradient buffer is all-reduced and the
    optimizer divides by the world size.  Usage mirrors the reference loop:

        dp = DataParallel(model)                       # instead of DistributedDataParallel(model, ...)
        loss = model(batch)[0]; loss.backward()        # gradients are summed across ranks here
        optimizer.step()                               # FusedAdamW(model=model) reads model.grad_div
    """

    def __init__(self, model, group=None, average=False, overlap=None, sm_reserve=16):
        """average=False: gradients are SUMMED over ranks and FusedAdamW(model=model) divides by the world size inside
        its update kernel (model.grad_div).  average=True: the buffer is scaled by 1/W right after the all-reduce, which
        is what DistributedDataParallel leaves in p.grad — for optimizers that know nothing about grad_div (the
        reference's vendored AdamW behind realise_b200.compat)."""
        self.model, self.group, self.average = model, group, average
        world = float(dist.get_world_size(group)) if dist.is_initialized() else 1.0
        model.grad_div = 1.0 if average else world
        self._inv_world = 1.0 / world
        model._post_backward = self.sync
        # overlap (opt-in: overlap=True or RL_DP_OVERLAP=1): the flat buffer is laid out in backward-completion order
        # (TrainEngine.buckets); each bucket is all-reduced on a communication stream as soon as the backward has issued
        # its last gradient kernel, behind the remaining backward.  The persistent GEMMs normally take every SM, which
        # would leave NCCL's CTAs waiting (or a GEMM's last CTAs waiting for NCCL): while a reduction may be in flight
        # the GEMMs are launched on `sm_reserve` fewer SMs.
        # MEASURED on 8 x B200 (profiles/r02_bench_train_n8*.json): 45.98 ms/step with the overlap vs 45.63 ms with one
        # all-reduce after the backward, 43.96 vs 43.90 at N = 2 — the SMs lent to NCCL cost the GEMMs what the overlap
        # hides of a 1.3 ms exchange, so the single all-reduce stays the default.
        if overlap is None:
            overlap = world > 1 and dist.get_backend(group) == "nccl" and os.environ.get("RL_DP_OVERLAP", "0") == "1"
        self.overlap, self.sm_reserve, self._comm, self._pending = bool(overlap), int(sm_reserve), None, False
        if self.overlap:
            model._bucket_ready = self.bucket_ready

    def bucket_ready(self, engine, k):
        from . import ops
        a, b = engine.buckets[k]
        if b <= a:
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(engine.flat.device)
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)                     # the bucket's gradient kernels were issued on `main`
        with torch.cuda.stream(self._comm):
            dist.all_reduce(engine.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)
        self._pending = True
        ops.SM_RESERVE = self.sm_reserve                 # GEMMs of the remaining backward leave SMs to NCCL

    def sync(self, engine):
        if self.overlap and self._pending:
            from . import ops
            torch.cuda.current_stream().wait_stream(self._comm)     # join: every bucket has been reduced
            ops.SM_RESERVE = 0
            self._pending = False
            w = dist.get_world_size(self.group)
        else:
            w = allreduce_sum_(engine.flat, self.group)
        if self.average and w > 1:
            engine.flat.mul_(self._inv_world)

    def broadcast_parameters(self, src=0):
        """DDP broadcasts rank-0 parameters/buffers at construction; do the same once."""
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)
            self.model._prepared = None
