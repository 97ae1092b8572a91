"""One optimizer step as ONE CUDA-graph replay.

The reference's loop body (src/run.py:186-212) is

    loss = model(batch)[0]; loss.backward(); clip_grad_norm_(...); optimizer.step(); scheduler.step(); model.zero_grad()

i.e. ~700 kernel launches per step on this implementation; issued one by one from Python they cost about as much
host time as the GPU needs to run them.  `GraphedTrainStep` captures forward + backward + (data-parallel gradient
all-reduce) + clip + AdamW of one batch shape once and replays it:

    step = GraphedTrainStep(model, optimizer)          # optimizer = realise_b200.optim.FusedAdamW(model=model)
    for batch in loader:
        loss = step(batch)                             # device scalar (static buffer); .item() when needed
        scheduler.step()                               # LambdaLR keeps working: lr is re-read before every replay

What makes the capture replayable:
  * dropout masks are counter based; the kernels add a device-resident step counter to their seed
    (the kernels' drop_counter argument), and the graph itself increments that counter, so every replay draws fresh masks;
  * lr and the Adam bias corrections are read from a 3-float device buffer written before each replay
    (rl_mt_adamw_dev);
  * the batch is copied into static input buffers; shapes (B, L, T) key the graph cache, the first step of a new
    shape runs eagerly (it is a real training step) and doubles as the warm-up the capture needs.
"""
import torch

from . import hostio, ops
from .batch import MAX_PHO_LEN


class GraphedTrainStep:
    def __init__(self, model, optimizer, max_graphs=2):
        if not hasattr(optimizer, "hyper_values"):
            raise TypeError("GraphedTrainStep needs realise_b200.optim.FusedAdamW (the update must be capturable)")
        self.model, self.opt = model, optimizer
        self.max_graphs = max_graphs
        self._graphs = {}
        self._seen = set()
        dev = next(model.parameters()).device
        self._hyper = torch.zeros(3, device=dev, dtype=torch.float32)
        self._counter = torch.zeros(1, device=dev, dtype=torch.int64)   # dropout step counter (read as uint64)
        self._ones = torch.ones(1, device=dev, dtype=torch.float32)
        self._pool = None                 # graphs of different shapes replay one at a time: they share one memory pool
        self._gen = None                  # model._prep_gen the cached graphs were captured against
        self.replays = 0

    # ---- batch plumbing --------------------------------------------------------------------------------------
    def _inputs(self, batch):
        dev = self._hyper.device
        c = self.model.config
        out = {}
        for k in ("src_idx", "masks", "tgt_idx", "loss_masks"):
            out[k] = batch[k]
        if c.with_pho == "yes":
            out["pho_lens"] = hostio.lens_to_device(batch["pho_lens"], dev)   # a Python list in the reference's batches
            pho = batch["pho_idx"]
            if pho.shape[1] < MAX_PHO_LEN:    # pad_sequence pads to the batch maximum (src/utils.py:92-96): fix T = 7 so that
                pho = torch.nn.functional.pad(pho, (0, MAX_PHO_LEN - pho.shape[1]))   # every batch replays the same graph
            out["pho_idx"] = pho
        return {k: v.to(device=dev, dtype=(torch.int32 if k == "pho_lens" else v.dtype), non_blocking=True).contiguous()
                for k, v in out.items()}

    def _eager(self, inputs):
        loss = self.model(inputs)[0]
        loss.backward()
        self.opt.step()
        return loss.detach()

    def _capture(self, inputs):
        m, opt = self.model, self.opt
        eng = m._engine
        static = {k: v.clone() for k, v in inputs.items()}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()          # the eager step's activations go back to the driver; the graph gets its own pool
        graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        opt.device_hyper = self._hyper
        seed0, step0 = eng.seed, opt._step
        try:
            with ops.dropout_counter(self._counter), torch.cuda.graph(graph, pool=self._pool):
                self._counter.add_(1)
                loss, _ = eng.forward(static)
                eng.backward_and_sync(self._ones)
                opt.step()
        finally:
            opt.device_hyper = None
            eng.seed, opt._step = seed0, step0     # recording a graph is not a training step
        return graph, static, loss

    def _validate(self):
        """A captured graph bakes in device pointers of the model's operand cache (16-bit weight copies, the optimizer's
        shadow table).  model.eval() / .train() / .to() / load_state_dict / tie_cls_weight drop that cache: graphs
        captured against an older generation must never be replayed (they would read and WRITE freed memory)."""
        m = self.model
        if m._prepared is None:
            m.prepare()
        if self._gen != m._prep_gen:
            self._graphs.clear()
            self._pool = None             # a memory pool dies with its last graph: never hand a stale handle to a capture
            self._gen = m._prep_gen
            self.opt._sig = None          # the shadow pointers in the optimizer table are stale too
            if any(p.grad is not None for g in self.opt.param_groups for p in g["params"]):
                self.opt._build()         # outside any capture (host-to-device copies)

    def __call__(self, batch):
        m = self.model
        if not m.training:
            raise RuntimeError("GraphedTrainStep: model.train() first")
        self._validate()
        inputs = self._inputs(batch)
        key = tuple((k, tuple(v.shape)) for k, v in sorted(inputs.items()))
        entry = self._graphs.get(key)
        if entry is None:
            if key not in self._seen or len(self._graphs) >= self.max_graphs:
                self._seen.add(key)
                return self._eager(inputs)             # first step of a shape: eager (real step + warm-up)
            entry = self._graphs[key] = self._capture(inputs)
        graph, static, loss = entry
        for k, v in inputs.items():
            static[k].copy_(v, non_blocking=True)
        self.opt._step += 1
        self._hyper.copy_(torch.tensor(self.opt.hyper_values(self.opt._step), dtype=torch.float32), non_blocking=True)
        graph.replay()
        self.replays += 1
        return loss
