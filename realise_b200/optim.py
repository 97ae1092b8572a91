"""Fused multi-tensor clip_grad_norm_ + AdamW on librealise_b200.so.

Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)` (src/run.py:207) followed by the
vendored `AdamW.step()` (transformers/optimization.py:113-169: bias-corrected Adam, decoupled weight decay applied
after the update) with two kernel launches over every parameter: sum of squared gradients, then the update which
reads the clip coefficient on device.  The same pass refreshes the bf16 operand copies the GEMMs consume.
"""
import ctypes

import torch

from ._lib import check, lib


class _Entry(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("shadow", ctypes.c_void_p), ("shadow32", ctypes.c_void_p), ("n", ctypes.c_int64), ("wd", ctypes.c_float),
                ("shadow_f16", ctypes.c_int32)]


CHUNK = 4096


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 correct_bias=True, model=None):
        """model: the realise_b200 model whose 16-bit operand copies the update kernel refreshes in place.  Without it
        the model notices the stale copies on its next forward (parameter version counters) and re-casts them."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.correct_bias = correct_bias
        self.model = model                # realise_b200 model: its bf16 / f32 operand copies are refreshed by the update
        self._step = 0
        self._sig = None
        self._sumsq = None
        self.device_hyper = None          # f32[3] device tensor {lr, 1-b1^t, 1-b2^t}: set by GraphedTrainStep so that a
                                          # captured step() reads its schedule from memory instead of kernel arguments

    def _build(self):
        entries, sig = [], []
        sh16, sh32 = ({}, {})
        if self.model is not None:
            if self.model._prepared is None:
                self.model.prepare()
            sh16, sh32 = self.model._shadow_bf16, self.model._shadow_f32
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                sh, s32 = sh16.get(id(p)), sh32.get(id(p))
                entries.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                sh.data_ptr() if sh is not None else 0, s32.data_ptr() if s32 is not None else 0,
                                p.numel(), float(group["weight_decay"]), int(sh is not None and sh.dtype is torch.float16)))
                sig.append((p.data_ptr(), p.grad.data_ptr(), entries[-1][4], entries[-1][5]))
        if sig == self._sig:
            return
        self._sig = sig
        dev = self.param_groups[0]["params"][0].device
        arr = (_Entry * len(entries))()
        chunks = []
        for i, (pp, g, m, v, sh, s32, n, wd, f16) in enumerate(entries):
            arr[i] = _Entry(pp, g, m, v, sh or None, s32 or None, n, wd, f16)
            chunks += [(i, c) for c in range((n + CHUNK - 1) // CHUNK)]
        self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int32).to(dev).contiguous()
        self._nchunks = len(chunks)
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        nws = int(lib().rl_workspace_bytes(b"mt_sumsq", ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0))) // 4
        self._partials = torch.zeros(max(nws, 1), device=dev, dtype=torch.float32)

    @torch.no_grad()
    def step(self, closure=None, grad_div=None):
        if grad_div is None:  # data parallel: gradients were summed over ranks (realise_b200.ddp.DataParallel)
            grad_div = float(getattr(self.model, "grad_div", 1.0)) if self.model is not None else 1.0
        self._build()
        self._step += 1
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            eng.grads_consumed()
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:          # one launch updates every tensor: the groups may differ in weight decay only
            if (g["lr"], tuple(g["betas"]), g["eps"]) != (g0["lr"], tuple(g0["betas"]), g0["eps"]):
                raise ValueError("FusedAdamW: lr / betas / eps must agree across param groups (only weight_decay is per group, "
                                 "as in src/run.py:146-152)")
        b1, b2 = g0["betas"]
        bc1 = 1.0 - b1 ** self._step if self.correct_bias else 1.0
        bc2 = 1.0 - b2 ** self._step if self.correct_bias else 1.0
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        tab, ck = ctypes.c_void_p(self._table.data_ptr()), ctypes.c_void_p(self._chunks.data_ptr())
        ss = ctypes.c_void_p(self._sumsq.data_ptr())
        f = ctypes.c_float
        check(lib().rl_mt_sumsq(tab, ck, ctypes.c_int64(self._nchunks), ss, ctypes.c_void_p(self._partials.data_ptr()),
                                ctypes.c_int64(self._partials.numel()), st), "rl_mt_sumsq")
        if self.device_hyper is not None:
            check(lib().rl_mt_adamw_dev(tab, ck, ctypes.c_int64(self._nchunks), ss, f(self.max_grad_norm or 0.0),
                                        ctypes.c_void_p(self.device_hyper.data_ptr()), f(b1), f(b2), f(g0["eps"]),
                                        f(grad_div), st), "rl_mt_adamw_dev")
            return
        check(lib().rl_mt_adamw(tab, ck, ctypes.c_int64(self._nchunks), ss, f(self.max_grad_norm or 0.0), f(g0["lr"]),
                                f(b1), f(b2), f(g0["eps"]), f(bc1), f(bc2), f(grad_div), st), "rl_mt_adamw")

    def state_dict(self):
        """torch layout plus the shared step counter (bias correction resumes where it stopped); every state entry also
        carries 'step' like the vendored AdamW's (transformers/optimization.py:136-140)."""
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = self._step
        sd = super().state_dict()
        sd["fused_step"] = self._step
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        step = state_dict.pop("fused_step", None)
        super().load_state_dict(state_dict)
        if step is None:     # a vendored-AdamW / torch AdamW checkpoint: per-parameter 'step'
            steps = [int(st["step"]) for st in self.state.values() if "step" in st]
            step = max(steps) if steps else 0
        self._step = int(step)
        self._sig = None     # moment tensors were replaced: rebuild the device table

    def hyper_values(self, step):
        """{lr, 1 - beta1^t, 1 - beta2^t} of optimizer step `step` (1-based) for the device-resident schedule."""
        g0 = self.param_groups[0]
        b1, b2 = g0["betas"]
        if not self.correct_bias:
            return [float(g0["lr"]), 1.0, 1.0]
        return [float(g0["lr"]), 1.0 - b1 ** step, 1.0 - b2 ** step]

    def grad_norm(self):
        """Global L2 norm of the last step's gradients (device scalar -> host)."""
        return float(self._sumsq.sqrt().item())
