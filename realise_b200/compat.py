"""Compatibility layer that lets the reference's own drivers (src/run.py via train.sh, src/test.py via test.sh) run
UNCHANGED on a current PyTorch with realise_b200 behind the model-class API (last row).

`install()` is called by the shim modules `shim/src/models.py` / `shim/src/models_abla.py` (which replace the two files
of the same name in a reference checkout) at import time — i.e. before src/run.py parses its arguments:

  * `--local-rank=N` (what torch.distributed.launch / torchrun pass since torch 2.0) is rewritten to the `--local_rank=N`
    that src/run.py:368 declares; when neither is given but LOCAL_RANK is in the environment (torchrun) it is appended.
  * `torch.load` defaults to `weights_only=False` again for the pickled argparse namespace `training_args.bin`
    (src/test.py:105, src/run.py:229) — the reference was written against torch 1.2 where that was the only behaviour.
  * `torch.nn.parallel.DistributedDataParallel(model, ...)` (src/run.py:165-167) returns, for a realise_b200 model, a thin
    wrapper with the same surface (`.module`, `__call__`, `.train()/.eval()`, `.parameters()`, `.zero_grad()`): the
    gradient exchange is ONE NCCL all-reduce of the engine's flat buffer after the backward (realise_b200.ddp), not
    autograd hooks on 400 parameters.  Gradients are averaged over ranks like DDP's.
  * third-party modules the reference imports but never uses on this path (`torchcrf`, `boto3`/`botocore`, `sacremoses`)
    are stubbed when absent (transformers/modeling_bert.py:25, transformers/file_utils.py:20-22).
"""
import os
import sys
import types

import torch

_INSTALLED = False


def fix_argv(argv, environ=os.environ):
    """`--local-rank` -> `--local_rank` (both `--x=N` and `--x N` forms); torchrun's LOCAL_RANK as a fallback."""
    out, seen = [], False
    for a in argv:
        if a == "--local-rank" or a.startswith("--local-rank="):
            a = "--local_rank" + a[len("--local-rank"):]
        seen = seen or a == "--local_rank" or a.startswith("--local_rank=")
        out.append(a)
    if not seen and "LOCAL_RANK" in environ and int(environ.get("WORLD_SIZE", "1")) > 1 and len(out) > 0 \
            and os.path.basename(out[0]) == "run.py":
        out.append(f"--local_rank={environ['LOCAL_RANK']}")
    return out


class DistributedModel(torch.nn.Module):
    """What `DistributedDataParallel(model, device_ids=[rank], ...)` returns for a realise_b200 model."""

    def __init__(self, module, process_group=None):
        super().__init__()
        from .ddp import DataParallel
        self.module = module
        self._dp = DataParallel(module, group=process_group, average=True)
        self._dp.broadcast_parameters()

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _ddp_factory(orig):
    def make(module, *a, **k):
        from .model import SpellBertPho2ResArch3Abla
        if isinstance(module, SpellBertPho2ResArch3Abla):
            return DistributedModel(module, process_group=k.get("process_group"))
        return orig(module, *a, **k)
    make.__wrapped__ = orig
    return make


def _stub(name, **attrs):
    if name in sys.modules:
        return
    try:
        __import__(name)
    except Exception:  # noqa: BLE001 — absent or broken: the reference only needs the name to exist
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    sys.argv[:] = fix_argv(sys.argv)
    orig_load = torch.load

    def load(*a, **k):
        k.setdefault("weights_only", False)
        return orig_load(*a, **k)

    load.__wrapped__ = orig_load
    torch.load = load
    ddp = torch.nn.parallel.DistributedDataParallel
    if not hasattr(ddp, "__wrapped__"):
        torch.nn.parallel.DistributedDataParallel = _ddp_factory(ddp)
    _stub("torchcrf", CRF=object)
    _stub("sacremoses")
    _stub("boto3")
    _stub("botocore")
    if isinstance(sys.modules.get("botocore"), types.ModuleType) and not hasattr(sys.modules["botocore"], "exceptions"):
        ex, cf = types.ModuleType("botocore.exceptions"), types.ModuleType("botocore.config")
        ex.ClientError, cf.Config = Exception, object
        sys.modules["botocore.exceptions"], sys.modules["botocore.config"] = ex, cf
        sys.modules["botocore"].exceptions, sys.modules["botocore"].config = ex, cf
