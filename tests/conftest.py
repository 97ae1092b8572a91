import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _release_gpu_memory_between_tests():
    """The benchmark-shape parity tests hold > 100 GB in the caching allocator; hand it back so that later tests (the
    two-rank worker processes of test_ddp_gpu.py share this GPU) start from a clean device."""
    yield
    if torch.cuda.is_available():
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(str(g["meta"]))
    return g, meta


_SD_CACHE = {}


def cached_state_dict(cfg, seed):
    """synth_state_dict is ~6 s for the full model: share it between tests of one session."""
    from realise_b200.synth import synth_state_dict
    key = (tuple(sorted(cfg.__dict__.items())), seed)
    if key not in _SD_CACHE:
        if len(_SD_CACHE) >= 2:
            _SD_CACHE.pop(next(iter(_SD_CACHE)))
        _SD_CACHE[key] = synth_state_dict(cfg, seed)
    return _SD_CACHE[key]
