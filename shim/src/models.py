"""Drop-in replacement for the reference's src/models.py: copy this file (and models_abla.py next to it) over the files
of the same name in a ReaLiSe checkout and `sh train.sh` / `sh test.sh` run unchanged with the B200-native
implementation behind `SpellBertPho2ResArch3` (the class train.sh selects with --model_type bert-pho2-res-arch3).

src/run.py:26-31 imports nine model classes from this module.  Only SpellBertPho2ResArch3 (and its ablation class in
models_abla.py) is on the path this package accelerates; the other research variants raise on
construction and tell the user to keep the reference file for them.
"""
import os
import sys

_ROOT = os.environ.get("REALISE_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.path.isdir(os.path.join(_ROOT, "realise_b200")) and _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)          # a copy inside a reference checkout finds the package through REALISE_B200_ROOT

from realise_b200 import compat as _compat  # noqa: E402

_compat.install()                      # argv / torch.load / DistributedDataParallel / third-party stubs (see compat.py)

from realise_b200.glyphs import is_chinese_char as _is_chinese_char  # noqa: E402,F401
from realise_b200.model import SpellBertPho2ResArch3  # noqa: E402,F401


def _out_of_scope(name):
    class _Unsupported:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is not on the realise_b200 hot path (SpellBertPho2ResArch3 and its "
                                      "ablations are); use the reference's own src/models.py for it")

        @classmethod
        def from_pretrained(cls, *a, **k):
            cls()

    _Unsupported.__name__ = _Unsupported.__qualname__ = name
    return _Unsupported


for _n in ("SpellBert", "SpellBertPho1", "SpellBertPho2", "SpellBertPho1Res", "SpellBertPho2Res", "SpellBertPho2ResArch2",
           "SpellBertPho2ResArch3MLM", "SpellBertPho2ResArch4"):
    globals()[_n] = _out_of_scope(_n)
