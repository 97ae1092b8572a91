"""Drop-in replacement for the reference's src/models_abla.py (see models.py in this directory)."""
import os
import sys

_ROOT = os.environ.get("REALISE_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.path.isdir(os.path.join(_ROOT, "realise_b200")) and _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from realise_b200 import compat as _compat  # noqa: E402

_compat.install()

from realise_b200.model import SpellBertPho2ResArch3Abla  # noqa: E402,F401
