#!/usr/bin/env python
"""Installs the UNMODIFIED reference into git-ignored baseline/_ref (it travels to the GPU box with the repo snapshot;
/root/reference itself does not exist there).  Run once in the build container:

    python baseline/install_ref.py

  1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>` — the
     reference's setup.py packages its vendored `transformers` 2.2.2 (the copy under /tmp is needed because the build
     writes into the source tree and /root/reference is read-only; --no-deps because boto3 / sacremoses / ... are not in
     the offline wheelhouse and none of them is on the arithmetic path);
  2. the reference's model scripts are not part of that package (train.sh runs them as `python src/run.py`), so the four
     files the hot path lives in — src/models.py, src/models_abla.py, src/char_cnn.py, src/utils.py — are copied next to
     it as baseline/_ref/src/.
Nothing under baseline/_ref is tracked by git (see .gitignore); bench.py --impl reference and the `incumbent` leg import
it through baseline/ref_model.py.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("REALISE_REFERENCE", "/root/reference")


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        sys.exit(f"{REF} not found: the reference can only be installed where its checkout exists")
    shutil.rmtree(DEST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(REF, copy, ignore=shutil.ignore_patterns(".git", "*.ttf", "assets"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
               "--no-deps", "--target", DEST, copy]
        subprocess.check_call(cmd)
    os.makedirs(os.path.join(DEST, "src"), exist_ok=True)
    for f in ("models.py", "models_abla.py", "char_cnn.py", "utils.py"):
        shutil.copy(os.path.join(REF, "src", f), os.path.join(DEST, "src", f))
    print("installed:", sorted(os.listdir(DEST)))


if __name__ == "__main__":
    main()
