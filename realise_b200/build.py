"""Builds librealise_b200.so (the sm_100a kernels + C ABI) in-tree with nvcc.

The .so lives next to this file so that it travels to the GPU box with the repo snapshot.
`python -m realise_b200.build` rebuilds; `ensure_built()` rebuilds only when a source is newer.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "librealise_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(os.path.join(r, f)) for r in (CSRC, INCLUDE)
                    for f in os.listdir(r) if f.endswith((".cuh", ".h")))
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_mtime)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("librealise_b200.so: nvcc compilation failed")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


def _stale():
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def ensure_built():
    if _stale():
        if subprocess.call(["which", os.environ.get("NVCC", "nvcc")], stdout=subprocess.DEVNULL) != 0:
            if os.path.exists(LIB):
                return LIB  # GPU box without a fresher build: use what travelled
            raise RuntimeError("librealise_b200.so is missing and nvcc is not available")
        # one process builds, the others (torchrun ranks importing at the same moment) wait for it
        import fcntl
        with open(os.path.join(HERE, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if _stale():
                    build()
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
