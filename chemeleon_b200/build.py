"""Build libchemeleon_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m chemeleon_b200.build            # build if sources are newer than the library
    python -m chemeleon_b200.build --force
"""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libchemeleon_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def have_nvcc() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


HASH_FILE = LIB + ".srchash"


def _deps():
    return sorted(sources() + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def source_hash() -> str:
    """Content hash of everything the library is built from (file mtimes do not survive a copy of the tree).
    CB2_NVCC_EXTRA (development builds, e.g. -DCB2_EDGE_TIMELINE) is deliberately not part of it: such a
    library is built with --force and must not be rebuilt by the next process that has a clean environment."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASH_FILE):
        return True
    with open(HASH_FILE) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu and link the library.  Safe under concurrent callers (one process per GPU all
    importing the package): an exclusive file lock serialises the build, the library is linked under a
    temporary name and renamed into place, so no process ever maps a half-written file."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libchemeleon_b200.so must be prebuilt (it travels with the tree)")
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    with open(os.path.join(build_dir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it while we waited
                return LIB
            want = source_hash()
            objs, procs = [], []
            for src in sources():
                obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
                objs.append(obj)
                cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("CB2_NVCC_EXTRA", "").split() + \
                    ["-c", src, "-o", obj]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            for src, p in procs:
                out, _ = p.communicate()
                if p.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}:\n{out}")
                if verbose and out.strip():
                    print(out)
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}")
            os.replace(tmp, LIB)
            with open(HASH_FILE + ".tmp", "w") as f:
                f.write(want + "\n")
            os.replace(HASH_FILE + ".tmp", HASH_FILE)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
