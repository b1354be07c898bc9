"""In-tree build of the C-ABI CUDA library (sm_100a only) and of the oracle's C pieces.

`python -m omni_avsr_b200.build` or `__graft_entry__.build()`.  The resulting
`omni_avsr_b200/libomni_avsr.so` is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libomni_avsr.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    """Content hash of the sources + flags.  Location-independent (the repo is copied to a scratch path on the GPU box:
    an absolute include path in the hash would force a rebuild there -- and a rebuild race between ranks)."""
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(f for f in NVCC_FLAGS if not os.path.isabs(f)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "stamp"
    dig = _digest(sources + headers)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    # one builder at a time (several ranks may import the package at once); the others wait, then find the stamp
    import fcntl
    with open(BUILD / "lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
                return LIB
            return _build_locked(sources, stamp, dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(sources, stamp, dig, verbose) -> Path:
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr}\n{r.stdout}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp), *[str(o) for o in objs], "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    os.replace(tmp, LIB)            # atomic: a process that already mapped the old library keeps its inode
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
