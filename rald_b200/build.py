"""Builds rald_b200/_C/librald_b200.so from rald_b200/csrc/*.cu with nvcc for sm_100a.

In-tree, incremental (per-file object cache keyed on source + header mtimes and a hash of the compiler flags),
parallel, serialised across processes by a file lock. The built
library travels to the GPU box with the repository snapshot; nothing is JIT-compiled at run time.

    python -m rald_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_C"
OBJ_DIR = OUT_DIR / "obj"
LIB_PATH = OUT_DIR / "librald_b200.so"
INCLUDE = ROOT.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    f"-I{INCLUDE}",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the rald_b200 CUDA library cannot be built")
    return exe


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def _compile_one(src: Path, obj: Path, verbose: bool) -> str:
    tmp = obj.with_suffix(f".{os.getpid()}.tmp.o")
    cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("RALD_B200_NVCC_FLAGS", "").split(), "-c", str(src), "-o", str(tmp)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, obj)
    (obj.with_suffix(".ptxas.txt")).write_text(res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return src.name


def _flags_key() -> str:
    """Hash of everything besides the sources that decides what an object file contains: the nvcc command line
    (incl. RALD_B200_NVCC_FLAGS, e.g. -DRALD_GELU_LOGISTIC=0) and the compiler version."""
    extra = os.environ.get("RALD_B200_NVCC_FLAGS", "")
    try:
        ver = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout
    except OSError:
        ver = ""
    return hashlib.sha256(("\0".join(NVCC_FLAGS) + "\0" + extra + "\0" + ver).encode()).hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Serialised across processes by an exclusive file lock (every rank of a torchrun job may find the library
    missing at the same time); objects and the library are written to temporaries and renamed into place."""
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    with open(OUT_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    key = _flags_key()
    key_file = OBJ_DIR / ".flags"
    if not key_file.exists() or key_file.read_text() != key:
        force = True   # objects built with other flags (or another nvcc) are stale whatever their mtimes say
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    hdr_m = _newest_header_mtime()
    todo = []
    objs = []
    for s in srcs:
        o = OBJ_DIR / (s.stem + ".o")
        objs.append(o)
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_m):
            todo.append((s, o))
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for name in ex.map(lambda a: _compile_one(a[0], a[1], verbose), todo):
                if verbose:
                    print(f"[rald_b200.build] compiled {name}")
        key_file.write_text(key)
    if todo or not LIB_PATH.exists():
        tmp = LIB_PATH.with_suffix(f".{os.getpid()}.tmp.so")
        cmd = [_nvcc(), "-shared", "-o", str(tmp), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
