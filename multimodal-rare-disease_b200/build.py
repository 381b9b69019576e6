"""Build the C-ABI CUDA library (libmrd_b200.so) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box as well as on a B200 box.
The .so lands next to this file (git-ignored, but shipped to the GPU box with the snapshot).
"""

from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libmrd_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
STAMP_PATH = os.path.join(HERE, ".libmrd_b200.stamp")
LOCK_PATH = os.path.join(HERE, ".libmrd_b200.lock")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    inc = os.path.join(os.path.dirname(HERE), "include", "mrd_b200.h")
    if os.path.exists(inc):
        with open(inc, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == _fingerprint()


def _compile_and_link(nvcc: str, verbose: bool) -> None:
    """nvcc every csrc/*.cu and link, all into per-process temporary names, then rename atomically: a
    concurrent reader (another rank that already holds a current library) never sees a half-written file."""
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    inc_dir = os.path.join(os.path.dirname(HERE), "include")
    tag = f".{os.getpid()}"
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + tag + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", inc_dir, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    tmp = LIB_PATH + tag + ".tmp"
    try:
        failed = None
        for src, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode != 0:
                sys.stderr.write(out.decode(errors="replace"))
            if p.returncode != 0 and failed is None:
                failed = src
        if failed:
            raise RuntimeError(f"nvcc failed on {failed}")
        link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            sys.stderr.write(r.stdout.decode(errors="replace"))
            raise RuntimeError("nvcc link failed")
        os.replace(tmp, LIB_PATH)
        stamp_tmp = STAMP_PATH + tag
        with open(stamp_tmp, "w") as fh:
            fh.write(_fingerprint())
        os.replace(stamp_tmp, STAMP_PATH)
    finally:
        for f in objs + [tmp]:
            try:
                os.remove(f)
            except OSError:
                pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu into libmrd_b200.so.  Returns the library path.

    Safe under torchrun: the build is serialised by an exclusive file lock and the freshness check is repeated
    once the lock is held, so N ranks starting with a stale library compile it once, not N times into the
    same files."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            # no compiler on this box: the prebuilt library shipped with the snapshot is all there is.  Say so
            # when it does not match the sources next to it (the coarse ABI check is the only other guard).
            if os.path.exists(STAMP_PATH):
                with open(STAMP_PATH) as fh:
                    if fh.read().strip() != _fingerprint():
                        warnings.warn("libmrd_b200.so was built from different sources than the ones in csrc/ and nvcc "
                                      "is not available to rebuild it; using the stale library", RuntimeWarning)
            return LIB_PATH
        raise RuntimeError("nvcc not found and no prebuilt libmrd_b200.so present")
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or not is_current():   # another process may have finished the build while we waited
                _compile_and_link(nvcc, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
