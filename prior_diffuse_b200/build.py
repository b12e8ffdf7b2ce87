"""Compile csrc/*.cu (sm_100a device code) and csrc/*.cpp (host-only C++) into prior_diffuse_b200/libpdse.so, in-tree,
so the .so travels with the repo snapshot to the GPU box.  One object per source, compiled in parallel and cached under
build/ by content hash of the source and the headers, then linked (``force=True`` recompiles everything)."""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpdse.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
if os.environ.get("PDSE_OPERANDS", "").lower() == "bf16":     # csrc/opfmt.h: fp16 operands unless asked otherwise
    NVCC_FLAGS.append("-DPDSE_OP_BF16")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + headers())


def _digest(path: str, hdr: bytes) -> str:
    h = hashlib.sha1(hdr)
    h.update(" ".join(NVCC_FLAGS).encode())
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ, exist_ok=True)
    hdr = b"".join(open(p, "rb").read() for p in headers())
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src) + "." + _digest(src, hdr) + ".o")
        jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        if os.path.exists(obj) and not verbose and not force:
            return ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj + ".tmp", src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + r.stdout + r.stderr)
        os.replace(obj + ".tmp", obj)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        logs = list(ex.map(compile_one, jobs))
    keep = {obj for _, obj in jobs}
    for old in glob.glob(os.path.join(OBJ, "*.o")):
        if old not in keep:
            os.remove(old)
    r = subprocess.run([nvcc, "-shared", "-o", LIB + ".tmp"] + [obj for _, obj in jobs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
