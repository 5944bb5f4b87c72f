"""Build libanemoi_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libanemoi_b200.so")
WIP_SOURCES: list = []  # sources that must keep compiling but are not linked (none at present)
SOURCES = ["abi.cu", "csr_build.cu", "gtconv.cu", "gtconv_tma.cu", "gtconv_fold.cu", "graphconv.cu", "host_api.cu", "peer_exchange.cu",
           "gemm_tc.cu", "layernorm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libanemoi_b200.so")


def _digest(paths, extra: str = "") -> str:
    """sha256 over the CONTENT of the files (and the flags): a shipped .so / .o is reused only when it was built from exactly
    these sources -- modification times do not survive a snapshot copy and can make a stale object look fresh."""
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(hashlib.sha256(f.read()).digest())
    return h.hexdigest()


def _stale(target: str, digest: str) -> bool:
    stamp = target + ".sha256"
    if not (os.path.exists(target) and os.path.exists(stamp)):
        return True
    with open(stamp) as f:
        return f.read().strip() != digest


def _stamp(target: str, digest: str) -> None:
    with open(target + ".sha256", "w") as f:
        f.write(digest + "\n")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu of csrc/ (in parallel) and link the shared library.  Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "gtconv_args.cuh"),
               os.path.join(os.path.dirname(HERE), "include", "anemoi_b200.h")]
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace("/", "_").replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        digest = _digest([path] + headers, " ".join(NVCC_FLAGS))
        if force or _stale(obj, digest):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, path]
            if verbose:
                print(" ".join(cmd))
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
            _stamp(obj, digest)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES) + len(WIP_SOURCES)) as ex:
        wip = [ex.submit(compile_one, src) for src in WIP_SOURCES]  # compile check only: NOT linked into the library
        objs = list(ex.map(compile_one, SOURCES))
        for f in wip:
            f.result()
    digest = _digest([o + ".sha256" for o in objs])
    if force or _stale(LIBPATH, digest):
        cmd = [nvcc, "-shared", "-o", LIBPATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        _stamp(LIBPATH, digest)
    return LIBPATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
