"""In-tree build of libivfadc_cuda.so (sm_100a only) with nvcc.  No GPU is needed to build."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libivfadc_cuda.so")
SOURCES = ["api.cu", "coarse.cu", "scan.cu", "encode.cu", "lists.cu", "shard.cu", "synth.cu", "train.cu"]
HEADERS = ["common.cuh", "warp_topk.cuh", "scan_impl.cuh", "scanq_impl.cuh", "tc_common.cuh", "scanu_impl.cuh", "scanw_impl.cuh", "coarse_tc.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    # bring-up: IVFADC_NVCC_EXTRA="-DX=1" IVFADC_LIB_OUT=/path/lib.so builds a variant beside the product library
    extra = os.environ.get("IVFADC_NVCC_EXTRA", "").split()
    out = os.environ.get("IVFADC_LIB_OUT")
    if extra or out:
        return _build_variant(extra, out or LIB + ".variant")
    os.makedirs(OBJDIR, exist_ok=True)
    common = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(ROOT, "include", "ivfadc.h")]
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + common):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


def _build_variant(extra, out: str) -> str:
    import tempfile
    nvcc = _nvcc()
    with tempfile.TemporaryDirectory() as tmp:
        def one(src):
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            subprocess.check_call([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj])
            return obj
        with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
            objs = list(ex.map(one, SOURCES))
        subprocess.check_call([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
