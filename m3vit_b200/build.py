"""Builds the C-ABI shared library `m3vit_b200/lib/libm3vit_moe.so` with nvcc for
sm_100a.  In-tree on purpose: the built .so travels to the GPU box with the repo
snapshot (a JIT cache under ~/.cache would not)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libm3vit_moe.so")
SOURCES = ["abi.cu", "gate.cu", "route.cu", "permute.cu", "ffn_f32.cu", "ffn_bf16.cu", "ffn_chain.cu", "block.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",     # no --use_fast_math: exact erf/exp/div, parity first
]
TRACE = os.environ.get("M3_GEMM_TRACE") == "1"
VARIANT = "trace" if TRACE else os.environ.get("M3_BUILD_VARIANT", "")
if TRACE:      # clock64 timeline in the tensor-core kernels (tools/gemm_timeline.py, tools/chain_timeline.py): a SEPARATE
    NVCC_FLAGS.append("-DM3_GEMM_TRACE")       # library, loaded with M3_LIB_PATH=.../libm3vit_moe_trace.so
elif VARIANT:  # A/B build with extra defines (tools/ab_libs.py): M3_BUILD_VARIANT=name M3_BUILD_DEFS="-DX=1 -DY=2"
    NVCC_FLAGS += os.environ.get("M3_BUILD_DEFS", "").split()
if VARIANT:
    LIB = os.path.join(LIBDIR, f"libm3vit_moe_{VARIANT}.so")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "m3vit_moe.h")]
    stamp = os.path.join(LIBDIR, f"build_{VARIANT}.sha256" if VARIANT else "build.sha256")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, "obj_trace" if TRACE else os.path.join("obj", VARIANT) if VARIANT else "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
