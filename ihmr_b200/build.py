"""Builds libihmr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m ihmr_b200.build [--force] [--verbose]

The library is the C-ABI of include/ihmr_b200.h; nothing here depends on torch.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libihmr_b200.so")
SOURCES = ["abi.cu", "mano.cu", "sdf.cu", "opt.cu", "blend_tc.cu", "eval.cu", "mlp.cu"]
HEADERS = ["common.cuh", "kernels.cuh", os.path.join("..", "..", "include", "ihmr_b200.h")]

# Extra libraries for the tests: name -> (path, source recompiled with extra flags).  `smallcaps` shrinks the shared-memory
# capacities of the penetration kernel so that ordinary frames take every multi-pass / overflow branch.
VARIANTS = {
    "smallcaps": (os.path.join(OUT_DIR, "libihmr_b200_smallcaps.so"), "sdf.cu",
                  ["-DSDF_PHI_CAP=64", "-DSDF_Q_CAP=512"]),
}

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build ihmr_b200")


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _build_variants(nvcc, ccbin, objs, verbose):
    for name, (path, src, flags) in VARIANTS.items():
        obj = os.path.join(OUT_DIR, f"{name}_{src.replace('.cu', '.o')}")
        cmd = [nvcc, *ccbin, *NVCC_FLAGS, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        others = [o for o in objs if os.path.basename(o) != src.replace(".cu", ".o")]
        subprocess.check_call([nvcc, *ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", obj, *others, "-o", path])


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp_file = os.path.join(OUT_DIR, "stamp.txt")
    stamp = _stamp()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp_file) and all(os.path.exists(v[0]) for v in VARIANTS.values()):
        if open(stamp_file).read().strip() == stamp:
            return LIB_PATH
    nvcc = _nvcc()
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ccbin, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"---- {src}\n{out}", flush=True)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, *ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB_PATH]
    subprocess.check_call(link)
    _build_variants(nvcc, ccbin, objs, verbose)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
