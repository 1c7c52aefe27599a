"""Build libpio_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m perceiverio_pytorch_b200.build [--force] [--verbose]

The library has no link-time dependency on libcuda or torch: it resolves cuTensorMapEncodeTiled through the
CUDA runtime's driver entry point at first use, so it loads (for symbol checks) on machines without a GPU.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpio_b200.so")
SOURCES = ["pio_host.cu", "pio_simt.cu", "pio_gemm.cu", "pio_gemm2.cu", "pio_flash.cu", "pio_flash2.cu", "pio_decode.cu", "pio_flash_qt.cu"]
HEADERS = ["pio_common.cuh", "pio_host.h", os.path.join("..", "..", "include", "pio_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"] + os.environ.get("PIO_NVCC_EXTRA", "").split()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    objs = []
    os.makedirs(os.path.join(PKG_DIR, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(PKG_DIR, "build", s.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for s, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
    with open(os.path.join(PKG_DIR, "build", "build.log"), "w") as f:
        f.write("\n".join(log))
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
