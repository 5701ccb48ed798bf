"""Compile librb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m remora_b200.build_native [--force]

nvcc cross-compiles without a GPU.  The .so lands in remora_b200/lib/ (git-ignored, shipped to the
GPU box with the snapshot)."""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librb200.so")
SOURCES = ["rb200_api.cu", "rb200_encode.cu", "rb200_layers.cu", "rb200_fused.cu", "rb200_chunks.cu",
           "rb200_tiled.cu", "rb200_refine.cu", "rb200_vbz.cu", "rb200_mega.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Wno-deprecated-gpu-targets",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _digest():
    h = hashlib.sha256()
    files = sorted(n for n in os.listdir(CSRC) if n.endswith((".cu", ".cuh", ".h"))) + \
        [os.path.join(ROOT, "include", "remora_b200.h")]
    for name in files:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        with open(path, "rb") as fh:
            h.update(name.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "librb200.sha256")
    digest = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp):
        if open(stamp).read().strip() == digest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v" if verbose else "-O3", "-c",
                                     os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        text = out.decode(errors="replace")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{text}\n")
        elif verbose:
            sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"], check=True)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
