"""In-tree build of libtasr_b200.so (sm_100a only) with nvcc.

    python -m telugu_asr_b200.build [--force] [--verbose]

nvcc cross-compiles here without a GPU; the .so is git-ignored but travels to the GPU box
with the gpurun snapshot.  One -gencode, no other architecture, no JIT cache.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtasr_b200.so")
STAMP = os.path.join(HERE, ".libtasr_b200.stamp")
SOURCES = ["api.cu", "absmax.cu", "conv2d_subsample.cu", "encoder_block.cu", "feature_post.cu", "ingest.cu", "logmel.cu", "logmel_generic.cu", "logmel_tc.cu", "sepconv_fp32.cu", "sepconv_tf32.cu", "sepconv_ws.cu", "specaugment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtasr_b200.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    files = [os.path.join(CSRC, s) for s in SOURCES]
    files += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".inc"))]
    files.append(os.path.join(HERE, "..", "include", "tasr.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into telugu_asr_b200/libtasr_b200.so; returns its path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtasr_b200.so:\n" + res.stderr[-4000:])
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
