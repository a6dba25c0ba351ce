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


def _headers_digest() -> bytes:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".inc"))]
    files.append(os.path.join(HERE, "..", "include", "tasr.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into telugu_asr_b200/libtasr_b200.so; returns its path.

    Each source is compiled to an object under telugu_asr_b200/.build/ (keyed by its own text, the headers and the flags;
    the compiles run in parallel) and the objects are linked by nvcc: an edit to one kernel recompiles one file."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, ".build")
    os.makedirs(objdir, exist_ok=True)
    hd = _headers_digest()
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    compile_flags = [f for i, f in enumerate(compile_flags) if not (f == "-cudart" or (i > 0 and compile_flags[i - 1] == "-cudart"))]
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        with open(src, "rb") as fh:
            key = hashlib.sha256(hd + fh.read()).hexdigest()[:20]
        obj = os.path.join(objdir, f"{os.path.splitext(s)[0]}.{key}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            for stale in os.listdir(objdir):
                if stale.startswith(os.path.splitext(s)[0] + ".") and stale.endswith(".o"):
                    os.remove(os.path.join(objdir, stale))
            cmd = [nvcc, *compile_flags, "-c"]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            cmd += ["-o", obj, src]
            jobs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    failed = []
    for s, obj, proc in jobs:
        out, err = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out + err)
        if proc.returncode != 0:
            failed.append((s, err))
            if os.path.exists(obj):
                os.remove(obj)
    if failed:
        raise RuntimeError("nvcc failed building libtasr_b200.so:\n" + "\n".join(f"[{s}]\n{e[-4000:]}" for s, e in failed))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", LIB, *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libtasr_b200.so:\n" + res.stderr[-4000:])
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
