"""The boundary is a C ABI: include/tasr.h compiles as C (gcc -std=c99 -pedantic), a C program can dlopen
libtasr_b200.so and bind its entry points, and argument validation answers before any CUDA call (so this runs on
the CPU-only box).  Also checks that the ctypes mirror of the structs has the C sizes."""
import ctypes
import os
import shutil
import subprocess

import pytest

from telugu_asr_b200 import _native

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_header_compiles_as_c_and_c_program_binds_the_library(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    _native.lib()                                   # builds the library if needed
    exe = tmp_path / "abi_check"
    src = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", str(exe), "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), _native.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI OK" in r.stdout
    sizes = dict(kv.split("=") for kv in r.stdout.replace(",", "").split() if "=" in kv)
    assert int(sizes["sizeof(TasrFeatParams)"]) == ctypes.sizeof(_native.TasrFeatParams)
    assert int(sizes["sizeof(TasrSepConvLayer)"]) == ctypes.sizeof(_native.TasrSepConvLayer)
