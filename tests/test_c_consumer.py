"""A plain-C program (tests/c/abi_smoke.c) compiled with gcc against include/arrowgpu.h and linked
with libarrowgpu.so: the boundary is a C ABI any FFI can bind.  On CPU it must build, link, run and
be refused by the library (no CPU fallback); on the GPU it must pass its own checks."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "abi_smoke")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "arrow_go_b200", "lib")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
           "-o", EXE, "-L", lib, "-larrowgpu", f"-Wl,-rpath,{lib}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return EXE


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_c_program_builds_links_and_is_refused_without_gpu():
    out = subprocess.run([build()], capture_output=True, text=True, timeout=120)
    assert out.returncode == 77, (out.returncode, out.stdout, out.stderr)
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_program_on_gpu():
    out = subprocess.run([build()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi_smoke ok" in out.stdout
