"""The C ABI from plain C (no Python, no torch in the process): compile tests/c/abi_harness.c
against include/amclj_b200.h and run it on the GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_c_harness(tmp_path):
    exe = tmp_path / "abi_harness"
    libdir = ROOT / "amclj_b200"
    subprocess.run(["gcc", "-O1", "-std=c11", str(ROOT / "tests/c/abi_harness.c"), "-I", str(ROOT / "include"),
                    "-L", str(libdir), "-lamclj_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "C_HARNESS_OK" in p.stdout


def test_c_harness_compiles(tmp_path):
    """CPU tier: the header is valid C11 and the harness links against the built library."""
    exe = tmp_path / "abi_harness"
    libdir = ROOT / "amclj_b200"
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-Werror", str(ROOT / "tests/c/abi_harness.c"), "-I",
                    str(ROOT / "include"), "-L", str(libdir), "-lamclj_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o",
                    str(exe)], check=True)
    assert exe.exists()
