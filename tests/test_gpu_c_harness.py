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


def _build(tmp_path, name, extra=()):
    exe = tmp_path / name
    libdir = ROOT / "amclj_b200"
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-Werror", *extra, str(ROOT / f"tests/c/{name}.c"), "-I",
                    str(ROOT / "include"), "-L", str(libdir), "-lamclj_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o",
                    str(exe)], check=True)
    return exe


def test_mgpu_harness_compiles(tmp_path):
    assert _build(tmp_path, "mgpu_harness").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("n,devices", [(50_000, "0"), (200_000, "0 0"), (100_003, "0 0 0")])
def test_mgpu_harness_emulated_shards(tmp_path, n, devices):
    """amclj_create_multi from plain C, shards emulated on GPU 0: identical to one ctx, bit for bit."""
    exe = _build(tmp_path, "mgpu_harness")
    p = subprocess.run([str(exe), str(n), *devices.split()], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "MGPU_OK" in p.stdout


@pytest.mark.gpu
def test_mgpu_harness_all_visible_gpus(tmp_path):
    """The real thing when the box has several GPUs (gpurun --gpus N): one shard per device."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("one GPU visible: the emulated-shard test above covers the code path")
    exe = _build(tmp_path, "mgpu_harness")
    p = subprocess.run([str(exe), "1000000", *[str(d) for d in range(g)]], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "MGPU_OK" in p.stdout
