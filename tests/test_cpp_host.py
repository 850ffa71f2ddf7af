"""C++ host-side mirror (include/b2mj_env.hpp: BatchEnv / BatchPlugin / BatchData) — compiled with g++ against
libb2mj.so and run as a native test binary (tests/cpp/test_batch_env.cpp mirrors the reference's gtests
mujoco_env_test.cpp, mujoco_ros_plugin_test.cpp and ros_interface_test.cpp for the stepping surface)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mujoco_ros_pkgs_b200")
BIN = os.path.join(ROOT, "build", "tests", "test_batch_env")
SRC = os.path.join(ROOT, "tests", "cpp", "test_batch_env.cpp")
HDRS = [os.path.join(ROOT, "include", h) for h in ("b2mj.h", "b2mj_env.hpp", "b2mj_plugins.hpp")]


def build_binary():
    deps = [SRC, os.path.join(PKG, "libb2mj.so")] + HDRS
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return BIN
    cxx = shutil.which("g++") or "/opt/gcc/bin/g++"
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", BIN,
           "-L" + PKG, "-lb2mj", "-Wl,-rpath," + PKG, "-pthread"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return BIN


def test_header_compiles_and_fails_loudly_without_gpu():
    """The C++ mirror builds warning-free against the C-ABI; without a CUDA device loading a model must fail with
    an error (no CPU fallback), which the binary reports as NO_DEVICE."""
    exe = build_binary()
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe, os.path.join(PKG, "models")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NO_DEVICE load=0" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("nenv", [1, 4, 33])
def test_batch_env_mirror_on_gpu(nenv):
    exe = build_binary()
    r = subprocess.run([exe, os.path.join(PKG, "models"), str(nenv)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1].startswith("OK")
