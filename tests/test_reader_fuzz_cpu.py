"""The host-side file readers (MJCF, binary model, STL / OBJ, PNG and binary height fields) under AddressSanitizer +
UBSan on mutated inputs: malformed files must end in an error message, never in a crash, an overrun or a leak.  The
reference hands whatever path or string its caller names to the loader (mujoco_env.cpp:771-911).  Harness:
tools/fuzz_readers.cpp, corpus: tools/fuzz_seeds.py; a longer run is recorded in profiles/r2d_reader_fuzz.txt."""
import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-std=c++17", "-g0", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
         f"-I{ROOT}/include", f"-I{ROOT}/mujoco_ros_pkgs_b200/csrc"]


@pytest.mark.timeout(600)
def test_mutated_model_files_are_refused_cleanly(tmp_path):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    srcs = [f"{ROOT}/tools/fuzz_readers.cpp"] + sorted(glob.glob(f"{ROOT}/mujoco_ros_pkgs_b200/csrc/model/*.cpp"))

    def cc(src):
        obj = str(tmp_path / (os.path.basename(src) + ".o"))
        return obj, subprocess.run([cxx, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    with ThreadPoolExecutor(8) as ex:
        built = list(ex.map(cc, srcs))
    for obj, r in built:
        if r.returncode and "sanitize" in r.stderr and "unsupported" in r.stderr.lower():
            pytest.skip("this g++ has no sanitizer runtime")
        assert r.returncode == 0, r.stderr[-2000:]
    exe = str(tmp_path / "fuzz_readers")
    r = subprocess.run([cxx, *FLAGS, *[o for o, _ in built], "-o", exe], capture_output=True, text=True)
    if r.returncode and ("asan" in r.stderr or "ubsan" in r.stderr):
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    seeds = str(tmp_path / "seeds")
    subprocess.run([sys.executable, f"{ROOT}/tools/fuzz_seeds.py", seeds], check=True, capture_output=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:allocator_may_return_null=1")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([exe, seeds, "2500", "17"], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    words = r.stdout.split()
    loaded, refused = int(words[words.index("loaded,") - 1]), int(words[words.index("refused") - 1])
    assert loaded > 200 and refused > 200, r.stdout  # both outcomes exercised
