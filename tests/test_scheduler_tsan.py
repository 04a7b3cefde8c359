"""The C++ scheduler under ThreadSanitizer (the reference wires up no sanitizer). Builds
tests/cpp/scheduler_stress.cpp + csrc/scheduler.cpp with -fsanitize=thread and runs it on the CPU backend."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sanitizer", ["thread", "address"])
def test_scheduler_stress_under_sanitizer(tmp_path, sanitizer):
    cxx = shutil.which("g++")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if cxx is None or not os.path.isdir(os.path.join(cuda, "include")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = tmp_path / f"stress_{sanitizer}"
    cmd = [cxx, "-std=c++17", "-O1", "-g", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", f"-I{REPO}/bagua_b200/csrc", f"-I{cuda}/include",
           f"{REPO}/tests/cpp/scheduler_stress.cpp", f"{REPO}/bagua_b200/csrc/scheduler.cpp", f"-L{cuda}/lib64", "-lcudart", "-ldl", "-lpthread", "-o", str(exe)]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip(f"sanitizer build unavailable here: {build.stderr[-400:]}")
    env = dict(os.environ, LD_LIBRARY_PATH=f"{cuda}/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""), TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=0")
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert "0 failures" in run.stdout
    assert "WARNING: ThreadSanitizer" not in run.stderr and "ERROR: AddressSanitizer" not in run.stderr


@pytest.mark.parametrize("sanitizer", ["thread", "address"])
def test_net_transport_stress_under_sanitizer(tmp_path, sanitizer):
    """The NCCL net plugin's transport engine (csrc/net) under TSAN / ASAN over loopback."""
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("needs g++")
    exe = tmp_path / f"net_stress_{sanitizer}"
    cmd = [cxx, "-std=c++17", "-O1", "-g", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", f"-I{REPO}/bagua_b200/csrc/net",
           f"{REPO}/tests/cpp/net_stress.cpp", f"{REPO}/bagua_b200/csrc/net/net_engine.cpp", "-lpthread", "-o", str(exe)]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip(f"sanitizer build unavailable here: {build.stderr[-400:]}")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=1")
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert "0 failures" in run.stdout
    assert "WARNING: ThreadSanitizer" not in run.stderr and "ERROR: AddressSanitizer" not in run.stderr
