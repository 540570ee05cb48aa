"""Builds tests/cpp/th_api_test.cpp against include/th/*.hpp + libth_b200.so (the way a C++ host would) and runs
it on the GPU: the th:: op surface, the two-phase pipeline idiom, validators, deferred command buffers and both
th_eval_gpu paths, from C++."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp_test():
    from token_hawk_b200 import build
    build.build()
    lib = os.path.join(ROOT, "token_hawk_b200", "lib")
    exe = os.path.join(lib, "th_api_test")
    src = os.path.join(ROOT, "tests", "cpp", "th_api_test.cpp")
    deps = [src, os.path.join(lib, "libth_b200.so")] + [os.path.join(ROOT, "include", "th", f) for f in os.listdir(os.path.join(ROOT, "include", "th"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):   # headers change the struct layouts
        subprocess.check_call([build.CXX, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L" + lib, "-lth_b200", "-lthk_sm100a", "-Wl,-rpath," + lib])
    return exe


def test_cpp_api_test_compiles():
    assert os.path.exists(build_cpp_test())


def test_cpp_host_surface_on_cpu():
    """tokenizer + sampler through include/th/th-llama.hpp from C++ (no device needed)."""
    from token_hawk_b200 import build
    build.build()
    lib = os.path.join(ROOT, "token_hawk_b200", "lib")
    exe = os.path.join(lib, "th_host_test")
    src = os.path.join(ROOT, "tests", "cpp", "th_host_test.cpp")
    subprocess.check_call([build.CXX, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L" + lib, "-lth_b200", "-lthk_sm100a", "-Wl,-rpath," + lib])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS th_host_test" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_api_on_gpu():
    exe = build_cpp_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS th_api_test" in r.stdout, r.stdout + r.stderr
