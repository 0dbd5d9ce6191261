"""Builds and runs the native C++ unit tests (tests/cpp/test_native.cc) against lib/libdglb200_cpu.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_cpp_suite(tmp_path):
    from dgl_b200._native import loader
    loader.cpu_lib()  # makes sure the library is built
    libdir = os.path.join(ROOT, "dgl_b200", "lib")
    exe = str(tmp_path / "test_native")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_native.cc"), "-o", exe,
                    "-L" + libdir, "-ldglb200_cpu", "-Wl,-rpath," + libdir, "-pthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
