"""Builds and runs the C++ host mirror's known-answer tests (include/dgx_algo.hpp over the
C ABI), which transcribe the Go tests of algo/uidlist_test.go."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp, name="test_algo_kat"):
    exe = os.path.join(tmp, name)
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
        "-L", os.path.join(ROOT, "dgraph_b200"), "-ldgx", "-Wl,-rpath," + os.path.join(ROOT, "dgraph_b200"),
    ])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    build(str(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_known_answers(tmp_path):
    exe = build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_wire_known_answers(tmp_path):
    """dgx::wire (host code): runs without a device."""
    exe = build(str(tmp_path), "test_wire_kat")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "WIRE_KAT_OK" in r.stdout, r.stdout + r.stderr
