"""Builds and runs the C++ unit tests of the DFT codelets (host build of csrc/fft_codelets.cuh)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["test_codelets", "test_codelets32"])
def test_codelets_against_naive_dft(tmp_path, name):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path / name
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpu", name + ".cpp")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    print(res.stdout)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr
