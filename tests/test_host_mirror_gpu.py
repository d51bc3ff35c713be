"""Builds and runs the C++ host-mirror test (tests/cpp/test_host_mirror.cpp): the reference's BaseClass usage
pattern (DBImproved with seeded cf, ICP.go_hell_ICP with zero matrices) on top of the C ABI, checked against the oracle."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_cpp_host_mirror(ctx, oracle, tmp_path):
    from vtkcloudpoint_b200 import _build
    exe = tmp_path / "test_host_mirror"
    cmd = ["g++", "-O1", "-std=c++17", str(ROOT / "tests/cpp/test_host_mirror.cpp"), "-o", str(exe),
           str(_build.LIB), str(oracle.LIB), f"-Wl,-rpath,{_build.LIB.parent}", f"-Wl,-rpath,{oracle.LIB.parent}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror ok" in out.stdout
