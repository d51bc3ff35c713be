"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vpc.h declares.
No compute calls here -- without a GPU vpc_create must fail loudly (there is no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "vpc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from vtkcloudpoint_b200 import capi
    dll = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in include/vpc.h but not exported by libvpc.so"
    assert sorted(capi.SIGNATURES) == declared, "capi.SIGNATURES must list exactly the symbols of include/vpc.h"
    assert b"sm_100a" in dll.vpc_version()


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    from vtkcloudpoint_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(_build.LIB)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback():
    import torch
    from vtkcloudpoint_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu suite")
    h = C.c_void_p()
    rc = capi.lib().vpc_create(C.byref(h), None, 0)
    assert rc == -4 and not h.value                 # VPC_E_NODEVICE
    from vtkcloudpoint_b200 import Context, VpcError
    with pytest.raises(VpcError):
        Context(0)


def test_product_never_imports_oracle():
    pkg = ROOT / "vtkcloudpoint_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        s = f.read_text()
        assert "oracle_py" not in s and "vpc_oracle" not in s and "vpco_" not in s, f


def test_copy_pool_runs_every_index_once(tmp_path):
    # csrc/host/staging.hpp: the worker pool behind the pageable-memory staging (pure host code; the Stager itself needs a GPU)
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    cuda_inc = Path(shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc").resolve().parent.parent / "include"
    exe = tmp_path / "test_copy_pool"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", f"-I{cuda_inc}", "-o", str(exe), str(root / "tests" / "cpp" / "test_copy_pool.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "copy pool ok" in out.stdout, out.stdout + out.stderr
