"""A/B builds of libvpc.so for kernel experiments: python tools/ab_build.py name "<extra nvcc flags>" -> vtkcloudpoint_b200/ab/libvpc_<name>.so
(bench.py picks one up through VPC_LIB=<path>)."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import _build
name, extra = sys.argv[1], sys.argv[2].split()
out = _build.PKG / "ab"; out.mkdir(exist_ok=True)
cmd = [_build.nvcc_path(), *_build.NVCC_FLAGS, *extra, "-o", str(out / f"libvpc_{name}.so"), str(_build.CSRC / "vpc_api.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
print(name, "rc", r.returncode, r.stderr[-500:])
