"""vtkcloudpoint_b200 -- B200-native (sm_100a) DBSCAN + ICP hot path of ZhiHuangHn/vtkCloudPoint.

The product is the C-ABI CUDA library libvpc.so (include/vpc.h); this package holds its
sources (csrc/), the build recipe, the ctypes binding and the host-side mirror of the
reference's BaseClass interface (api.py; C++ mirror csrc/host/vpc_host.hpp; C# shim csharp/).  Nothing here imports oracle/.
"""
from .api import Context, DbscanResult, IcpResult  # noqa: F401
from .capi import VpcError  # noqa: F401

__all__ = ["Context", "DbscanResult", "IcpResult", "VpcError"]
