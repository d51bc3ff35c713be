"""World-size 1/2/3 gloo tests of the C5 pipeline's HOST logic (cluster-sharded statistics, all_to_all in rawData order,
result gather, sharded ICP) with checker-backed CPU backends.  Expected = the oracle pipeline on the whole cloud."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pipeline_ref
        from dist_cpu_backend import CpuPipelineBackend
        from vtkcloudpoint_b200.pipeline import run_pipeline
        mx, my, xyz = pipeline_ref.scene(0xC5, 9, 4200)
        pre = pipeline_ref.run(mx, my, xyz, np.zeros((2, 1)), 0.07, 7, 0.088, -1.0, 1)
        truth = pipeline_ref.truth_for(pre["means"][:2, pre["kept"]])
        ref = pipeline_ref.run(mx, my, xyz, truth, 0.07, 7, 0.088, 1e-9, 6, match_distance=0.05)
        assert ref["filtered"].sum() >= 2 and len(ref["kept"]) >= 60          # the scene exercises the filter
        n = len(mx)
        a, b = n * rank // world, n * (rank + 1) // world
        res = run_pipeline(CpuPipelineBackend(), torch.from_numpy(mx[a:b].copy()), torch.from_numpy(my[a:b].copy()),
                           torch.from_numpy(np.ascontiguousarray(xyz[:, a:b])), a, torch.from_numpy(truth), eps=0.07, min_pts=7,
                           radius_threshold=0.088, icp_e=1e-9, icp_max_iters=6, match_distance=0.05)
        pipeline_ref.check(res, ref, a, b)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_pipeline_matches_whole_cloud(world):
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)
