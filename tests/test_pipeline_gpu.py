"""GPU parity of the C5 pipeline on one GPU (DBSCAN -> centroids + circles -> radius filter -> ICP -> match) against the
oracle pipeline on the same scene: labels, counts, kept ids, correspondences exact; centroids and circles bit-exact; R/T/SSE 1e-6."""
import numpy as np
import pytest
import torch

import pipeline_ref
from vtkcloudpoint_b200.pipeline import GpuPipelineBackend, run_pipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid,n_total", [(9, 4200), (45, 104_000)])
def test_pipeline_single_gpu(ctx, grid, n_total):
    mx, my, xyz = pipeline_ref.scene(0xC5, grid, n_total)
    pre = pipeline_ref.run(mx, my, xyz, np.zeros((2, 1)), 0.07, 7, 0.088, -1.0, 1)
    truth = pipeline_ref.truth_for(pre["means"][:2, pre["kept"]])
    ref = pipeline_ref.run(mx, my, xyz, truth, 0.07, 7, 0.088, 1e-9, 8, match_distance=0.05)
    assert ref["filtered"].sum() >= 2
    dev = torch.device("cuda", 0)
    res = run_pipeline(GpuPipelineBackend(ctx), torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev), torch.from_numpy(xyz).to(dev), 0,
                       torch.from_numpy(truth).to(dev), eps=0.07, min_pts=7, radius_threshold=0.088, icp_e=1e-9, icp_max_iters=8, match_distance=0.05)
    pipeline_ref.check(res, ref, 0, len(mx))
    # every kept centre finds its own truth point again (the scene is a rigid motion + small noise)
    assert (res.matched.cpu().numpy() >= 0).mean() > 0.99
