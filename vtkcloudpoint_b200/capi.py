"""ctypes binding of the C ABI declared in include/vpc.h.

This is the same boundary the reference's C# would P/Invoke.  Loading
fails loudly when the CUDA library is missing or cannot be built: there is no CPU
fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

from . import _build

VPC_OK = 0
ERRORS = {-1: "VPC_E_BADARG", -2: "VPC_E_CUDA", -3: "VPC_E_NOMEM", -4: "VPC_E_NODEVICE",
          -5: "VPC_E_TOOBIG", -6: "VPC_E_STATE"}

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f64 = C.c_double

# name -> (restype, argtypes); every symbol include/vpc.h declares
SIGNATURES = {
    "vpc_create": (C.c_int, [C.POINTER(_p), C.POINTER(C.c_int), C.c_int]),
    "vpc_destroy": (None, [_p]),
    "vpc_last_error": (C.c_char_p, [_p]),
    "vpc_version": (C.c_char_p, []),
    "vpc_launch_count": (_i64, [_p]),
    "vpc_profile_enable": (C.c_int, [_p, C.c_int]),
    "vpc_profile_report": (_i64, [_p, C.c_char_p, _i64]),
    "vpc_host_alloc": (C.c_int, [C.POINTER(_p), _i64]),
    "vpc_host_free": (None, [_p]),
    "vpc_host_register": (C.c_int, [_p, _i64]),
    "vpc_host_unregister": (C.c_int, [_p]),
    "vpc_dbscan_l1_2d": (C.c_int, [_p, _p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p]),
    "vpc_dbscan_l1_2d_dev": (C.c_int, [_p, _p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p, _p]),
    "vpc_dbscan_l1_2d_cells": (C.c_int, [_p, _p, _p, _i64, _p, _i32, _f64, _i32, _p, _p, _p, _p]),
    "vpc_dbscan_l1_2d_cells_dev": (C.c_int, [_p, _p, _p, _i64, _p, _i32, _f64, _i32, _p, _p, _p, _p, _p]),
    "vpc_dbscan_slab_local_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f64, _i32, _p, _p, _p]),
    "vpc_dbscan_slab_finish_dev": (C.c_int, [_p, _p, _p, _i64, _p, _p]),
    "vpc_uf_edges_dev": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p]),
    "vpc_slab_halo_pack_dev": (C.c_int, [_p, _p, _p, _i64, _i32, _f64, _f64, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "vpc_slab_assemble_dev": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _i32, _p, _p, _p, _p]),
    "vpc_slab_pairs_dev": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _f64, _f64, _f64, _i32, _i32, _i32, _p, _p, _p]),
    "vpc_slab_heads_dev": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _p, _p]),
    "vpc_slab_ids_dev": (C.c_int, [_p, _p, _p, _i64, _p, _i64, _i32, _p, _p, _p, _p]),
    "vpc_slab_pairs_ws_dev": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _f64, _f64, _f64, _i32, _i32, _i32, _p, _p, _p]),
    "vpc_slab_merge_table_bytes": (_i64, [_i32, _i32]),
    "vpc_dbscan_takes_banded_path": (C.c_int, [_i64]),
    "vpc_dbscan_slab_finish_merge_dev": (C.c_int, [_p, _p, _i32, _i32, _p, _i64, _p, _p]),
    "vpc_closest_point_set": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _p]),
    "vpc_icp_rigid": (C.c_int, [_p, _p, _i64, _p, _i64, _f64, _i32, _p, _p, _p, _p, _p]),
    "vpc_icp_set_model_dev": (C.c_int, [_p, _p, _i64, _p]),
    "vpc_closest_point_set_dev": (C.c_int, [_p, _p, _i64, _p, _p, _p]),
    "vpc_icp_rigid_dev": (C.c_int, [_p, _p, _i64, _f64, _i32, _p, _p, _p]),
    "vpc_match_within": (C.c_int, [_p, _p, _i64, _p, _i64, _f64, _p, _p]),
    "vpc_match_within_dev": (C.c_int, [_p, _p, _i64, _f64, _p, _p, _p]),
    "vpc_cluster_means_dev": (C.c_int, [_p, _p, _i64, _i32, _p, _i32, _p, _p, _p]),
    "vpc_dbscan_blocked_ref": (C.c_int, [_p, _p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "vpc_dbscan_blocked_ref_ex": (C.c_int, [_p, _p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vpc_merge_ids_by_distance": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _f64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vpc_sort_pairs_dev": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "vpc_argsort_f64_dev": (C.c_int, [_p, _p, _i64, _p, _p]),
    "vpc_cluster_groups_dev": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p]),
    "vpc_cluster_means_ordered_dev": (C.c_int, [_p, _p, _p, _i32, _p, _i64, _i32, _p, _p, _p]),
    "vpc_cluster_circles_dev": (C.c_int, [_p, _p, _p, _i32, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "vpc_radius_filter_dev": (C.c_int, [_p, _p, _p, _i32, _f64, _p, _p]),
    "vpc_cluster_stats": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vpc_nearest_truth_2d": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p, _i64, _f64, _p]),
    "vpc_nearest_truth_2d_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f64, _p, _p, _p, _p]),
    "vpc_polar_to_xyz_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f64, _f64, _i32, _i32, _p, _p, _p]),
    "vpc_dedupe_xyz_dev": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _p]),
    "vpc_ingest_text": (C.c_int, [_p, _p, _i64, _f64, _f64, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vpc_synth_dbscan_cloud_dev": (C.c_int, [_p, C.c_uint64, _i32, _i32, _i64, _f64, _f64, _f64, _f64, _i64, _i64, _p, _p, _p]),
    "vpc_icp_shard_begin_dev": (C.c_int, [_p, _i64, _p]),
    "vpc_icp_shard_nn_dev": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p]),
    "vpc_icp_shard_select_dev": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "vpc_icp_shard_accumulate_dev": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _p]),
    "vpc_icp_shard_solve_dev": (C.c_int, [_p, _p, _i64, _f64, _i32, _p, _p]),
    "vpc_comm_create": (C.c_int, [_p, _i32, _i32, _i64, C.POINTER(_p)]),
    "vpc_comm_handle": (C.c_int, [_p, _p]),
    "vpc_comm_connect": (C.c_int, [_p, _p]),
    "vpc_comm_connect_local": (C.c_int, [_p, C.POINTER(_p)]),
    "vpc_comm_error": (C.c_int, [_p, _p]),
    "vpc_comm_barrier_dev": (C.c_int, [_p, _p]),
    "vpc_comm_disconnect": (C.c_int, [_p]),
    "vpc_comm_destroy": (None, [_p]),
    "vpc_slab_plan_heap_bytes": (_i64, [_i32, _i64, _i32, _i32]),
    "vpc_slab_plan_create": (C.c_int, [_p, _p, _p, _p, _f64, _i32, _f64, _i32, _i32, C.POINTER(_p)]),
    "vpc_slab_plan_io": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "vpc_slab_step_dev": (C.c_int, [_p, _i32, _p]),
    "vpc_slab_step_phase_dev": (C.c_int, [_p, _i32, _i32, _p]),
    "vpc_slab_plan_destroy": (None, [_p]),
    "vpc_icp_dist_heap_bytes": (_i64, [_i32, _i64]),
    "vpc_icp_dist_create": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, C.POINTER(_p)]),
    "vpc_icp_dist_begin_dev": (C.c_int, [_p, _p]),
    "vpc_icp_dist_round_phase_dev": (C.c_int, [_p, _i32, _f64, _i32, _p]),
    "vpc_icp_dist_rounds_dev": (C.c_int, [_p, _f64, _i32, _i32, _p, _p, _p]),
    "vpc_icp_dist_destroy": (None, [_p]),
}


class VpcError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


@lru_cache(maxsize=None)
def lib() -> C.CDLL:
    import os
    alt = os.environ.get("VPC_LIB")    # developer knob: an A/B build of the same sources (tools/ab_build.py)
    path = alt if alt else _build.build_lib()          # raises if nvcc is missing or the build fails
    dll = C.CDLL(str(path))            # raises OSError if the library cannot be loaded
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(dll, name)        # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return dll
