"""Multi-GPU DBSCAN / ICP over peer memory (NVLink / NVSwitch): Python plumbing above the C ABI (include/vpc.h, "across GPUs
over peer memory").  The step itself is libvpc kernels only -- loads / stores to mapped peer addresses plus epoch flags
(csrc/comm.cuh, slab.cuh, icp_dist.cuh); torch.distributed is used ONCE, at set-up, to exchange the 64-byte heap handles
and (for the calibration) a few counters.  There is no CPU fallback.

The reference has no counterpart: it is one process with a thread pool over halo-less cells (FrmMain.cs:1356-1359).

  PeerComm            one rank's exchange heap, connected to everybody's (one process per GPU: cudaIpc handles through
                      torch.distributed; one process: PeerComm.local_group)
  SlabPeerPlan        exact DBSCAN of a pre-cut cloud, one u-slab per rank (vpc_slab_plan_*)
  calibrated_slab_plan  capacities measured by one probe step
  IcpDistPlan         ICP with the target (mode 0) or the source (mode 1) sharded (vpc_icp_dist_*)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi

_NP2TORCH = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8}


class _DevArray:
    """A raw device pointer dressed for torch.as_tensor (zero copy)."""

    def __init__(self, ptr: int, n: int, np_dtype):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": np.dtype(np_dtype).str, "data": (int(ptr), False), "version": 2}


def _wrap(ptr: int, n: int, np_dtype, device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, n, np_dtype), device=device)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class PeerComm:
    def __init__(self, ctx, rank: int, world: int, heap_bytes: int):
        self.ctx, self.rank, self.world, self.heap_bytes = ctx, int(rank), int(world), int(heap_bytes)
        self._lib = ctx._lib
        self._h = C.c_void_p()
        self._ipc = False
        ctx._check(self._lib.vpc_comm_create(ctx._h, self.rank, self.world, self.heap_bytes, C.byref(self._h)))

    @classmethod
    def connected(cls, ctx, heap_bytes: int, device, group=None) -> "PeerComm":
        """One process per GPU: create the heap, all_gather the handles, open everybody's heap."""
        rank, world = _world(group)
        comm = cls(ctx, rank, world, heap_bytes)
        if world > 1:
            buf = C.create_string_buffer(64)
            ctx._check(comm._lib.vpc_comm_handle(comm._h, buf))
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(device)
            everybody = torch.empty(world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(everybody, mine, group=group)
            raw = everybody.cpu().numpy().tobytes()
            ctx._check(comm._lib.vpc_comm_connect(comm._h, raw))
            comm._ipc = True
            dist.barrier(group=group)          # nobody frees or reuses a heap before everybody has opened it
        return comm

    @staticmethod
    def local_group(ctxs, heap_bytes: int) -> list["PeerComm"]:
        """One process: one comm per context (contexts may share a device -- emulation mode, drive the ranks phase by phase)."""
        comms = [PeerComm(c, r, len(ctxs), heap_bytes) for r, c in enumerate(ctxs)]
        arr = (C.c_void_p * len(comms))(*[c._h for c in comms])
        for c in comms:
            c.ctx._check(c._lib.vpc_comm_connect_local(c._h, arr))
        return comms

    def barrier_dev(self, device):
        """A barrier over the ranks as one tiny kernel on torch's current stream (no host synchronisation)."""
        self.ctx._check(self._lib.vpc_comm_barrier_dev(self._h, torch.cuda.current_stream(device).cuda_stream))

    def error_bits(self) -> int:
        v = C.c_int32(0)
        self.ctx._check(self._lib.vpc_comm_error(self._h, C.cast(C.byref(v), C.c_void_p)))
        return int(v.value)

    def close(self, group=None):
        """One process per GPU: every rank must call this (the ranks synchronise before anybody frees its heap)."""
        if self._h.value:
            if self.world > 1 and dist.is_available() and dist.is_initialized() and self._ipc:
                self._lib.vpc_comm_disconnect(self._h)
                dist.barrier(group=group)
            self._lib.vpc_comm_destroy(self._h)
            self._h = C.c_void_p()


class SlabPeerPlan:
    """vpc_slab_plan: rank r of `comm` holds slab r of ONE pre-cut cloud.  x / y are the plan's own input buffers (write the slab
    into them); cluster_id / is_key / is_classed / status are its outputs (include/vpc.h)."""

    def __init__(self, comm: PeerComm, n_per_rank, splitters, eps: float, min_pts: int, coord_bound: float, cap_halo: int, cap_pairs: int, device):
        self.comm, self.ctx, self._lib = comm, comm.ctx, comm._lib
        self.n = int(n_per_rank[comm.rank])
        self.cap_halo, self.cap_pairs = int(cap_halo), int(cap_pairs)
        npr = (C.c_int64 * comm.world)(*[int(v) for v in n_per_rank])
        spl = (C.c_double * max(comm.world - 1, 1))(*[float(v) for v in splitters])
        self._h = C.c_void_p()
        self.ctx._check(self._lib.vpc_slab_plan_create(self.ctx._h, comm._h, npr, spl, float(eps), int(min_pts), float(coord_bound), self.cap_halo,
                                                       self.cap_pairs, C.byref(self._h)))
        ptrs = [C.c_void_p() for _ in range(6)]
        self.ctx._check(self._lib.vpc_slab_plan_io(self._h, *[C.byref(p) for p in ptrs]))
        self.x = _wrap(ptrs[0].value, self.n, np.float64, device)
        self.y = _wrap(ptrs[1].value, self.n, np.float64, device)
        self.cluster_id = _wrap(ptrs[2].value, self.n, np.int32, device)
        self.is_key = _wrap(ptrs[3].value, self.n, np.uint8, device)
        self.is_classed = _wrap(ptrs[4].value, self.n, np.uint8, device)
        self.status = _wrap(ptrs[5].value, 16, np.int32, device)
        self.device = device

    @staticmethod
    def heap_bytes(lib, world: int, n_max: int, cap_halo: int, cap_pairs: int) -> int:
        return int(lib.vpc_slab_plan_heap_bytes(world, n_max, cap_halo, cap_pairs)) + 4096

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def step(self, first_cluster_id: int = 0):
        self.ctx._check(self._lib.vpc_slab_step_dev(self._h, int(first_cluster_id), self._stream()))
        return self.cluster_id, self.is_key, self.is_classed, self.status

    def step_phase(self, phase: int, first_cluster_id: int = 0):
        self.ctx._check(self._lib.vpc_slab_step_phase_dev(self._h, int(phase), int(first_cluster_id), self._stream()))

    def close(self):
        if self._h.value:
            self._lib.vpc_slab_plan_destroy(self._h)
            self._h = C.c_void_p()


def slab_heap_bytes(lib, world, n_max, cap_halo, cap_pairs):
    return SlabPeerPlan.heap_bytes(lib, world, n_max, cap_halo, cap_pairs)


def calibrated_slab_plan(ctx, x, y, n_per_rank, splitters, eps, min_pts, coord_bound, device, group=None, margin: float = 1.5):
    """(comm, plan) with MEASURED exchange capacities: one probe step with generous buffers, the largest halo strip / pair count over
    all ranks read back once (the only host synchronisation, at plan creation), then margin x that.  x, y: this rank's slab."""
    rank, world = _world(group)
    n_max = int(max(n_per_rank))
    cap_h, cap_p = max(1024, n_max // 4), max(1024, n_max // 2)
    comm = PeerComm.connected(ctx, slab_heap_bytes(ctx._lib, world, n_max, cap_h, cap_p), device, group)
    probe = SlabPeerPlan(comm, n_per_rank, splitters, eps, min_pts, coord_bound, cap_h, cap_p, device)
    probe.x.copy_(x); probe.y.copy_(y)
    probe.step()
    need = probe.status[1:4].to(torch.int64).clone()
    if world > 1:
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=group)
    err, halo, pairs = (int(v) for v in need.tolist())
    probe.close()
    comm.close(group)
    if err:
        raise ValueError(f"the probe step failed (error bits {err}: 1 = a rank timed out, 2 = overflow of generous buffers)")
    cap_h, cap_p = int(halo * margin) + 1024, int(pairs * margin) + 1024
    comm = PeerComm.connected(ctx, slab_heap_bytes(ctx._lib, world, n_max, cap_h, cap_p), device, group)
    plan = SlabPeerPlan(comm, n_per_rank, splitters, eps, min_pts, coord_bound, cap_h, cap_p, device)
    plan.x.copy_(x); plan.y.copy_(y)
    return comm, plan


class IcpDistPlan:
    """vpc_icp_dist: mode 0 = target sharded (this context's model is shard `rank`, first global index idx_offset), mode 1 = source
    sharded (whole target on every rank).  data: (3, n) float64 CUDA tensor with ALL data points, identical on every rank."""

    def __init__(self, comm: PeerComm, mode: int, data: torch.Tensor, idx_offset: int = 0):
        self.comm, self.ctx, self._lib = comm, comm.ctx, comm._lib
        self.data = data.contiguous()
        self.n, self.device, self.mode = int(data.shape[1]), data.device, int(mode)
        self._h = C.c_void_p()
        self.ctx._check(self._lib.vpc_icp_dist_create(self.ctx._h, comm._h, self.mode, self.data.data_ptr(), self.n, int(idx_offset), C.byref(self._h)))
        self.state = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.order = torch.zeros(self.n, dtype=torch.int32, device=self.device)

    @staticmethod
    def heap_bytes(lib, world: int, n: int) -> int:
        return int(lib.vpc_icp_dist_heap_bytes(world, n)) + 4096

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def begin(self):
        self.ctx._check(self._lib.vpc_icp_dist_begin_dev(self._h, self._stream()))

    def rounds(self, e: float, max_iters: int, rounds: int | None = None):
        """Enqueue `rounds` (default max_iters) rounds and the export; returns (state f64[16], order int32[n]) device tensors."""
        self.ctx._check(self._lib.vpc_icp_dist_rounds_dev(self._h, float(e), int(max_iters), int(max_iters if rounds is None else rounds),
                                                          self.state.data_ptr(), self.order.data_ptr(), self._stream()))
        return self.state, self.order

    def round_phase(self, phase: int, e: float, max_iters: int):
        self.ctx._check(self._lib.vpc_icp_dist_round_phase_dev(self._h, int(phase), float(e), int(max_iters), self._stream()))

    def export(self):
        return self.rounds(0.0, 1, 0)

    def run(self, e: float, max_iters: int):
        self.begin()
        return self.rounds(e, max_iters)

    def close(self):
        if self._h.value:
            self._lib.vpc_icp_dist_destroy(self._h)
            self._h = C.c_void_p()


class GraphedStep:
    """Any enqueue-only callable captured once into a CUDA graph and replayed (the slab step has ~20 kernels, the ICP loop 100-150:
    issued from Python one by one they are bound by the host's launch rate).  Every rank must construct and replay in lockstep."""

    def __init__(self, fn, device, warmup: int = 2):
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def replay(self):
        self.graph.replay()
        return self.out
