// ctx.cuh -- the context object behind vpc_ctx and the launch / error macros shared by the host-side sources of libvpc.
#pragma once

#include "../../include/vpc.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <new>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "dbscan.cuh"
#include "icp.cuh"
#include "host/staging.hpp"

using namespace vpc;

namespace {

struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  void reset() { off = 0; }
  template <class T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

}  // namespace

struct vpc_ctx {
  int device = 0;
  std::mutex mu;
  std::string err;
  int64_t launches = 0;
  cudaStream_t own_stream = nullptr;
  Arena db;        // DBSCAN workspace
  Arena io;        // device copies of host inputs / outputs (host-pointer entry points)
  Arena icp_model; // model cell list (persists between calls)
  Arena icp_work;  // per-call ICP workspace
  Arena st;        // sort / cluster statistics / ingest workspace
  Arena blk;       // blocked clustering / centroid merge buffers (blocked_api.cuh)
  // ICP model state
  IcpModel model{};
  bool model_set = false;
  IcpState* icp_state = nullptr;
  double* icp_partial = nullptr;
  unsigned* icp_ticket = nullptr;
  int icp_partial_blocks = 0;
  int sm_count = 148;
  DbArgs db_slab{};       // arguments of the last vpc_dbscan_slab_local_dev, for ..._finish_dev
  bool db_slab_valid = false;
  DbArgs db_pre{};        // workspace laid out by the slab step's phase 0 (pre-cut mode), consumed by its phase 1
  bool db_pre_valid = false;
  int64_t db_ws_n = -1;  // n the DBSCAN workspace is currently laid out and initialised for
  bool db_ws_banded = false;
  // optional per-kernel CUDA-event timing (bench.py's roofline leg)
  bool profile = false;
  bool pdl = true;        // programmatic dependent launch between the kernels of a chain (VPC_PDL=0: off)
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  // pageable caller memory: worker threads + a page-locked ring (host/staging.hpp); created on first use
  vpc_host::CopyPool* pool = nullptr;
  bool pool_tried = false;
  int copy_workers_forced = 0;   // VPC_COPY_THREADS: every copy uses this many workers (0 = chosen by size)
  vpc_host::Stager stager;
  // single-process multi-GPU mode (vpc_create with n_devices > 1): one sub-context per rank, see group_api.cuh
  struct vpc_group* group = nullptr;
};

namespace {

#define VPC_CUDA(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                       \
      return (_e == cudaErrorMemoryAllocation) ? VPC_E_NOMEM : VPC_E_CUDA;                   \
    }                                                                                        \
  } while (0)

// pdl: the kernel begins with pdl_enter() (common.cuh) and may be launched while its predecessor in the stream is still draining
// (programmatic dependent launch; the edge survives stream capture into a CUDA graph).  VPC_PDL=0 switches it off.
template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

#define VPC_LAUNCH_IMPL(ctx, use_pdl, kernel, grid, block, stream, ...)                          \
  do {                                                                                       \
    cudaEvent_t _ea = nullptr, _eb = nullptr;                                                \
    if ((ctx)->profile) {                                                                    \
      cudaEventCreate(&_ea); cudaEventCreate(&_eb); cudaEventRecord(_ea, (stream));          \
    }                                                                                        \
    launch_kernel(kernel, dim3(grid), dim3(block), (stream), (use_pdl) && (ctx)->pdl && !(ctx)->profile, __VA_ARGS__); \
    if ((ctx)->profile) {                                                                    \
      cudaEventRecord(_eb, (stream)); (ctx)->prof.push_back({#kernel, _ea, _eb});            \
    }                                                                                        \
    (ctx)->launches++;                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#kernel) + ": " + cudaGetErrorString(_e);                     \
      return VPC_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)
#define VPC_LAUNCH(ctx, kernel, grid, block, stream, ...) VPC_LAUNCH_IMPL(ctx, false, kernel, grid, block, stream, __VA_ARGS__)
#define VPC_LAUNCH_PDL(ctx, kernel, grid, block, stream, ...) VPC_LAUNCH_IMPL(ctx, true, kernel, grid, block, stream, __VA_ARGS__)

// worker threads for pageable host memory: VPC_COPY_THREADS (0 = plain cudaMemcpyAsync).  The pool holds up to eight; a copy engages
// two of them below 24 MiB and more above (host/staging.hpp).  With VPC_COPY_THREADS set, every copy may use all of them.
vpc_host::CopyPool* ctx_pool(vpc_ctx* ctx) {
  if (!ctx->pool_tried) {
    ctx->pool_tried = true;
    const unsigned hw = std::thread::hardware_concurrency();
    int t = hw >= 12 ? 8 : (hw >= 6 ? 4 : (hw >= 4 ? 2 : 1));
    if (const char* e = std::getenv("VPC_COPY_THREADS")) { t = std::atoi(e); ctx->copy_workers_forced = t; }
    if (t > 0) ctx->pool = new (std::nothrow) vpc_host::CopyPool(t);
    ctx->stager.forced_workers = ctx->copy_workers_forced;
  }
  return ctx->pool;
}

int fail(vpc_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

// Grow-only device arena.  Growing synchronises the device (old buffer may be in flight).
int arena_reserve(vpc_ctx* ctx, Arena& a, size_t bytes) {
  a.reset();
  if (bytes <= a.cap) return VPC_OK;
  VPC_CUDA(ctx, cudaDeviceSynchronize());
  if (a.base) VPC_CUDA(ctx, cudaFree(a.base));
  a.base = nullptr; a.cap = 0;
  size_t want = bytes + bytes / 8 + (1u << 20);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ctx->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    return VPC_E_NOMEM;
  }
  a.base = static_cast<char*>(p); a.cap = want;
  return VPC_OK;
}

}  // namespace
