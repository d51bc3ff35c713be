// comm.cuh -- peer-memory exchange between the GPUs of one NVLink / NVSwitch box, without NCCL calls on the step.
//
// Every rank (= one GPU) owns a "heap": one device allocation of the same size and layout on every rank.  A kernel on rank r
// reaches rank q's heap through a peer pointer (cudaIpcOpenMemHandle across processes, cudaDeviceEnablePeerAccess inside one
// process), so an exchange is ordinary loads / stores to mapped peer addresses over NVLink plus a flag:
//   producer:  data stores ... (last block) ONE __threadfence_system(), then st.relaxed.sys flag[q][phase][me] = epoch << 32 | count
//   consumer:  spin ld.relaxed.sys on flag[me][phase][src] until its epoch arrives, one __threadfence_system(), then read (local data,
//              or the producer's heap by peer loads)
// Flags carry a monotonically increasing epoch, so they are never reset and a CUDA-graph replay needs no host work.  Every spin is
// bounded (kSpinLimit cycles): a rank that never shows up raises the heap's error word instead of hanging the GPU.
// Peer loads use ld.relaxed.sys (peer lines are cached in L1 only, SURVEY/B300_MICROARCH: L2 is bypassed), so a value written by
// the owner in step k is never served stale in step k+1.
//
// The reference has no counterpart (single process, thread pool, FrmMain.cs:1356-1359); this is the exchange layer of SURVEY.md 8e.
#pragma once

#include "common.cuh"

namespace vpc {

constexpr int kMaxWorld = 16;
constexpr int kPhases = 8;                       // flag rows per heap
constexpr long long kSpinLimit = 6000000000ll;   // ~3 s at 1.9 GHz

// phases (rows of the flag table)
constexpr int kPhHalo = 0, kPhPairs = 1, kPhHeads = 2, kPhIcpNn = 3, kPhIcpSums = 4, kPhGather = 5, kPhHome = 6, kPhBarrier = 7;

struct HeapHeader {                              // offset 0 of every heap
  unsigned long long flag[kPhases][kMaxWorld];   // flag[phase][src]: written by rank src, read by the heap's owner
  unsigned long long reserved[kPhases][kMaxWorld];
  int error;                                     // != 0: a bounded spin gave up (bit 0) / an exchange buffer overflowed (bit 1)
  int pad[15];
  unsigned long long epoch[8];                   // step counters of the heap's OWNER ([0] slab DBSCAN, [1] ICP): they live with the flags, so
                                                 // every plan ever created on this heap continues the same monotone sequence
};
constexpr size_t kHeapHeaderBytes = (sizeof(HeapHeader) + 255) & ~size_t(255);

struct Peers {                                   // passed to kernels by value
  char* base[kMaxWorld];                         // base[q] = rank q's heap as seen from THIS rank's address space
  int rank, world;
  __device__ __forceinline__ HeapHeader* hdr(int q) const { return reinterpret_cast<HeapHeader*>(base[q]); }
  template <class T>
  __device__ __forceinline__ T* at(int q, size_t off) const { return reinterpret_cast<T*>(base[q] + off); }
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_sys_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// A flag word = (epoch << 32) | payload: ONE relaxed system-scope store per destination publishes both, so a count travels with the
// flag without a second ordered store.  The CALLER orders its data before the flags with ONE __threadfence_system() (fence + relaxed
// store = a release pattern); a st.release.sys per destination would repeat that fence `world` times -- on 4 x B200 that alone was
// ~60 us per ICP round (profiles/r02_multi_gpu.md).
__device__ __forceinline__ void comm_signal(const Peers& P, int dst, int phase, unsigned long long epoch, unsigned long long payload) {
  st_relaxed_sys_u64(&P.hdr(dst)->flag[phase][P.rank], (epoch << 32) | (payload & 0xffffffffull));
}
// Wait until rank `src` has signalled `phase` of step `epoch` (bounded): relaxed polling, one acquiring fence at the end.
// Returns false on timeout (and raises the error word).
__device__ __forceinline__ bool comm_wait(const Peers& P, int src, int phase, unsigned long long epoch) {
  HeapHeader* h = P.hdr(P.rank);
  const unsigned long long* f = &h->flag[phase][src];
  bool ok = true;
  if ((ld_relaxed_sys_u64(f) >> 32) < epoch) {
    const long long t0 = clock64();
    while ((ld_relaxed_sys_u64(f) >> 32) < epoch) {
      if (clock64() - t0 > kSpinLimit) { atomicOr(&h->error, 1); ok = false; break; }
      __nanosleep(32);
    }
  }
  __threadfence_system();
  return ok;
}
__device__ __forceinline__ unsigned long long comm_payload(const Peers& P, int src, int phase) {
  return ld_relaxed_sys_u64(&P.hdr(P.rank)->flag[phase][src]) & 0xffffffffull;
}

// One thread per source rank waits, then the block is released.  Call from ALL threads of a block.
__device__ __forceinline__ void comm_wait_all_block(const Peers& P, int phase, unsigned long long epoch) {
  if ((int)threadIdx.x < P.world) comm_wait(P, (int)threadIdx.x, phase, epoch);
  __syncthreads();
}

// A barrier over the ranks as one tiny kernel: signal everybody, then wait for everybody (its own epoch counter, epoch[2]).  Only for
// ranks that really run at the same time (one process per GPU / one stream per device): it signals before it waits, so it must not be
// used in the phase-by-phase emulation of several ranks on one GPU.
__global__ void k_comm_barrier(Peers P) {
  __shared__ unsigned long long s_e;
  if (threadIdx.x == 0) {
    unsigned long long* e = &P.hdr(P.rank)->epoch[2];
    s_e = *e + 1; *e = s_e;
    __threadfence_system();
  }
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    comm_signal(P, (int)threadIdx.x, kPhBarrier, s_e, 0ull);
    comm_wait(P, (int)threadIdx.x, kPhBarrier, s_e);
  }
}

}  // namespace vpc
