// staging.hpp -- host <-> device copies for PAGEABLE caller memory (what a P/Invoke marshaller hands over: a GC-pinned but not
// page-locked double[]).  The driver's own pageable path is a single-threaded bounce; here a few worker threads (and the calling
// thread) copy chunks into / out of a page-locked ring IN PARALLEL and each issues the DMA of its own chunk, so the host-side copy
// runs at several cores' memcpy rate and overlaps with the PCIe transfers.  Page-locked caller memory (vpc_host_alloc,
// vpc_host_register) is detected and copied directly.
//
// The reference has no counterpart: its points live in List<Point3D> objects (DataModel.cs:102-160); the shim flattens them
// into arrays for the call (csharp/DBImprovedGpu.cs).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vpc_host {

// parallel_for over a handful of persistent threads.  Workers spin briefly for the next job (a call makes several in a row), then
// sleep; one job at a time, the caller takes part.  A job names how many workers may join (small copies are fastest with two,
// measured on the 16-core B200 host: 8 MB copies 1 / 2 / 4 / 8 workers -> 0.77 / 1.00 / 0.98 / 0.95 Gpts/s end to end; 64 MB copies
// 2 / 4 / 8 workers -> 0.93 / 1.24 / 1.24 Gpts/s on one GPU, 0.89 / 1.09 / 1.29 on two); the others go back to sleep at once.
class CopyPool {
 public:
  explicit CopyPool(int n_threads) {
    for (int t = 0; t < n_threads; ++t) th_.emplace_back([this, t] { run(t); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_.fetch_add(1); }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return (int)th_.size(); }

  void parallel_for(size_t n, const std::function<void(size_t)>& body, int max_workers = 1 << 30) {
    if (n == 0) return;
    if (n == 1 || th_.empty() || max_workers <= 0) { for (size_t i = 0; i < n; ++i) body(i); return; }
    {
      std::lock_guard<std::mutex> lk(m_);
      body_ = &body; n_ = n; next_.store(0); done_.store(0); open_ = true; limit_ = max_workers;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    work(body, n);
    while (done_.load(std::memory_order_acquire) < n) std::this_thread::yield();
    { std::lock_guard<std::mutex> lk(m_); open_ = false; }
    while (inside_.load(std::memory_order_acquire) != 0) std::this_thread::yield();   // nobody still holds a pointer to `body`
  }

 private:
  void work(const std::function<void(size_t)>& body, size_t n) {
    size_t i, mine = 0;
    while ((i = next_.fetch_add(1, std::memory_order_relaxed)) < n) { body(i); ++mine; }
    if (mine) done_.fetch_add(mine, std::memory_order_acq_rel);
  }
  void run(int id) {
    unsigned long long seen = 0;
    bool sat_out = false;
    for (;;) {
      // wait for a new generation: spin for a short while (the next copy of the same call is microseconds away), then sleep
      int spins = sat_out ? 4000 : 0;
      sat_out = false;
      while (gen_.load(std::memory_order_acquire) == seen) {
        if (++spins < 4000) { std::this_thread::yield(); continue; }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
      }
      const std::function<void(size_t)>* body = nullptr;
      size_t n = 0;
      {
        std::lock_guard<std::mutex> lk(m_);
        seen = gen_.load(std::memory_order_acquire);
        if (stop_) return;
        if (!open_) continue;
        if (id >= limit_) { sat_out = true; continue; }
        body = body_; n = n_;
        inside_.fetch_add(1, std::memory_order_acq_rel);
      }
      work(*body, n);
      inside_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> gen_{0};
  std::atomic<size_t> next_{0}, done_{0};
  std::atomic<int> inside_{0};
  const std::function<void(size_t)>* body_ = nullptr;
  size_t n_ = 0;
  int limit_ = 1 << 30;
  bool open_ = false, stop_ = false;
};

// One page-locked ring per device stream, used in two halves so that filling one half overlaps the DMA out of the other.
// Not thread safe: the owning context serialises its calls.
class Stager {
 public:
  static constexpr size_t kChunk = 512u << 10;

  Stager() = default;
  Stager(const Stager&) = delete;
  ~Stager() { release(); }

  void release() {
    for (auto& e : half_ev_) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : chunk_ev_) if (e) cudaEventDestroy(e);
    chunk_ev_.clear();
    if (ring_) cudaFreeHost(ring_);
    ring_ = nullptr; cap_ = 0;
  }

  // page-locked (cudaHostAlloc / cudaHostRegister) memory goes straight to the DMA engine
  static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
  }

  // the ring holds two halves of min(bytes, 64 MiB) each; grows only
  cudaError_t reserve(size_t bytes) {
    // (+ 4 chunks: the arrays of a multi-array job each start on a chunk boundary)
    const size_t half = ((std::min<size_t>(std::max<size_t>(bytes, kChunk), 64u << 20) + kChunk - 1) / kChunk + 4) * kChunk;
    if (2 * half <= cap_) return cudaSuccess;
    release();
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ring_), 2 * half, cudaHostAllocPortable);
    if (e != cudaSuccess) { ring_ = nullptr; return e; }
    cap_ = 2 * half;
    chunk_ev_.assign(cap_ / kChunk, nullptr);
    half_busy_[0] = half_busy_[1] = false;
    turn_ = 0;
    return cudaSuccess;
  }

  // host -> device on `stream` of `device`; on return every chunk has been handed to the DMA engine (the host range may be reused)
  // workers: how many pool threads may join (0 = by size: two below 24 MiB, four above)
  static int auto_workers(size_t bytes) { return bytes < (24u << 20) ? 2 : 4; }

  cudaError_t h2d(CopyPool* pool, void* dev, const void* host, size_t bytes, cudaStream_t stream, int device = -1, int workers = 0) {
    if (bytes == 0) return cudaSuccess;
    if (forced_workers > 0) workers = forced_workers; else if (workers <= 0) workers = auto_workers(bytes);
    if (!pool || is_pinned(host)) return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream);
    cudaError_t e = cap_ ? cudaSuccess : reserve(bytes);
    if (e != cudaSuccess) return e;
    if (device < 0) cudaGetDevice(&device);
    const size_t half = cap_ / 2;
    std::atomic<int> err{0};
    for (size_t off = 0; off < bytes; off += half) {
      const size_t piece = std::min(half, bytes - off), n_chunks = (piece + kChunk - 1) / kChunk;
      const int h = next_half();
      if ((e = wait_half(h)) != cudaSuccess) return e;
      char* base = ring_ + (size_t)h * half;
      const char* src = static_cast<const char*>(host) + off;
      char* dst = static_cast<char*>(dev) + off;
      pool->parallel_for(n_chunks, [&](size_t i) {
        const size_t o = i * kChunk, len = std::min(kChunk, piece - o);
        std::memcpy(base + o, src + o, len);
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        if (cudaMemcpyAsync(dst + o, base + o, len, cudaMemcpyHostToDevice, stream) != cudaSuccess) err.store(1);
      }, workers);
      if (err.load()) return cudaErrorUnknown;
      if ((e = mark_half(h, stream)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }

  // device -> host on `stream`; synchronous with respect to the caller's memory: on return `host` holds the data
  cudaError_t d2h(CopyPool* pool, void* host, const void* dev, size_t bytes, cudaStream_t stream, int device = -1, int workers = 0) {
    if (bytes == 0) return cudaSuccess;
    if (forced_workers > 0) workers = forced_workers; else if (workers <= 0) workers = auto_workers(bytes);
    if (!pool || is_pinned(host)) return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cap_ ? cudaSuccess : reserve(bytes);
    if (e != cudaSuccess) return e;
    if (device < 0) cudaGetDevice(&device);
    const size_t half = cap_ / 2;
    std::atomic<int> err{0};
    for (size_t off = 0; off < bytes; off += half) {
      const size_t piece = std::min(half, bytes - off), n_chunks = (piece + kChunk - 1) / kChunk;
      const int h = next_half();
      if ((e = wait_half(h)) != cudaSuccess) return e;
      char* base = ring_ + (size_t)h * half;
      const size_t ev0 = (size_t)h * (half / kChunk);
      for (size_t i = 0; i < n_chunks; ++i) {            // all DMAs first (they queue behind the kernels), one event per chunk
        const size_t o = i * kChunk, len = std::min(kChunk, piece - o);
        if ((e = cudaMemcpyAsync(base + o, static_cast<const char*>(dev) + off + o, len, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        cudaEvent_t& ev = chunk_ev_[ev0 + i];
        if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev, stream)) != cudaSuccess) return e;
      }
      char* dst = static_cast<char*>(host) + off;
      pool->parallel_for(n_chunks, [&](size_t i) {       // as each chunk lands a thread copies it into the caller's memory
        const size_t o = i * kChunk, len = std::min(kChunk, piece - o);
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        if (cudaEventSynchronize(chunk_ev_[ev0 + i]) != cudaSuccess) { err.store(1); return; }
        std::memcpy(dst + o, base + o, len);
      }, workers);
      if (err.load()) return cudaErrorUnknown;
      half_busy_[h] = false;                             // drained
    }
    return cudaSuccess;
  }

  // Several arrays in ONE job (one pool wake-up, all DMAs queued before the first wait): the results of a call -- ids, core flags,
  // isClassed -- come back as one stream of chunks instead of three copies that each drain before the next starts.
  struct Seg { void* host; void* dev; size_t bytes; };
  // VPC_STAGE_MULTI (diagnostics): bit 0 = host->device jobs, bit 1 = device->host jobs.  Default 2: measured on one box at 1M points,
  // mask 0 / 1 / 2 / 3 -> 0.98 / 0.98 / 0.91 / 0.93 ms per call: the results gain from one job, the inputs do not.
  static int multi_mask() {
    static const int m = [] { const char* e = std::getenv("VPC_STAGE_MULTI"); return e ? std::atoi(e) : 2; }();
    return m;
  }

  cudaError_t h2d_multi(CopyPool* pool, const Seg* segs, int n_seg, cudaStream_t stream, int device = -1, int workers = 0) {
    size_t total = 0, chunks = 0;
    bool pinned = false;
    for (int k = 0; k < n_seg; ++k) { total += segs[k].bytes; chunks += (segs[k].bytes + kChunk - 1) / kChunk; pinned = pinned || (segs[k].bytes && is_pinned(segs[k].host)); }
    cudaError_t e = (pool && !pinned && !cap_) ? reserve(total) : cudaSuccess;
    if (e != cudaSuccess) return e;
    if (!pool || pinned || chunks * kChunk > cap_ / 2 || n_seg > 8 || !(multi_mask() & 1)) {   // not one ring half: one array after the other
      for (int k = 0; k < n_seg; ++k) if ((e = h2d(pool, segs[k].dev, segs[k].host, segs[k].bytes, stream, device, workers)) != cudaSuccess) return e;
      return cudaSuccess;
    }
    if (chunks == 0) return cudaSuccess;
    if (forced_workers > 0) workers = forced_workers; else if (workers <= 0) workers = auto_workers(total);
    if (device < 0) cudaGetDevice(&device);
    const int h = next_half();
    if ((e = wait_half(h)) != cudaSuccess) return e;
    char* base = ring_ + (size_t)h * (cap_ / 2);
    size_t first[8];                                           // first chunk of every segment
    { size_t c = 0; for (int k = 0; k < n_seg; ++k) { first[k] = c; c += (segs[k].bytes + kChunk - 1) / kChunk; } }
    std::atomic<int> err{0};
    pool->parallel_for(chunks, [&](size_t i) {
      int k = n_seg - 1;
      while (k > 0 && first[k] > i) --k;
      const size_t o = (i - first[k]) * kChunk, len = std::min(kChunk, segs[k].bytes - o);
      std::memcpy(base + i * kChunk, static_cast<const char*>(segs[k].host) + o, len);
      int cur = -1;
      cudaGetDevice(&cur);
      if (cur != device) cudaSetDevice(device);
      if (cudaMemcpyAsync(static_cast<char*>(segs[k].dev) + o, base + i * kChunk, len, cudaMemcpyHostToDevice, stream) != cudaSuccess) err.store(1);
    }, workers);
    if (err.load()) return cudaErrorUnknown;
    return mark_half(h, stream);
  }

  cudaError_t d2h_multi(CopyPool* pool, const Seg* segs, int n_seg, cudaStream_t stream, int device = -1, int workers = 0) {
    size_t total = 0, chunks = 0;
    bool pinned = false;
    for (int k = 0; k < n_seg; ++k) { total += segs[k].bytes; chunks += (segs[k].bytes + kChunk - 1) / kChunk; pinned = pinned || (segs[k].bytes && is_pinned(segs[k].host)); }
    cudaError_t e = (pool && !pinned && !cap_) ? reserve(total) : cudaSuccess;
    if (e != cudaSuccess) return e;
    if (!pool || pinned || chunks * kChunk > cap_ / 2 || n_seg > 8 || !(multi_mask() & 2)) {
      for (int k = 0; k < n_seg; ++k) if ((e = d2h(pool, segs[k].host, segs[k].dev, segs[k].bytes, stream, device, workers)) != cudaSuccess) return e;
      return cudaSuccess;
    }
    if (chunks == 0) return cudaSuccess;
    if (forced_workers > 0) workers = forced_workers; else if (workers <= 0) workers = auto_workers(total);
    if (device < 0) cudaGetDevice(&device);
    const int h = next_half();
    if ((e = wait_half(h)) != cudaSuccess) return e;
    char* base = ring_ + (size_t)h * (cap_ / 2);
    const size_t ev0 = (size_t)h * ((cap_ / 2) / kChunk);
    size_t first[8];
    { size_t c = 0; for (int k = 0; k < n_seg; ++k) { first[k] = c; c += (segs[k].bytes + kChunk - 1) / kChunk; } }
    for (int k = 0; k < n_seg; ++k)                            // all DMAs first (they queue behind the kernels), one event per chunk
      for (size_t o = 0, i = first[k]; o < segs[k].bytes; o += kChunk, ++i) {
        const size_t len = std::min(kChunk, segs[k].bytes - o);
        if ((e = cudaMemcpyAsync(base + i * kChunk, static_cast<const char*>(segs[k].dev) + o, len, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        cudaEvent_t& ev = chunk_ev_[ev0 + i];
        if (!ev && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev, stream)) != cudaSuccess) return e;
      }
    std::atomic<int> err{0};
    pool->parallel_for(chunks, [&](size_t i) {                 // as each chunk lands a thread copies it into the caller's memory
      int k = n_seg - 1;
      while (k > 0 && first[k] > i) --k;
      const size_t o = (i - first[k]) * kChunk, len = std::min(kChunk, segs[k].bytes - o);
      int cur = -1;
      cudaGetDevice(&cur);
      if (cur != device) cudaSetDevice(device);
      if (cudaEventSynchronize(chunk_ev_[ev0 + i]) != cudaSuccess) { err.store(1); return; }
      std::memcpy(static_cast<char*>(segs[k].host) + o, base + i * kChunk, len);
    }, workers);
    if (err.load()) return cudaErrorUnknown;
    half_busy_[h] = false;                                     // drained
    return cudaSuccess;
  }

  cudaError_t finish(CopyPool*) { return cudaSuccess; }  // d2h is complete on return

  int forced_workers = 0;        // VPC_COPY_THREADS: every copy engages this many workers

 private:
  int next_half() { const int h = turn_; turn_ ^= 1; return h; }
  cudaError_t wait_half(int h) {                         // the DMA that last read this half has completed
    if (!half_busy_[h]) return cudaSuccess;
    half_busy_[h] = false;
    return cudaEventSynchronize(half_ev_[h]);
  }
  cudaError_t mark_half(int h, cudaStream_t stream) {
    cudaError_t e;
    if (!half_ev_[h] && (e = cudaEventCreateWithFlags(&half_ev_[h], cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(half_ev_[h], stream)) != cudaSuccess) return e;
    half_busy_[h] = true;
    return cudaSuccess;
  }

  char* ring_ = nullptr;
  size_t cap_ = 0;
  int turn_ = 0;
  cudaEvent_t half_ev_[2] = {nullptr, nullptr};
  bool half_busy_[2] = {false, false};
  std::vector<cudaEvent_t> chunk_ev_;
};

}  // namespace vpc_host
