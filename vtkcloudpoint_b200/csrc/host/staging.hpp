// staging.hpp -- host <-> device copies for PAGEABLE caller memory (what a P/Invoke marshaller hands over: a GC-pinned but not
// page-locked double[]).  cudaMemcpyAsync from pageable memory is a synchronous bounce through the driver's small staging buffer
// (~3 GB/s measured on the B200 box, profiles/r01d_time_around.md); here worker threads copy chunks into / out of a page-locked ring
// while the DMA engine moves the previous chunks, so the copy runs near the PCIe rate.  Page-locked caller memory (vpc_host_alloc,
// vpc_host_register) is detected and copied directly.
//
// The reference has no counterpart: its points live in List<Point3D> objects (DataModel.cs:102-160); the shim flattens them
// into arrays for the call (csharp/DBImprovedGpu.cs).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace vpc_host {

class CopyPool {
 public:
  explicit CopyPool(int n_threads) {
    for (int t = 0; t < n_threads; ++t) th_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void submit(std::function<void()> f) {
    { std::lock_guard<std::mutex> lk(m_); q_.push_back(std::move(f)); }
    cv_.notify_one();
  }
  int threads() const { return (int)th_.size(); }

 private:
  void run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        f = std::move(q_.front()); q_.pop_front();
      }
      f();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  bool stop_ = false;
};

// One page-locked ring per device stream.  Not thread safe: the owning context serialises its calls.
class Stager {
 public:
  static constexpr size_t kChunk = 1u << 20;   // 1 MiB pieces: 16 of them cover the 1M-point cloud's x or y

  Stager() = default;
  Stager(const Stager&) = delete;
  ~Stager() { release(); }

  void release() {
    for (auto& s : slots_) if (s.ev) cudaEventDestroy(s.ev);
    slots_.clear();
    if (ring_) cudaFreeHost(ring_);
    ring_ = nullptr; cap_ = 0;
  }

  // page-locked (cudaHostAlloc / cudaHostRegister) memory goes straight to the DMA engine
  static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
  }

  // grows the ring (call while no transfer is pending: at the start of an API call); 2 chunks of head-room, capped
  cudaError_t reserve(size_t bytes) {
    const size_t want = ((std::min<size_t>(bytes, 128u << 20) + kChunk - 1) / kChunk) * kChunk + 2 * kChunk;
    if (want <= cap_ || !pending_.empty() || outstanding_.load() != 0) return cap_ ? cudaSuccess : cudaErrorNotReady;
    release();
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ring_), want, cudaHostAllocDefault);
    if (e != cudaSuccess) { ring_ = nullptr; return e; }
    cap_ = want;
    slots_.resize(cap_ / kChunk);
    out_busy_.reset(new std::atomic<int>[slots_.size()]);
    for (size_t k = 0; k < slots_.size(); ++k) { slots_[k].ev = nullptr; slots_[k].busy = false; out_busy_[k].store(0); }
    next_ = 0;
    return cudaSuccess;
  }

  // host -> device on `stream`; returns after every chunk has been handed to the DMA engine (the host range may be reused at once)
  cudaError_t h2d(CopyPool* pool, void* dev, const void* host, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (!pool || is_pinned(host)) return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream);
    cudaError_t e = cap_ ? cudaSuccess : reserve(bytes);
    if (e != cudaSuccess) return e;
    drain_pool_ = pool;
    const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    std::vector<int> slot_of(n_chunks);
    std::vector<std::atomic<int>> done(n_chunks);
    for (auto& d : done) d.store(0);
    size_t issued = 0, submitted = 0;
    // keep at most slots_.size() chunks in flight: a slot is reused once its DMA has completed
    while (issued < n_chunks) {
      while (submitted < n_chunks && submitted - issued < slots_.size()) {
        const int sl = acquire_slot();
        if (sl < 0) return cudaErrorUnknown;
        slot_of[submitted] = sl;
        const size_t off = submitted * kChunk, len = std::min(kChunk, bytes - off);
        char* dst = ring_ + (size_t)sl * kChunk;
        const char* src = static_cast<const char*>(host) + off;
        std::atomic<int>* flag = &done[submitted];
        pool->submit([dst, src, len, flag] { std::memcpy(dst, src, len); flag->store(1, std::memory_order_release); });
        ++submitted;
      }
      while (done[issued].load(std::memory_order_acquire) == 0) std::this_thread::yield();
      const size_t off = issued * kChunk, len = std::min(kChunk, bytes - off);
      Slot& s = slots_[slot_of[issued]];
      e = cudaMemcpyAsync(static_cast<char*>(dev) + off, ring_ + (size_t)slot_of[issued] * kChunk, len, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) return e;
      if (!s.ev && (e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(s.ev, stream)) != cudaSuccess) return e;
      s.busy = true;
      ++issued;
    }
    return cudaSuccess;
  }

  // device -> host on `stream`; the copies into the caller's memory are finished by finish()
  cudaError_t d2h(CopyPool* pool, void* host, const void* dev, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (!pool || is_pinned(host)) return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cap_ ? cudaSuccess : reserve(bytes);
    if (e != cudaSuccess) return e;
    const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    for (size_t c = 0; c < n_chunks; ++c) {
      const int sl = acquire_slot_for_d2h(pool);
      if (sl < 0) return cudaErrorUnknown;
      const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
      Slot& s = slots_[sl];
      e = cudaMemcpyAsync(ring_ + (size_t)sl * kChunk, static_cast<const char*>(dev) + off, len, cudaMemcpyDeviceToHost, stream);
      if (e != cudaSuccess) return e;
      if (!s.ev && (e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(s.ev, stream)) != cudaSuccess) return e;
      s.busy = true;
      pending_.push_back(Pending{sl, static_cast<char*>(host) + off, len});
    }
    return cudaSuccess;
  }

  // drains the device -> host pipeline: as each chunk's DMA completes a worker copies it into the caller's memory
  cudaError_t finish(CopyPool* pool) {
    if (!pool) return cudaSuccess;
    cudaError_t e = drain(pool, pending_.size());
    while (outstanding_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    return e;
  }

 private:
  struct Slot { cudaEvent_t ev; bool busy; };
  struct Pending { int slot; char* dst; size_t len; };

  int acquire_slot() {                       // next slot of the ring, once its previous DMA and its previous copy-out have completed
    const int sl = (int)next_;
    next_ = (next_ + 1) % slots_.size();
    Slot& s = slots_[sl];
    // a chunk that still waits to be copied out of this slot: drain the queue up to and including it
    for (size_t k = 0; k < pending_.size(); ++k)
      if (pending_[k].slot == sl) { if (drain(drain_pool_, k + 1) != cudaSuccess) return -1; break; }
    while (out_busy_[sl].load(std::memory_order_acquire) != 0) std::this_thread::yield();
    if (s.busy) { if (cudaEventSynchronize(s.ev) != cudaSuccess) return -1; s.busy = false; }
    return sl;
  }
  int acquire_slot_for_d2h(CopyPool* pool) { drain_pool_ = pool; return acquire_slot(); }
  cudaError_t drain(CopyPool* pool, size_t count) {
    for (size_t k = 0; k < count && !pending_.empty(); ++k) {
      Pending p = pending_.front(); pending_.pop_front();
      Slot& s = slots_[p.slot];
      cudaError_t e = cudaEventSynchronize(s.ev);
      if (e != cudaSuccess) return e;
      s.busy = false;
      const char* src = ring_ + (size_t)p.slot * kChunk;
      outstanding_.fetch_add(1, std::memory_order_acq_rel);
      out_busy_[p.slot].store(1, std::memory_order_release);
      std::atomic<int>* out = &outstanding_;
      std::atomic<int>* mine = &out_busy_[p.slot];
      pool->submit([p, src, out, mine] { std::memcpy(p.dst, src, p.len); mine->store(0, std::memory_order_release); out->fetch_sub(1, std::memory_order_acq_rel); });
    }
    return cudaSuccess;
  }

  char* ring_ = nullptr;
  size_t cap_ = 0, next_ = 0;
  std::vector<Slot> slots_;
  std::deque<Pending> pending_;
  std::atomic<int> outstanding_{0};
  std::unique_ptr<std::atomic<int>[]> out_busy_;
  CopyPool* drain_pool_ = nullptr;
};

}  // namespace vpc_host
