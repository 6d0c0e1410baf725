// host_copy.h - moving the caller's P / V between ordinary (pageable) host memory and the device.
//
// Serenity's matrices are Eigen objects in pageable memory.  cudaMemcpyAsync on such memory makes the driver stage the data with
// one thread (about 10 GB/s): 19 MB of P and V of a 1536-function system then cost more than 2 ms around a 5.8 ms build.  The
// host-buffer entry points therefore stage large transfers themselves: a few persistent worker threads copy between the
// caller's buffer and a page-locked staging buffer slice by slice while the DMA engine moves the slices that are ready.
// Downloads are streamed: the job is opened once (one wake-up of the workers), the thread that owns the CUDA stream publishes how
// many bytes of the staging buffer the DMA engine has delivered so far, and the workers copy every slice below that mark - the
// memcpy of chunk c runs while chunk c + 1 is on the bus, with no per-chunk hand-shake.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sxc {

class HostCopier {
 public:
  explicit HostCopier(int nthreads) {
    for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostCopier() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  HostCopier(const HostCopier&) = delete;
  HostCopier& operator=(const HostCopier&) = delete;

  // memcpy(dst, src, bytes) with all workers and the calling thread; returns when the copy is complete
  void copy(void* dst, const void* src, size_t bytes) {
    if (bytes < 4 * SLICE || workers_.empty()) {
      std::memcpy(dst, src, bytes);
      return;
    }
    open(dst, src, bytes, bytes);
    close();
  }

  // streamed form: open() hands the job to the workers with the first `avail` bytes of src valid, publish() raises that mark
  // (monotonically, up to bytes), help() lets the caller copy slices that are ready without waiting, close() copies what is left
  // and returns when every slice has landed.  With no workers the caller does all of it in help() / close().
  void open(void* dst, const void* src, size_t bytes, size_t avail) {
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      avail_.store(avail, std::memory_order_release);
      next_.store(0);
      pending_ = (int)workers_.size();
      ++generation_;
    }
    cv_.notify_all();
  }
  void publish(size_t avail) { avail_.store(avail, std::memory_order_release); }
  // copies at most one slice; false when the next unclaimed slice is not available yet (or none is left)
  bool help() {
    size_t off = next_.load(std::memory_order_relaxed);
    for (;;) {
      if (off >= bytes_) return false;
      const size_t end = std::min(off + SLICE, bytes_);
      if (avail_.load(std::memory_order_acquire) < end) return false;
      if (next_.compare_exchange_weak(off, off + SLICE)) {
        std::memcpy(dst_ + off, src_ + off, end - off);
        return true;
      }
    }
  }
  void close() {
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  static constexpr size_t SLICE = 256 * 1024;
  void work() {
    for (;;) {
      const size_t off = next_.fetch_add(SLICE);
      if (off >= bytes_) break;
      const size_t end = std::min(off + SLICE, bytes_);
      while (avail_.load(std::memory_order_acquire) < end) cpu_relax();  // the DMA engine has not delivered this slice yet
      std::memcpy(dst_ + off, src_ + off, end - off);
    }
  }
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void loop() {
    unsigned long seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      lk.unlock();
      work();
      lk.lock();
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }

  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0;
  std::atomic<size_t> next_{0};
  std::atomic<size_t> avail_{0};
  int pending_ = 0;
  unsigned long generation_ = 0;
  bool stop_ = false;
};

}  // namespace sxc
