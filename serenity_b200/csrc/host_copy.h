// host_copy.h - moving the caller's P / V between ordinary (pageable) host memory and the device.
//
// Serenity's matrices are Eigen objects in pageable memory.  cudaMemcpyAsync on such memory makes the driver stage the data with
// one thread (about 10 GB/s): 19 MB of P and V of a 1536-function system then cost more than 2 ms around a 5.8 ms build.  The
// host-buffer entry points therefore stage large transfers themselves: a few persistent worker threads copy between the
// caller's buffer and a page-locked staging buffer slice by slice while the DMA engine moves the slices that are ready.
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
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      next_.store(0);
      pending_ = (int)workers_.size();
      ++generation_;
    }
    cv_.notify_all();
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
      std::memcpy(dst_ + off, src_ + off, std::min(SLICE, bytes_ - off));
    }
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
  int pending_ = 0;
  unsigned long generation_ = 0;
  bool stop_ = false;
};

}  // namespace sxc
