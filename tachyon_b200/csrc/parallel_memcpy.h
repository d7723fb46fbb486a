// A few persistent host threads that copy one buffer in slices.  Used to move PAGEABLE
// caller memory (the std::vector / Rust Vec a caller of the reference's C API passes,
// msm_gpu_unittest.cc:33-67, bn254_msm_gpu.cc:21-34) into pinned bounce buffers at several
// times the rate of a single memcpy, so the H2D stage of the range pipeline runs near PCIe
// speed instead of the ~11 GB/s of a pageable cudaMemcpyAsync.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace tb200 {

class ParallelMemcpy {
 public:
  explicit ParallelMemcpy(int threads) {
    if (threads < 1) threads = 1;
    for (int i = 1; i < threads; ++i) workers_.emplace_back([this, i] { Loop(i); });
    parts_ = threads;
  }
  ~ParallelMemcpy() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  ParallelMemcpy(const ParallelMemcpy&) = delete;
  ParallelMemcpy& operator=(const ParallelMemcpy&) = delete;

  int threads() const { return parts_; }

  // Blocking.  The calling thread copies slice 0.
  void Copy(void* dst, const void* src, size_t bytes) {
    if (parts_ == 1 || bytes < (size_t(1) << 20)) {
      memcpy(dst, src, bytes);
      return;
    }
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      pending_ = parts_ - 1;
      ++generation_;
    }
    cv_.notify_all();
    Slice(0);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void Slice(int i) {
    size_t per = (bytes_ / parts_ + 4095) & ~size_t(4095);
    size_t lo = per * i, hi = lo + per;
    if (lo >= bytes_) return;
    if (hi > bytes_ || i == parts_ - 1) hi = bytes_;
    memcpy(dst_ + lo, src_ + lo, hi - lo);
  }
  void Loop(int i) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      Slice(i);
      {
        std::lock_guard<std::mutex> l(mu_);
        --pending_;
      }
      done_.notify_one();
    }
  }

  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  uint64_t generation_ = 0;
  bool stop_ = false;
  int parts_ = 1, pending_ = 0;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0;
};

}  // namespace tb200
