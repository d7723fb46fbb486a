// Host-side MSM engine: owns the device workspace and stream of one GPU and
// enqueues the kernel pipeline of msm_kernels.cuh.  Plays the role of
// tachyon/math/elliptic_curves/msm/variable_base_msm_gpu.h:11-31 +
// algorithms/icicle/icicle_msm.h:35-76 (context = mem pool + stream, Run(bases,
// scalars) -> one point), without icicle.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "host_math.h"
#include "msm_kernels.cuh"

namespace tb200 {

extern std::atomic<uint64_t> g_kernel_launches;

struct CudaError {
  cudaError_t code;
  const char* what;
  const char* file;
  int line;
};

#define TB_CUDA(expr)                                                         \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) throw CudaError{e_, #expr, __FILE__, __LINE__};    \
  } while (0)

// Grow-only device buffer.
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  void Reserve(size_t want) {
    if (want <= bytes) return;
    if (ptr) TB_CUDA(cudaFree(ptr));
    ptr = nullptr;
    bytes = 0;
    // round up to 2 MiB, 12.5 % slack so slightly larger follow-up calls reuse it
    size_t sz = want + want / 8;
    sz = (sz + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
    TB_CUDA(cudaMalloc(&ptr, sz));
    bytes = sz;
  }
  void Free() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(ptr);
  }
};

struct MsmTiming {
  float h2d_ms = 0, sort_ms = 0, accumulate_ms = 0, reduce_ms = 0, total_ms = 0, host_ms = 0;
  uint32_t window_bits = 0, windows = 0, tasks = 0, entries = 0, kernel_launches = 0, devices = 1;
};

struct MsmOptions {
  uint32_t window_bits = 0;  // 0 = choose from n
  uint32_t segment = 0;      // 0 = default
  int aggregate = -1;        // -1 = default
};

// Window choice.  Cost in units of one mixed addition:
//   sort + accumulation  kEntryCost  * n * W
//   reduction            kBucketCost * W * 2^(c-1)   (2 full additions per bucket)
// W is the smallest count with W * c >= bits + 1, so the signed top digit
// never carries out (see for_each_digit).
inline uint32_t WindowsFor(uint32_t bits, uint32_t c) { return (bits + 1 + c - 1) / c; }

// c >= 4 keeps W <= 64 (the size of the pinned result buffer).
constexpr uint32_t kMinWindowBits = 4;
constexpr uint32_t kMaxWindows = 64;

inline uint32_t ChooseWindowBits(size_t n, uint32_t scalar_bits) {
  // measured on B200 (BN254 2^24, c = 20): 0.155 ns per mixed addition, 0.057 ns of
  // sorting per entry, 0.73 ns of reduction per bucket
  constexpr double kEntryCost = 1.37, kBucketCost = 4.7;
  constexpr uint32_t kMaxBuckets = 1u << 24;  // scan_top_kernel capacity
  double best = 1e300;
  uint32_t best_c = kMinWindowBits;
  for (uint32_t c = kMinWindowBits; c <= 22; ++c) {
    uint32_t W = WindowsFor(scalar_bits, c);
    double buckets = (double)W * (double)(1u << (c - 1));
    if (buckets > kMaxBuckets) break;
    double cost = kEntryCost * (double)n * W + kBucketCost * buckets;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

template <class C>
class MsmEngine {
 public:
  using Fq = typename C::Fq;
  using Fr = typename C::Fr;
  using Point = HostXYZZ<Fq>;
  static constexpr size_t kAffineBytes = 2 * Fq::kLimbs64 * 8;
  static constexpr size_t kScalarBytes = Fr::kLimbs64 * 8;
  static constexpr size_t kXyzzBytes = 4 * Fq::kLimbs64 * 8;
  static constexpr int kXyzzWords = 4 * Fq::kLimbs32;
  // n * W must stay below 2^32 (u32 offsets) and n below 2^31 (sign bit)
  static constexpr size_t kMaxChunk = size_t(1) << 26;

  explicit MsmEngine(int device) : device_(device) {
    TB_CUDA(cudaSetDevice(device_));
    TB_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    for (auto& e : ev_) TB_CUDA(cudaEventCreate(&e));
    TB_CUDA(cudaMallocHost(&host_out_, kHostOutBytes));
    TB_CUDA(cudaMalloc(&totals_, sizeof(MsmTotals)));
    int sms = 0;
    TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
    sm_count_ = sms;
  }
  ~MsmEngine() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    for (DeviceBuffer* b : {&bases_stage_, &scalars_stage_, &count_, &offset_, &cursor_,
                            &task_base_, &tasks_, &multi_, &sorted_, &digits_, &task_out_, &block_sums_, &order_, &len_hist_,
                            &lvl_a_[0], &lvl_a_[1], &lvl_c_[0], &lvl_c_[1], &tree_[0], &tree_[1]})
      b->Free();
    if (totals_) cudaFree(totals_);
    if (host_out_) cudaFreeHost(host_out_);
    for (auto& e : ev_) cudaEventDestroy(e);
    cudaStreamDestroy(own_stream_);
  }
  MsmEngine(const MsmEngine&) = delete;
  MsmEngine& operator=(const MsmEngine&) = delete;

  int device() const { return device_; }
  void SetStream(cudaStream_t s) { stream_ = s ? s : own_stream_; }
  MsmOptions& options() { return options_; }
  const MsmTiming& timing() const { return timing_; }

  // bases / scalars: n elements each, host (pageable or pinned) or device
  // memory of this engine's device.  Blocking.
  Point Run(const void* bases, const void* scalars, size_t n) {
    TB_CUDA(cudaSetDevice(device_));
    timing_ = MsmTiming{};
    Point total = Point::Zero();
    if (n == 0) return total;  // pippenger_adapter.h:62-65
    // Sequential chunks only when the index arithmetic requires it; unlike
    // icicle_msm_bn254_g1.cc:56-62 the last chunk keeps its remainder.
    for (size_t off = 0; off < n; off += kMaxChunk) {
      size_t len = n - off < kMaxChunk ? n - off : kMaxChunk;
      Point part = RunChunk(static_cast<const char*>(bases) + off * kAffineBytes,
                            static_cast<const char*>(scalars) + off * kScalarBytes, len);
      total = (off == 0) ? part : total.Add(part);
    }
    return total;
  }

 private:
  static bool IsDevicePointer(const void* p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
  }

  template <class K, class... Args>
  void Launch(K kernel, uint32_t grid, uint32_t block, Args... args) {
    kernel<<<grid, block, 0, stream_>>>(args...);
    TB_CUDA(cudaGetLastError());
    ++launches_;
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }

  template <class K, class... Args>
  void LaunchGrid(K kernel, dim3 grid, uint32_t block, Args... args) {
    kernel<<<grid, block, 0, stream_>>>(args...);
    TB_CUDA(cudaGetLastError());
    ++launches_;
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }

  MsmPlan MakePlan(size_t n) const {
    MsmPlan p{};
    p.n = (uint32_t)n;
    p.c = options_.window_bits ? options_.window_bits : ChooseWindowBits(n, Fr::kBits);
    if (p.c < kMinWindowBits) p.c = kMinWindowBits;
    if (p.c > 24) p.c = 24;
    p.W = WindowsFor(Fr::kBits, p.c);
    p.B = 1u << (p.c - 1);
    p.TB = p.W * p.B;
    // Tasks hold up to 4x the mean bucket size, so ordinary buckets are one task and only
    // genuinely oversized buckets (skewed scalars) are split and folded.
    uint32_t seg = 128;
    while (seg < (uint32_t)kMaxSegment && seg < 4 * (n / p.B + 1)) seg <<= 1;
    p.seg = options_.segment ? options_.segment : seg;
    if (p.seg > (uint32_t)kMaxSegment) p.seg = kMaxSegment;
    p.aggregate = options_.aggregate < 0 ? 1u : (uint32_t)options_.aggregate;
    uint64_t entries = (uint64_t)n * p.W;
    uint64_t nonempty = entries < p.TB ? entries : p.TB;
    p.max_tasks = (uint32_t)(nonempty + entries / p.seg);
    return p;
  }

  Point RunChunk(const void* bases, const void* scalars, size_t n) {
    auto wall0 = std::chrono::steady_clock::now();
    MsmPlan plan = MakePlan(n);
    launches_ = 0;
    TB_CUDA(cudaEventRecord(ev_[0], stream_));

    // ---- inputs -----------------------------------------------------------
    const uint32_t* d_bases;
    const uint32_t* d_scalars;
    if (IsDevicePointer(scalars)) {
      d_scalars = static_cast<const uint32_t*>(scalars);
    } else {
      scalars_stage_.Reserve(n * kScalarBytes);
      TB_CUDA(cudaMemcpyAsync(scalars_stage_.ptr, scalars, n * kScalarBytes,
                              cudaMemcpyHostToDevice, stream_));
      d_scalars = scalars_stage_.as<uint32_t>();
    }
    if (IsDevicePointer(bases)) {
      d_bases = static_cast<const uint32_t*>(bases);
    } else {
      bases_stage_.Reserve(n * kAffineBytes);
      TB_CUDA(cudaMemcpyAsync(bases_stage_.ptr, bases, n * kAffineBytes, cudaMemcpyHostToDevice,
                              stream_));
      d_bases = bases_stage_.as<uint32_t>();
    }
    TB_CUDA(cudaEventRecord(ev_[1], stream_));

    // ---- workspace --------------------------------------------------------
    uint32_t scan_blocks = (plan.TB + kScanItems - 1) / kScanItems;
    if (scan_blocks > (uint32_t)kScanItems) throw CudaError{cudaErrorInvalidValue, "too many buckets", __FILE__, __LINE__};
    count_.Reserve((size_t)(plan.TB + 1) * 4);
    offset_.Reserve((size_t)(plan.TB + 1) * 4);
    cursor_.Reserve((size_t)(plan.TB + 1) * 4);
    task_base_.Reserve((size_t)plan.TB * 4);
    tasks_.Reserve((size_t)plan.max_tasks * sizeof(uint2));
    multi_.Reserve((size_t)plan.TB * 4);
    sorted_.Reserve((size_t)n * plan.W * 4);
    digits_.Reserve((size_t)n * plan.W * 4);
    task_out_.Reserve((size_t)plan.max_tasks * kXyzzBytes);
    block_sums_.Reserve((size_t)scan_blocks * 8);
    order_.Reserve((size_t)plan.max_tasks * 4);
    len_hist_.Reserve((size_t)(kMaxSegment + 1) * 4);

    // ---- sort: histogram, scan, tasks, scatter ----------------------------
    TB_CUDA(cudaMemsetAsync(count_.ptr, 0, (size_t)(plan.TB + 1) * 4, stream_));
    uint32_t sgrid = (plan.n + 255) / 256;
    Launch(digits_hist_kernel<C>, sgrid, 256, d_scalars, plan, digits_.as<uint32_t>(),
           count_.as<uint32_t>());
    Launch(scan_block_sums_kernel, scan_blocks, kScanThreads, count_.as<uint32_t>(), plan.TB,
           plan.seg, block_sums_.as<uint64_t>());
    Launch(scan_top_kernel, 1, kScanThreads, block_sums_.as<uint64_t>(), scan_blocks, totals_);
    Launch(scan_apply_build_tasks_kernel, scan_blocks, kScanThreads, count_.as<uint32_t>(),
           plan.TB, plan.seg, block_sums_.as<uint64_t>(), offset_.as<uint32_t>(),
           cursor_.as<uint32_t>(), task_base_.as<uint32_t>(), tasks_.as<uint2>(),
           multi_.as<uint32_t>(), totals_);
    LaunchGrid(digits_scatter_kernel, dim3(sgrid, plan.W), 256, digits_.as<uint32_t>(), plan,
               cursor_.as<uint32_t>(), sorted_.as<uint32_t>());
    // tasks by descending length
    TB_CUDA(cudaMemsetAsync(len_hist_.ptr, 0, (size_t)(kMaxSegment + 1) * 4, stream_));
    uint32_t ogrid = (plan.max_tasks + kOrderThreads * kOrderPerThread - 1) /
                     (kOrderThreads * kOrderPerThread);
    Launch(order_hist_kernel, ogrid, kOrderThreads, tasks_.as<uint2>(), totals_,
           len_hist_.as<uint32_t>());
    Launch(order_scan_kernel, 1, 1024, len_hist_.as<uint32_t>());
    Launch(order_scatter_kernel, ogrid, kOrderThreads, tasks_.as<uint2>(), totals_,
           len_hist_.as<uint32_t>(), order_.as<uint32_t>());
    TB_CUDA(cudaEventRecord(ev_[2], stream_));

    // ---- accumulate -------------------------------------------------------
    uint32_t agrid = (plan.max_tasks + kAccThreads - 1) / kAccThreads;
    Launch(accumulate_kernel<C>, agrid, kAccThreads, d_bases, sorted_.as<uint32_t>(),
           tasks_.as<uint2>(), order_.as<uint32_t>(), totals_, task_out_.as<uint32_t>());
    Launch(fold_partials_kernel<C>, sm_count_ * 4, kFoldThreads, multi_.as<uint32_t>(), totals_,
           offset_.as<uint32_t>(), task_base_.as<uint32_t>(), plan.seg,
           task_out_.as<uint32_t>());
    TB_CUDA(cudaEventRecord(ev_[3], stream_));

    // ---- bucket reduction: one blocked running-sum level, then a merge tree ---------
    uint32_t L0 = ChooseLevelLength(plan.B, plan.W);
    uint32_t m = plan.B / L0;  // blocks per window, a power of two
    lvl_a_[0].Reserve((size_t)plan.W * m * kXyzzBytes);
    lvl_c_[0].Reserve((size_t)plan.W * m * kXyzzBytes);
    {
      uint32_t threads = plan.W * m;
      Launch(reduce_level_kernel<C, true>, (threads + kReduceThreads - 1) / kReduceThreads,
             kReduceThreads, task_out_.as<uint32_t>(), (const uint32_t*)nullptr,
             offset_.as<uint32_t>(), task_base_.as<uint32_t>(), plan.B, m, L0, 0u, plan.W,
             lvl_a_[0].as<uint32_t>(), lvl_c_[0].as<uint32_t>());
    }
    uint32_t M = Log2(m);
    const uint32_t* tin = lvl_a_[0].as<uint32_t>();
    const uint32_t* tin_p = lvl_c_[0].as<uint32_t>();
    for (uint32_t s = 0; s < M; ++s) {
      uint32_t m_out = m >> (s + 1);
      DeviceBuffer& dst = tree_[s & 1];
      dst.Reserve((size_t)plan.W * m_out * (s + 3) * kXyzzBytes);
      uint32_t threads = plan.W * m_out * (s + 3);
      Launch(reduce_merge_kernel<C>, (threads + kReduceThreads - 1) / kReduceThreads,
             kReduceThreads, tin, tin_p, s, m_out, plan.W, dst.as<uint32_t>());
      tin = dst.as<uint32_t>();
      tin_p = nullptr;
    }
    // per window: (A, P, D_0 .. D_(M-1)); finished on the host
    uint32_t vals = M + 2;
    size_t win_bytes = (size_t)plan.W * vals * kXyzzBytes;
    if (M == 0) {
      TB_CUDA(cudaMemcpy2DAsync(host_out_, 2 * kXyzzBytes, lvl_a_[0].ptr, kXyzzBytes, kXyzzBytes,
                                plan.W, cudaMemcpyDeviceToHost, stream_));
      TB_CUDA(cudaMemcpy2DAsync(host_out_ + kXyzzBytes, 2 * kXyzzBytes, lvl_c_[0].ptr, kXyzzBytes,
                                kXyzzBytes, plan.W, cudaMemcpyDeviceToHost, stream_));
    } else {
      TB_CUDA(cudaMemcpyAsync(host_out_, tin, win_bytes, cudaMemcpyDeviceToHost, stream_));
    }
    TB_CUDA(cudaMemcpyAsync(host_out_ + win_bytes, totals_, sizeof(MsmTotals),
                            cudaMemcpyDeviceToHost, stream_));
    TB_CUDA(cudaEventRecord(ev_[4], stream_));
    TB_CUDA(cudaStreamSynchronize(stream_));

    // ---- host epilogue ------------------------------------------------------------
    // total = sum_w 2^(c w) [A_w + P_w + L0 sum_j 2^j D_(w,j)]: one Horner over bit
    // positions from the top (the c doublings per window of pippenger_base.h:59-77),
    // adding every term at its own bit, so the bucket-tree weights cost no extra doubling.
    auto host0 = std::chrono::steady_clock::now();
    const Point* hv = reinterpret_cast<const Point*>(host_out_);
    uint32_t l0 = Log2(L0);
    Point result = Point::Zero();
    for (uint32_t w = plan.W; w-- > 0;) {
      const Point* v = hv + (size_t)w * vals;
      for (uint32_t bit = plan.c; bit-- > 0;) {
        if (w + 1 < plan.W || bit + 1 < plan.c) result = result.Dbl();
        if (bit >= l0 && bit - l0 < M) result = result.Add(v[2 + (bit - l0)]);
        if (bit == 0) result = result.Add(v[0]).Add(v[1]);
      }
    }
    auto host1 = std::chrono::steady_clock::now();

    MsmTotals tot;
    memcpy(&tot, host_out_ + win_bytes, sizeof(tot));
    float ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
    timing_.h2d_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev_[1], ev_[2]));
    timing_.sort_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
    timing_.accumulate_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev_[3], ev_[4]));
    timing_.reduce_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[4]));
    timing_.total_ms += ms;
    timing_.host_ms += std::chrono::duration<float, std::milli>(host1 - host0).count();
    timing_.window_bits = plan.c;
    timing_.windows = plan.W;
    timing_.tasks += tot.tasks;
    timing_.entries += tot.entries;
    timing_.kernel_launches += launches_;
    (void)wall0;
    return result;
  }

  static uint32_t Log2(uint32_t x) {
    uint32_t r = 0;
    while ((1u << r) < x) ++r;
    return r;
  }

  // Blocks of the running-sum level: as long as possible while the grid still fills the
  // chip (a thread costs 2 L full additions, the merge tree ~3 per block).
  uint32_t ChooseLevelLength(uint32_t buckets_per_window, uint32_t windows) const {
    uint64_t items = (uint64_t)buckets_per_window * windows;
    uint64_t want_threads = (uint64_t)sm_count_ * 384;
    uint32_t L = 64;
    while (L > 4 && items / L < want_threads) L >>= 1;
    if (L > buckets_per_window) L = buckets_per_window;
    return L;
  }

  // per window (A, P, D_0..D_(M-1)), M <= 22, + totals
  static constexpr size_t kHostOutBytes = kMaxWindows * 24 * kXyzzBytes + 64;

  int device_;
  int sm_count_ = 148;
  cudaStream_t own_stream_ = nullptr;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev_[5];
  MsmOptions options_;
  MsmTiming timing_;
  uint32_t launches_ = 0;
  MsmTotals* totals_ = nullptr;
  char* host_out_ = nullptr;
  DeviceBuffer bases_stage_, scalars_stage_, count_, offset_, cursor_, task_base_, tasks_, multi_,
      sorted_, digits_, task_out_, block_sums_, order_, len_hist_, lvl_a_[2], lvl_c_[2], tree_[2];
};

}  // namespace tb200
