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
#include <initializer_list>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "host_math.h"
#include "nccl_dl.h"
#include "parallel_memcpy.h"
#include "msm_kernels.cuh"
#include "msm_sort.cuh"

namespace tb200 {

extern std::atomic<uint64_t> g_kernel_launches;

struct CudaError {
  cudaError_t code;
  const char* what;
  const char* file;
  int line;
  int device = -1;  // filled in by multi-device callers: which GPU's worker failed
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
  bool in_arena = false;  // a slice of the engine's workspace arena (not freed on its own)
  // 2 MiB granules, 12.5 % slack so slightly larger follow-up calls reuse the buffer
  static size_t Rounded(size_t want) {
    size_t sz = want + want / 8;
    return (sz + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
  }
  void Reserve(size_t want) {
    if (want <= bytes) return;
    Free();
    size_t sz = Rounded(want);
    TB_CUDA(cudaMalloc(&ptr, sz));
    bytes = sz;
  }
  void Free() {
    if (ptr && !in_arena) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    in_arena = false;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(ptr);
  }
};

struct MsmTiming {
  float h2d_ms = 0, sort_ms = 0, accumulate_ms = 0, reduce_ms = 0, total_ms = 0, host_ms = 0;
  uint32_t window_bits = 0, windows = 0, tasks = 0, entries = 0, kernel_launches = 0, devices = 1;
  uint32_t ranges = 0;
  float enqueue_ms = 0, wait_ms = 0;  // host wall clock: queueing the work / blocked on the device
  uint32_t pair_rounds = 0;
  // the accumulate launches that run alone on the device (every range's launch except the low
  // window group's, which overlaps the high group's reduction): what the kernel roofline uses
  float acc_kernel_ms = 0;
  uint32_t acc_kernel_entries = 0;
  uint32_t low_windows = 0;  // windows in the low group (0 = the windows were not split)
  float combine_ms = 0;      // window_combine_kernel of the high group (hidden when split)
};

struct MsmOptions {
  uint32_t window_bits = 0;  // 0 = choose from n
  uint32_t segment = 0;      // 0 = default
  int aggregate = -1;        // -1 = default
  uint32_t ranges = 0;       // point ranges per MSM; 0 = 1 for device inputs, pipelined for host
  int sample_scalars = 1;    // choose the window from a sample of the scalars' bit lengths
  int sort_mode = -1;        // 0 = one-level atomic counting sort, 1 = two-level shared-memory
                             // sort (msm_sort.cuh) where eligible, -1 = automatic
  int pair_rounds = -1;      // batched-affine pair rounds before the XYZZ accumulation:
                             // -1 = none (default), -2 = from bucket occupancy, >= 0 = forced
  uint32_t host_ranges = 8;  // most ranges the automatic host-input pipeline cuts an MSM into
  int reduce_mode = 1;       // 1 = two threads per block of buckets (reduce_blocks_kernel; G2: two
                             // lane pairs, reduce_blocks_pair_kernel), 2 = two threads per block
                             // for G2 as well, 0 = one thread per block (reduce_level_kernel)
  uint32_t level_fill = 0;   // blocks per SM the running-sum level wants before it shortens its
                             // blocks (0 = default: 384, 768 from 1.5 M bucket slots)
  int balance = 1;           // balanced windows (WideWindowsFor); 0 = equal widths, slack on top
  int precompute = 0;        // RegisterBases also builds the table of window multiples
                             // 2^(bit offset of window w) * P (precompute_factor of
                             // icicle_msm.h:21, here always the full factor): MSMs over the
                             // registered bases then use ONE set of buckets for all windows
  int acc_variant = -1;      // G2 groups: 0 = one thread per accumulation task (accumulate_kernel),
                             // 1 = one LANE PAIR per task, a lane per Fq2 component
                             // (accumulate_pair_kernel; the default), 2 = the same at the
                             // alternative register budget (PairMinBlocksAlt).
                             // G1 groups: 0 = free-running warps, 3 = the warps of a CTA in step
                             // (accumulate_lockstep_kernel), -1 = the curve's default
  int acc_lockstep = -1;     // G2 pair kernel: 1 = warps of a CTA in step, 0 = free, -1 = curve default
  int reduce_inline = -1;    // G1 running-sum kernel: 1 = both roles share one inlined call site of the
                             // addition, 0 = out-of-line addition, -1 = the curve's default
  int reduce_roll = -1;      // code shape of the field multiplications in the running-sum kernel:
                             // 0 = unrolled, 1 = looped multiplications (fp_mul_rolled: a third of
                             // the code), 2 = looped, squarings through the multiplier too,
                             // -1 = the curve's measured default (C::kReduceRoll)
  int stage_points = 0;      // 1 = accumulate_staged_kernel: the next point of a task travels
                             // through shared memory (cp.async) instead of registers
  int device_ladder = 0;     // where the final ladder sum_w 2^(offset of w) S_w runs: 0 = on the
                             // host (default: ~255 strictly sequential point doublings take
                             // 65 us on one host core, 430 us on the GPU's fastest form,
                             // tools/probe/chain_probe.cu), 1 = on the device, hidden behind the
                             // accumulation of the low windows where the cost model finds room
  int low_windows = -1;      // device ladder: windows of the low group (accumulated last, while the high group's
                             // reduction and window combination run on the tail stream):
                             // -1 = from the cost model, 0 = no split
};

// Window choice.  Cost in units of one mixed addition:
//   sort + accumulation  kEntryCost  * n * W
//   reduction            kBucketCost * W * 2^(c-1)   (2 full additions per bucket)
// W is the smallest count with W * c >= bits + 1, so the signed top digit
// never carries out (see for_each_digit).
inline uint32_t WindowsFor(uint32_t bits, uint32_t c) { return (bits + 1 + c - 1) / c; }

// Balanced windows.  W * c usually exceeds bits + 1, and with equal widths the whole slack
// lands in the top window: BN254 at c = 20 has 13 windows for 255 bits, the top one holds 15
// bits, i.e. 2^14 of its 2^19 buckets take all n entries (1024 per bucket at 2^24 points, split
// into tasks and folded again, in every point range).  Taking one bit from each of the top
// `slack` windows instead gives every window a full set of 2^(cw-1) buckets and leaves fewer
// non-empty buckets to reduce (8 * 2^19 + 5 * 2^18 instead of 13 * 2^19).  Returns how many
// (low) windows keep c bits.
inline uint32_t WideWindowsFor(uint32_t bits, uint32_t c, bool balance = true) {
  uint32_t W = WindowsFor(bits, c);
  if (!balance || c < 5) return W;
  uint32_t slack = W * c - (bits + 1);
  // slack >= W would make every window c - 1 bits wide: that is the plan of c - 1 with twice
  // the bucket slots, so leave such a c unbalanced (the cost model then never prefers it)
  return slack >= W ? W : W - slack;
}
// Buckets that can be non-empty.
inline double PopulatedBuckets(uint32_t bits, uint32_t c, bool balance = true) {
  uint32_t W = WindowsFor(bits, c), wide = WideWindowsFor(bits, c, balance);
  return (double)wide * (double)(1u << (c - 1)) + (double)(W - wide) * (double)(1u << (c - 2));
}

// c >= 4 keeps W <= 64 (the size of the pinned result buffer).
constexpr uint32_t kMinWindowBits = 4;
constexpr uint32_t kMaxWindows = 64;

// Every point range after the first pays the O(buckets) bookkeeping again (counter reset, scan,
// task list, one read-modify-write of each bucket value): ~0.65 ms per range for the 5.5 M
// buckets of a 2^24-point BN254 MSM, i.e. 0.76 mixed additions per bucket and range.
constexpr double kRangeBucketCost = 0.76;

// Mixed additions the accumulation kernel is charged for, uniform scalars.  Throughput-bound it
// is n * W; but the kernel ends when its LONGEST task ends, and one task is one bucket: with few,
// long buckets (small n, or the narrow top windows of a balanced plan, whose buckets hold twice
// the entries) the resident threads all wait for ~mean + 4 sigma sequential additions.  Measured
// (B200, BN254): 2^18 points 1.98 ms at c = 14 (110 K buckets of 45 / 90 entries) against 1.25 ms
// at c = 15; 2^17 1.19 -> 0.89 ms.
inline double AccumulateWork(size_t n, uint32_t scalar_bits, uint32_t c, bool balance = true,
                             uint32_t resident_threads = 148 * 512) {
  const uint32_t W = WindowsFor(scalar_bits, c), wide = WideWindowsFor(scalar_bits, c, balance);
  const double per_bucket = (double)n / (double)(1u << (c - (wide < W ? 2 : 1)));  // narrow windows
  const double longest = per_bucket + 4.0 * std::sqrt(per_bucket);
  const double tasks = PopulatedBuckets(scalar_bits, c, balance);
  // a thread's addition takes 1 / (resident threads) of the chip's throughput at full occupancy
  // (10.5 us) and never less than ~5.8 us, the latency of one mixed addition with the SM to
  // itself (2^14 points: 20 K tasks of ~55 additions take 0.33 ms, 74 K tasks of ~19 0.13 ms)
  double running = tasks < (double)resident_threads ? tasks : (double)resident_threads;
  if (running < 0.55 * resident_threads) running = 0.55 * resident_threads;
  const double throughput = (double)n * W, critical = longest * running;
  return throughput > critical ? throughput : critical;
}

// Sorting costs the same 0.057 ns per entry for every group while a mixed addition costs 0.155 ns
// (BN254 G1) to 1.3 ns (BLS12-381 G2): the entry cost in units of one mixed addition.
inline double EntryCost(double madd_ns) { return 1.0 + 0.057 / madd_ns; }

inline uint32_t ChooseWindowBits(size_t n, uint32_t scalar_bits, size_t ranges = 1,
                                 double kEntryCost = 1.37) {
  // measured on B200 (BN254 2^24, c = 20): 0.155 ns per mixed addition, 0.057 ns of
  // sorting per entry, 0.75 ns of reduction per bucket (2 full additions + the merge tree = 5.4
  // mixed additions at 0.138 ns, the same ratio for every field; 2^22 points: c = 17 10.8 ms,
  // c = 19 11.1 ms)
  const double kBucketCost = 5.4 + kRangeBucketCost * (double)(ranges > 1 ? ranges - 1 : 0);
  constexpr uint32_t kMaxBuckets = 1u << 24;  // scan_top_kernel capacity
  double best = 1e300;
  uint32_t best_c = kMinWindowBits;
  for (uint32_t c = kMinWindowBits; c <= 22; ++c) {
    uint32_t W = WindowsFor(scalar_bits, c);
    double buckets = (double)W * (double)(1u << (c - 1));
    if (buckets > kMaxBuckets) break;
    // with several point ranges every range is its own accumulation of n / ranges points
    const size_t per_range = ranges > 1 ? (n + ranges - 1) / ranges : n;
    double cost = kEntryCost * AccumulateWork(per_range, scalar_bits, c) * (double)(ranges > 1 ? ranges : 1) +
                  kBucketCost * PopulatedBuckets(scalar_bits, c);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// Window choice when all windows share one bucket set (precomputed table of window multiples):
// the reduction covers 2^(c-1) buckets once instead of once per window, so larger windows pay.
inline uint32_t ChooseWindowBitsShared(size_t n, uint32_t scalar_bits, double kEntryCost = 1.37,
                                       uint32_t resident_threads = 148 * 512) {
  constexpr double kBucketCost = 5.4;
  double best = 1e300;
  uint32_t best_c = kMinWindowBits;
  for (uint32_t c = kMinWindowBits; c <= 24; ++c) {
    uint32_t W = WindowsFor(scalar_bits, c);
    if ((uint64_t)n * W >= (uint64_t(1) << 31)) continue;  // table index + sign bit in one word
    // the same longest-task bound as AccumulateWork: B buckets of n W / B entries on average
    // (the lower half, shared by the narrow windows, holds more)
    const double buckets = (double)(1u << (c - 1));
    const double per_bucket = (double)n * W / buckets * 1.5;
    const double longest = per_bucket + 4.0 * std::sqrt(per_bucket);
    double running = buckets < (double)resident_threads ? buckets : (double)resident_threads;
    if (running < 0.55 * resident_threads) running = 0.55 * resident_threads;
    const double work = (double)n * W > longest * running ? (double)n * W : longest * running;
    double cost = kEntryCost * work + kBucketCost * buckets;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// The same cost model with the scalars' actual lengths: `bit_hist[b]` = how many of `samples`
// sampled scalars have bit length b (0 = the zero scalar).  A scalar of b bits contributes
// about ceil(b / c) non-zero digits, so witness-like vectors (mostly 0 / 1 / small values)
// have far fewer entries per point than the uniform draw the plain model assumes, and a
// smaller window (fewer buckets to reduce) wins.
inline uint32_t ChooseWindowBitsSampled(size_t n, uint32_t scalar_bits, const uint32_t* bit_hist,
                                        uint32_t samples, size_t ranges = 1, double kEntryCost = 1.37) {
  const double kBucketCost = 5.4 + kRangeBucketCost * (double)(ranges > 1 ? ranges - 1 : 0);
  constexpr uint32_t kMaxBuckets = 1u << 24;
  double best = 1e300;
  uint32_t best_c = kMinWindowBits;
  for (uint32_t c = kMinWindowBits; c <= 22; ++c) {
    uint32_t W = WindowsFor(scalar_bits, c);
    double buckets = (double)W * (double)(1u << (c - 1));
    if (buckets > kMaxBuckets) break;
    double digits = 0;
    for (uint32_t b = 1; b <= scalar_bits + 1; ++b)
      if (bit_hist[b]) digits += (double)bit_hist[b] * (double)((b + c - 1) / c);
    double entries = (double)n * digits / (double)samples;
    // every (point, window) slot is still recoded, written and read once by the sort
    double cost = kEntryCost * entries + kBucketCost * PopulatedBuckets(scalar_bits, c) +
                  0.05 * (double)n * W;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// Host element type of a device field kind.
template <class K>
struct HostElOf;
template <class F>
struct HostElOf<FpField<F>> {
  using type = HostFp<F>;
};
template <class F>
struct HostElOf<Fp2Field<F>> {
  using type = HostFp2<F>;
};
template <class C>
using HostElT = typename HostElOf<typename C::Field>::type;

template <class C>
class MsmEngine {
 public:
  using Fr = typename C::Fr;
  using Point = HostPointXYZZ<HostElT<C>>;
  static constexpr size_t kElBytes = C::Field::kWords * 4;
  static constexpr size_t kAffineBytes = 2 * kElBytes;
  static constexpr size_t kScalarBytes = Fr::kLimbs64 * 8;
  static constexpr size_t kXyzzBytes = 4 * kElBytes;
  static constexpr int kXyzzWords = 4 * C::Field::kWords;
  // n * W must stay below 2^32 (u32 offsets) and n below 2^31 (sign bit)
  static constexpr size_t kMaxPiece = size_t(1) << 26;
  static constexpr size_t kMaxRanges = 64;   // point ranges per MSM
  static constexpr size_t kStageSlots = 3;   // H2D staging ring
  static constexpr size_t kBounceSlots = 4;  // pinned bounce buffers for pageable sources
  static constexpr size_t kBounceBytes = size_t(16) << 20;
  static constexpr uint32_t kMaxPairRounds = 4;
  static constexpr uint32_t kMinPairBatch = 96;  // pairs per thread below which a round is skipped

  explicit MsmEngine(int device) : device_(device) {
    TB_CUDA(cudaSetDevice(device_));
    TB_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    TB_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    TB_CUDA(cudaStreamCreateWithFlags(&sample_stream_, cudaStreamNonBlocking));
    {
      // the tail stream (reduction + window combination of the high windows) outranks the
      // compute stream: its small latency-bound kernels must get SM slots as soon as CTAs of
      // the concurrent low-window accumulation retire
      int lo_prio = 0, hi_prio = 0;
      TB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
      TB_CUDA(cudaStreamCreateWithPriority(&tail_stream_, cudaStreamNonBlocking, hi_prio));
    }
    TB_CUDA(cudaMallocHost(&host_out_, 2 * kHostOutBytes));
    TB_CUDA(cudaMalloc(&totals_, sizeof(MsmTotals)));
    int sms = 0;
    TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
    sm_count_ = sms;
  }
  ~MsmEngine() {
    LeaveRanks();
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    cudaStreamSynchronize(tail_stream_);
    cudaStreamSynchronize(copy_stream_);
    for (const DeviceBuffer* b : AllBuffers()) const_cast<DeviceBuffer*>(b)->Free();
    if (arena_) cudaFree(arena_);
    if (totals_) cudaFree(totals_);
    if (host_out_) cudaFreeHost(host_out_);
    if (bounce_) cudaFreeHost(bounce_);
    for (auto& e : events_) cudaEventDestroy(e);
    cudaStreamDestroy(copy_stream_);
    cudaStreamDestroy(sample_stream_);
    cudaStreamDestroy(tail_stream_);
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
    if (n == 0) return world_ > 1 ? GatherHostPoint(total) : total;  // pippenger_adapter.h:62-65
    // Independent pieces only when the u32 index arithmetic requires it (entries = points x
    // windows must stay below 2^32, which a forced small window can hit before 2^26 points);
    // unlike icicle_msm_bn254_g1.cc:56-62 the last piece keeps its remainder.
    const size_t piece = PieceLimit(n);
    if (n <= piece) {  // the usual case: with several ranks the gather rides the MSM's stream
      Pending p = Enqueue(bases, scalars, n, 0, false, /*gather=*/world_ > 1);
      Point part = Finish(p);
      return (world_ > 1 && !p.gathered) ? GatherHostPoint(part) : part;
    }
    for (size_t off = 0; off < n; off += piece) {
      size_t len = n - off < piece ? n - off : piece;
      Pending p = Enqueue(static_cast<const char*>(bases) + off * kAffineBytes,
                          static_cast<const char*>(scalars) + off * kScalarBytes, len, 0);
      Point part = Finish(p);
      total = (off == 0) ? part : total.Add(part);
    }
    return world_ > 1 ? GatherHostPoint(total) : total;
  }

  // Point-range sharding over `world` processes, one GPU each: after this call every
  // Run() returns the SUM of all ranks' partial sums.  The exchange is one ncclAllGather of the
  // ranks' XYZZ partials, issued on this engine's stream right behind the window-combination
  // kernel; the `world` points are added on the host ("only 8 points are combined").  `id` is
  // the 128-byte ncclUniqueId that rank 0 obtained from tachyon_b200_nccl_unique_id().
  void JoinRanks(const void* id, int rank, int world) {
    TB_CUDA(cudaSetDevice(device_));
    std::string why;
    const NcclApi* nccl = NcclApi::Get(&why);
    if (!nccl) {
      static std::string keep;
      keep = why;
      throw CudaError{cudaErrorNotSupported, keep.c_str(), __FILE__, __LINE__};
    }
    LeaveRanks();
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    int rc = nccl->CommInitRank(&comm_, world, uid, rank);
    if (rc != 0) throw CudaError{cudaErrorUnknown, nccl->GetErrorString(rc), __FILE__, __LINE__};
    rank_ = rank;
    world_ = world;
    TB_CUDA(cudaMalloc(&gather_dev_, (size_t)world * kXyzzBytes));
    TB_CUDA(cudaMallocHost(&gather_host_, 2 * (size_t)world * kXyzzBytes));
  }
  void LeaveRanks() {
    if (!comm_) return;
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    if (const NcclApi* nccl = NcclApi::Get(nullptr)) nccl->CommDestroy(comm_);
    comm_ = nullptr;
    world_ = 1;
    rank_ = 0;
    if (gather_dev_) cudaFree(gather_dev_);
    if (gather_host_) cudaFreeHost(gather_host_);
    gather_dev_ = nullptr;
    gather_host_ = nullptr;
  }
  int world() const { return world_; }

  // Frees the grow-only workspace (not the registered bases); the next call re-allocates.
  void ReleaseWorkspace() {
    TB_CUDA(cudaSetDevice(device_));
    TB_CUDA(cudaStreamSynchronize(copy_stream_));
    TB_CUDA(cudaStreamSynchronize(stream_));
    TB_CUDA(cudaStreamSynchronize(tail_stream_));
    for (const DeviceBuffer* b : AllBuffers())
      if (b != &registered_) const_cast<DeviceBuffer*>(b)->Free();
    if (arena_) cudaFree(arena_);
    arena_ = nullptr;
    for (auto& u : stage_used_) u = false;
    budget_ = 0;
  }
  size_t workspace_bytes() const { return OwnedBytes(); }

  // Allocates everything an n-point MSM with host inputs needs (workspace, staging ring,
  // bounce buffers, copy threads) so that the first call does not pay for it — what the
  // advisory `degree` of tachyon_<c>_g1_create_msm_gpu is good for (msm_gpu.h:35 ignores it).
  void Prewarm(size_t n) {
    TB_CUDA(cudaSetDevice(device_));
    if (n == 0) return;
    if (n > kMaxPiece) n = kMaxPiece;
    Enqueue(nullptr, nullptr, n, 0, /*reserve_only=*/true);
    EnsureBounce();
  }

  // Keeps a private device copy of `n` bases (host or device source) for later MSMs — the
  // SRS of kzg.h:91-113, uploaded once instead of once per commitment.
  // With option "precompute" the copy is followed by the table of window multiples
  // T[w][i] = 2^(bit offset of window w) * P_i for the window size the cost model picks for n
  // points sharing one bucket set; slice 0 of the table is the bases themselves.
  void RegisterBases(const void* bases, size_t n) {
    TB_CUDA(cudaSetDevice(device_));
    TB_CUDA(cudaStreamSynchronize(stream_));
    TB_CUDA(cudaStreamSynchronize(copy_stream_));
    table_c_ = 0;
    uint32_t W = 1;
    if (options_.precompute && n) {
      table_c_ = options_.window_bits ? options_.window_bits
                                      : ChooseWindowBitsShared(n, Fr::kBits, EntryCost(MaddNanos()));
      if (table_c_ < kMinWindowBits) table_c_ = kMinWindowBits;
      W = WindowsFor(Fr::kBits, table_c_);
      if ((uint64_t)n * W >= (uint64_t(1) << 31))
        throw CudaError{cudaErrorInvalidValue, "precomputed table too large for 31-bit indices", __FILE__,
                        __LINE__};
    }
    registered_.Reserve(n * W * kAffineBytes);
    // on the engine's stream: a plain cudaMemcpy from pageable memory may return while its last
    // staged chunk is still on its way, and this (non-blocking) stream does not wait for the
    // legacy stream — the table kernel below read stale bytes beyond the first MiB
    if (n) {
      TB_CUDA(cudaMemcpyAsync(registered_.ptr, bases, n * kAffineBytes, cudaMemcpyDefault, stream_));
      TB_CUDA(cudaStreamSynchronize(stream_));
    }
    registered_n_ = n;
    if (table_c_) {
      const uint32_t wide = WideWindowsFor(Fr::kBits, table_c_, options_.balance != 0);
      table_wide_ = wide;
      // in place: the kernel reads slice 0 (the bases) and writes every slice
      precompute_table_kernel<C><<<(uint32_t)((n + 127) / 128), 128, 0, stream_>>>(
          registered_.as<uint32_t>(), (uint32_t)n, (uint32_t)n, W, table_c_, wide,
          registered_.as<uint32_t>());
      TB_CUDA(cudaGetLastError());
      g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
      TB_CUDA(cudaStreamSynchronize(stream_));
    }
  }
  uint32_t table_window_bits() const { return table_c_; }
  const void* registered_bases() const { return registered_.ptr; }
  size_t registered_size() const { return registered_n_; }

  // `count` MSMs, MSM i over bases[i] (nullptr: the registered bases) and scalars[i], both
  // sizes[i] long — the commit loop of tachyon/crypto/commitments/kzg/kzg.h:217-313 and the
  // G1 queries of zk/r1cs/groth16/prove.h:100-131.  Two MSMs are kept in flight: the inputs
  // of MSM i+1 cross PCIe and its kernels are queued while MSM i runs, and the host
  // epilogue of MSM i overlaps the device work of MSM i+1.
  void RunBatch(const void* const* bases, const void* const* scalars, const size_t* sizes,
                size_t count, Point* out) {
    TB_CUDA(cudaSetDevice(device_));
    timing_ = MsmTiming{};
    size_t biggest = 0;
    for (size_t i = 0; i < count; ++i) {
      biggest = sizes[i] > biggest ? sizes[i] : biggest;
      if (!bases[i] && sizes[i] > registered_n_)
        throw CudaError{cudaErrorInvalidValue, "MSM larger than the registered bases", __FILE__,
                        __LINE__};
    }
    auto bases_of = [&](size_t i) { return bases[i] ? bases[i] : registered_.ptr; };
    if (biggest > PieceLimit(biggest) && !table_c_) {  // rare: fall back to one blocking call each
      MsmTiming sum{};
      for (size_t i = 0; i < count; ++i) {
        out[i] = Run(bases_of(i), scalars[i], sizes[i]);
        sum.total_ms += timing_.total_ms;
        sum.kernel_launches += timing_.kernel_launches;
      }
      timing_ = sum;
      return;
    }
    auto wall0 = std::chrono::steady_clock::now();
    Pending pend[2];
    bool live[2] = {false, false};
    for (size_t i = 0; i < count + 1; ++i) {
      int slot = (int)(i & 1);
      if (i < count) {
        if (sizes[i] == 0) {
          out[i] = Point::Zero();
        } else {
          // the other slot's MSM is still in flight: growing a buffer would wait for it,
          // which is correct (cudaFree synchronises) but serialises; sizes are usually equal
          // from the second MSM on the copies overlap the PREVIOUS MSM's kernels, so one
          // range is enough and the per-range bookkeeping is saved
          in_batch_tail_ = live[slot ^ 1];
          pend[slot] = Enqueue(bases_of(i), scalars[i], sizes[i], slot, false, false,
                               /*table=*/!bases[i] && table_c_ != 0);
          in_batch_tail_ = false;
          pend[slot].index = i;
          live[slot] = true;
        }
      }
      int prev = slot ^ 1;
      if (i >= 1 && live[prev]) {
        out[pend[prev].index] = Finish(pend[prev]);
        live[prev] = false;
      }
    }
    timing_.total_ms =
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  }

 private:
  // reduction-tree buffers of a window group: leaves (A_t, P_t), two scratch regions, two
  // stage outputs, all cut into per-window slices of 2 * (blocks per window) points
  struct TreeBuffers {
    DeviceBuffer leaves, ping, pong, out[2];
  };

  // One enqueued MSM: everything the host epilogue needs once the device is done.
  struct Pending {
    MsmPlan plan{};
    MsmPlan red{};  // the plan as the reduction / window combination see it (ReductionPlan)
    uint32_t L0 = 0, M = 0;
    uint32_t low = 0;     // windows of the low group
    uint32_t L0_low = 0;  // its running-sum block length
    size_t host_bytes = 0;  // bytes that crossed PCIe for this MSM
    bool gathered = false;  // the result is the all-gathered set of rank partials
    bool device_ladder = false;  // the device left one point (else: one sum per window)
    size_t K = 0;
    bool any_host = false;
    int slot = 0;       // host result buffer / event set
    uint32_t launches = 0;
    size_t index = 0;
  };

  static bool IsDevicePointer(const void* p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
  }

  static bool IsPageable(const void* p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
  }

  void EnsureBounce() {
    if (bounce_) return;
    TB_CUDA(cudaMallocHost(&bounce_, kBounceSlots * kBounceBytes));
    int hw = (int)std::thread::hardware_concurrency();
    int want = hw >= 4 ? (hw * 3 / 4 > 16 ? 16 : hw * 3 / 4) : 1;
    if (const char* e = getenv("TACHYON_B200_COPY_THREADS")) want = atoi(e);
    copier_.reset(new ParallelMemcpy(want));
  }

  // Host -> device on the copy stream.  Pinned sources go straight to the DMA engine;
  // pageable ones are copied by a few host threads into a ring of pinned bounce buffers
  // first (a pageable cudaMemcpyAsync runs at ~11 GB/s on this platform, PCIe at ~55).
  void CopyToDevice(void* dst, const void* src, size_t bytes, bool pageable) {
    if (!pageable || bytes < (size_t(8) << 20)) {  // small copies: the driver's own staging wins
      TB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, copy_stream_));
      return;
    }
    EnsureBounce();
    for (size_t off = 0; off < bytes; off += kBounceBytes) {
      size_t len = bytes - off < kBounceBytes ? bytes - off : kBounceBytes;
      size_t slot = bounce_seq_++ % kBounceSlots;
      cudaEvent_t ev = Event(2 * kEventsPerSlot + kStageSlots + slot);
      if (bounce_used_[slot]) TB_CUDA(cudaEventSynchronize(ev));  // its last DMA has drained
      char* b = bounce_ + slot * kBounceBytes;
      copier_->Copy(b, static_cast<const char*>(src) + off, len);
      TB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, b, len, cudaMemcpyHostToDevice,
                              copy_stream_));
      TB_CUDA(cudaEventRecord(ev, copy_stream_));
      bounce_used_[slot] = true;
    }
  }

  template <class K, class... Args>
  void Launch(K kernel, uint32_t grid, uint32_t block, Args... args) {
    kernel<<<grid, block, 0, stream_>>>(args...);
    TB_CUDA(cudaGetLastError());
    ++launches_;
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }

  template <class K, class... Args>
  void LaunchOn(cudaStream_t st, K kernel, uint32_t grid, uint32_t block, Args... args) {
    kernel<<<grid, block, 0, st>>>(args...);
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

  // Plan of one point range of `n` points inside an MSM whose window size is `c`.
  MsmPlan MakePlan(size_t n, uint32_t c, bool table = false) const {
    MsmPlan p{};
    p.n = (uint32_t)n;
    p.c = c;
    p.W = WindowsFor(Fr::kBits, p.c);
    p.B = 1u << (p.c - 1);
    p.shared = table ? 1u : 0u;
    p.stride = table ? (uint32_t)registered_n_ : 0u;
    p.TB = table ? p.B : p.W * p.B;
    p.wide = table ? table_wide_ : WideWindowsFor(Fr::kBits, p.c, options_.balance != 0);
    // Tasks hold up to 4x the mean bucket size, so ordinary buckets are one task and only
    // genuinely oversized buckets (skewed scalars) are split and folded.
    uint32_t seg = 128;
    const uint32_t run_buckets = p.wide < p.W ? p.B / 2 : p.B;  // narrow windows: half the buckets
    const size_t per_set = table ? n * p.W : n;                // entries per bucket set
    while (seg < (uint32_t)kMaxSegment && seg < 4 * (per_set / run_buckets + 1)) seg <<= 1;
    p.seg = options_.segment ? options_.segment : seg;
    if (p.seg > (uint32_t)kMaxSegment) p.seg = kMaxSegment;
    p.aggregate = options_.aggregate < 0 ? 1u : (uint32_t)options_.aggregate;
    uint64_t entries = (uint64_t)n * p.W;
    uint64_t nonempty = entries < p.TB ? entries : p.TB;
    // p.seg is the UPPER limit: choose_segment_kernel picks the actual task length on the
    // device from the measured bucket occupancy, down to kMinSegment
    uint32_t seg_floor = options_.segment ? p.seg : kMinSegment;
    p.max_tasks = (uint32_t)(nonempty + entries / seg_floor);
    // Pair rounds (batched-affine pre-reduction) are OFF unless asked for: measured on B200
    // they lose to the XYZZ path at every size (BN254 2^24: accumulate 31.4 ms -> 37.7 ms at
    // R = 3).  "pair_rounds" = -2 picks R from the bucket occupancy:
    // leave ~4-8 points per bucket for the XYZZ stage and stop while a round still has enough
    // pairs for every thread of a full grid to amortise its inversion.
    uint32_t R = 0;
    if (options_.pair_rounds >= 0) {
      R = (uint32_t)options_.pair_rounds;
    } else if (options_.pair_rounds == -2) {
      uint64_t mean = n / p.B;
      while (R < kMaxPairRounds && (mean >> (R + 3)) >= 1 &&
             (entries >> (R + 1)) >= (uint64_t)sm_count_ * 512 * kMinPairBatch)
        ++R;
    }
    if (R > kMaxPairRounds) R = kMaxPairRounds;
    p.R = R;
    return p;
  }

  // The two-level sort needs at least two coarse bins per window and at most
  // kMaxCoarsePerWindow of them; tiny ranges are not worth its extra passes.
  bool UseTwoLevelSort(const MsmPlan& p) const {
    if (p.shared) return false;  // shared buckets: the one-level sort (its keys carry no window)
    bool eligible = p.c >= 12 && (p.B >> kFineBits) <= kMaxCoarsePerWindow && p.n >= (1u << 15);
    if (options_.sort_mode == 0) return false;
    if (options_.sort_mode == 1) return eligible;
    // Automatic: measured on B200 with uniform scalars the two-level sort wins once a window
    // has >= 2^18 buckets (c >= 19: 2^24 points 4.44 -> 3.78 ms, 2^23 2.30 -> 2.19) and loses a
    // few percent below (2^21: 0.63 -> 0.68 ms).  On skewed scalars it wins everywhere (witness-
    // like 2^24: 4.13 -> 2.27 ms), but the skew is not known before the first pass.
    // Skewed scalars (flagged by the sample): the two-level sort wins from ~2^20 points per range
    // (witness-like 2^24: sort 4.09 -> 2.51 ms, 2^20: 0.455 -> 0.374); on the ~2^18-point ranges
    // of a pipelined 2^20 MSM its extra passes cost more than the atomics they avoid.
    return eligible && (p.c >= 19 || (skew_hint_ && p.n >= (1u << 20)));
  }
  static SortPlan MakeSortPlan(const MsmPlan& p) {
    SortPlan sp{};
    sp.Cw = p.B >> kFineBits;
    sp.regions = p.W * sp.Cw;
    sp.max_tiles = (uint32_t)(((uint64_t)p.n * p.W) / kSortTile + sp.regions + 1);
    return sp;
  }

  // Upper bound of the padded entry count of a range (what totals->entries can reach).
  static uint64_t PaddedBound(const MsmPlan& p) {
    uint64_t entries = (uint64_t)p.n * p.W;
    uint64_t nonempty = entries < p.TB ? entries : p.TB;
    uint64_t a = (uint64_t(1) << p.R) - 1;
    return ((entries + a * nonempty) + a) & ~a;
  }

  // Window size from a sample of the scalars (every n / 1024-th one): bit-length histogram ->
  // ChooseWindowBitsSampled.  Host scalars are read in place; device scalars cost one small
  // strided D2H copy (~30 us), so tiny MSMs keep the size-only rule.
  uint32_t WindowBitsFromSample(const void* scalars, size_t n, bool scalars_dev, size_t ranges) {
    // Measured: the sampled choice pays from ~2^22 points (witness-like 2^24: 13.9 -> 10.0 ms);
    // below, the few heavy buckets of such vectors dominate and the size-only window is as good.
    // The sample also tells whether the scalars are skewed (many tiny values: all of them meet in
    // a handful of buckets, where the one-level sort's atomics serialise): then the two-level
    // shared-memory sort is used at any eligible window size.  Host scalars are sampled in place
    // from 2^16 points; device scalars (one strided D2H copy) from 2^22.
    skew_hint_ = false;
    if (!options_.sample_scalars || n < (size_t(1) << (scalars_dev ? 22 : 16)))
      return WindowBitsFor(n, ranges);
    constexpr uint32_t kSamples = 1024;
    using FrEl = HostFp<Fr>;
    static_assert(sizeof(FrEl) == kScalarBytes, "scalar layout");
    const size_t stride = n / kSamples;
    FrEl sample[kSamples];
    if (scalars_dev) {
      // on the copy stream: the compute stream may still be busy with the previous MSM of a batch
      TB_CUDA(cudaMemcpy2DAsync(sample, kScalarBytes, scalars, stride * kScalarBytes, kScalarBytes,
                                kSamples, cudaMemcpyDeviceToHost, sample_stream_));
      TB_CUDA(cudaStreamSynchronize(sample_stream_));
    } else {
      for (uint32_t k = 0; k < kSamples; ++k)
        memcpy(&sample[k], static_cast<const char*>(scalars) + k * stride * kScalarBytes,
               kScalarBytes);
    }
    uint32_t hist[Fr::kBits + 2] = {};
    FrEl one = FrEl::Zero();
    one.v[0] = 1;
    for (uint32_t k = 0; k < kSamples; ++k) {
      FrEl canon = sample[k].Mul(one);  // from Montgomery
      uint32_t bits = 0;
      for (int i = FrEl::N; i-- > 0;) {
        if (canon.v[i]) {
          bits = 64 * i + 64 - (uint32_t)__builtin_clzll(canon.v[i]);
          break;
        }
      }
      if (bits > Fr::kBits + 1) bits = Fr::kBits + 1;  // unreduced garbage: treat as full length
      hist[bits]++;
    }
    uint32_t tiny = 0;
    for (uint32_t b = 0; b <= 8; ++b) tiny += hist[b];
    skew_hint_ = tiny * 20 > kSamples;  // more than 5 % of the scalars are below 2^8
    if (options_.window_bits || n < (size_t(1) << 22)) return WindowBitsFor(n, ranges);
    uint32_t c = ChooseWindowBitsSampled(n, Fr::kBits, hist, kSamples, ranges, EntryCost(MaddNanos()));
    uint32_t by_size = WindowBitsFor(n, ranges);
    if (c > by_size) c = by_size;  // the sample may only argue for FEWER buckets
    if (c < kMinWindowBits) c = kMinWindowBits;
    while (c < 22 && (uint64_t)n * WindowsFor(Fr::kBits, c) > 0xE0000000ull) ++c;  // u32 offsets
    return c;
  }

  // Largest number of points one piece may hold: n * W (+ padding) must fit u32 offsets.
  size_t PieceLimit(size_t n) const {
    size_t lim = kMaxPiece;
    uint32_t W = WindowsFor(Fr::kBits, WindowBitsFor(n < lim ? n : lim));
    size_t by_entries = (size_t)0xE0000000u / W;
    return by_entries < lim ? by_entries : lim;
  }

  uint32_t WindowBitsFor(size_t n, size_t ranges = 1) const {
    uint32_t c =
        options_.window_bits ? options_.window_bits
                             : ChooseWindowBits(n, Fr::kBits, ranges, EntryCost(MaddNanos()));
    if (c < kMinWindowBits) c = kMinWindowBits;
    if (c > 24) c = 24;
    return c;
  }

  // Device bytes one point range of m points needs besides the bucket values (the role of
  // the footprint model in icicle_msm_utils.cc:10-68, for this pipeline's buffers).
  size_t RangeFootprint(size_t m, uint32_t c, bool stage_bases, bool stage_scalars,
                        bool table = false) const {
    MsmPlan p = MakePlan(m, c, table);
    size_t b = (size_t)m * p.W * 4 + PaddedBound(p) * 4;  // digits + sorted
    if (p.R) b += PaddedBound(p) * (kAffineBytes + kAffineBytes / 4);  // pair outputs + prefixes
    if (UseTwoLevelSort(p)) b += (size_t)m * p.W * 8;                  // coarse-sorted entries
    b += (size_t)p.max_tasks * (8 + 4 + 4 + kXyzzBytes);  // tasks, meta, order, task_out
    b += ((size_t)p.max_tasks / kFoldThreads + p.TB) * 8;  // fold jobs
    b += (size_t)(p.TB + 1) * 4 * 5;                      // count, offset, cursor, task_base, multi
    if (stage_bases) b += (size_t)m * kAffineBytes * kStageSlots;
    if (stage_scalars) b += (size_t)m * kScalarBytes * kStageSlots;
    return b;
  }

  cudaEvent_t Event(size_t i) {
    while (events_.size() <= i) {
      cudaEvent_t e;
      TB_CUDA(cudaEventCreate(&e));
      events_.push_back(e);
    }
    return events_[i];
  }
  // event sets: per Pending slot, 0 begin, 1 end of accumulation, 2 end, 3 copy begin, 4 high
  // window group accumulated, 5 high group combined (tail stream), 6 / 7 around the high group's
  // window_combine_kernel, 8 low group's tree done, 9 high group's sum available to the compute
  // stream, 10 tail stream: high group's running-sum level done, then per range r: 12 + 5r copied,
  // +1 sort start, +2 sort end, +3 accumulated, +4 after the range's first accumulate launch
  static constexpr size_t kSlotEvents = 12, kRangeEvents = 5;
  static constexpr size_t kEventsPerSlot = kSlotEvents + kRangeEvents * kMaxRanges;
  cudaEvent_t SlotEvent(int slot, size_t i) { return Event((size_t)slot * kEventsPerSlot + i); }
  cudaEvent_t StageFreeEvent(size_t stage) { return Event(2 * kEventsPerSlot + stage); }

  // Grow-only workspace.  A buffer that has to grow is freed and reallocated, which is only
  // safe once nothing in flight uses it: drain both streams first.
  void ReserveAll(std::initializer_list<std::pair<DeviceBuffer*, size_t>> wants) {
    bool grow = false;
    for (auto& w : wants) grow = grow || w.second > w.first->bytes;
    if (!grow) return;
    static const bool trace = getenv("TACHYON_B200_TRACE") != nullptr;
    // TACHYON_B200_ARENA=0: one cudaMalloc per buffer (the round-1 layout)
    static const bool use_arena = !(getenv("TACHYON_B200_ARENA") && atoi(getenv("TACHYON_B200_ARENA")) == 0);
    // experiment hook: a gap of TACHYON_B200_PAD_MB between consecutive buffers
    static const size_t pad = (getenv("TACHYON_B200_PAD_MB") ? (size_t)atol(getenv("TACHYON_B200_PAD_MB")) : 0) << 20;
    auto t0 = std::chrono::steady_clock::now();
    TB_CUDA(cudaStreamSynchronize(copy_stream_));
    TB_CUDA(cudaStreamSynchronize(stream_));
    TB_CUDA(cudaStreamSynchronize(tail_stream_));
    if (use_arena) {
      // ONE allocation for the whole workspace, cut into 2 MiB-aligned slices in list order: the
      // placement of the buffers relative to each other no longer depends on what the driver's
      // allocator did before (sizes stay grow-only per buffer; contents are scratch)
      std::vector<size_t> sz;
      size_t total = 0;
      for (auto& w : wants) {
        size_t have = w.first->bytes, s = w.second > have ? DeviceBuffer::Rounded(w.second) : have;
        sz.push_back(s);
        total += s + (s ? pad : 0);
      }
      for (auto& w : wants) w.first->Free();
      if (arena_) TB_CUDA(cudaFree(arena_));
      arena_ = nullptr;
      TB_CUDA(cudaMalloc(&arena_, total ? total : 1));
      size_t off = 0, k = 0;
      for (auto& w : wants) {
        if (sz[k]) {
          w.first->ptr = static_cast<char*>(arena_) + off;
          w.first->bytes = sz[k];
          w.first->in_arena = true;
          off += sz[k] + pad;
        }
        ++k;
      }
      if (trace)
        fprintf(stderr, "[tachyon_b200] workspace arena: %zu bytes at %p, %.2f ms\n", total, arena_,
                std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
    } else {
      int index = 0;
      for (auto& w : wants) {
        if (trace && w.second > w.first->bytes) {
          auto t1 = std::chrono::steady_clock::now();
          size_t had = w.first->bytes;
          w.first->Reserve(w.second);
          fprintf(stderr, "[tachyon_b200] workspace buffer %d grows %zu -> %zu bytes at %p: %.2f ms\n", index,
                  had, w.first->bytes, w.first->ptr,
                  std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t1).count());
        } else {
          w.first->Reserve(w.second);
        }
        ++index;
      }
      if (trace)
        fprintf(stderr, "[tachyon_b200] workspace growth took %.2f ms\n",
                std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    for (auto& u : stage_used_) u = false;
    budget_ = 0;
  }

  // G2: one accumulation task per lane pair (accumulate_pair_kernel).  acc_variant picks the
  // register budget: 1 (and the default) = the first, 2 = the second entry of PairShapes<C>.
  template <int kThreads, int kMinBlocks>
  void LaunchPairShape(uint32_t max_tasks, const uint32_t* d_bases, uint32_t part) {
    if constexpr (C::Field::kDegree == 2) {
      uint32_t grid = (uint32_t)(((uint64_t)max_tasks * 2 + kThreads - 1) / kThreads);
      if (grid == 0) grid = 1;
      const bool lockstep = options_.acc_lockstep < 0 ? C::kAccLockstep : options_.acc_lockstep != 0;
      Launch(lockstep ? accumulate_pair_kernel<C, kThreads, kMinBlocks, true>
                      : accumulate_pair_kernel<C, kThreads, kMinBlocks, false>,
             grid, (uint32_t)kThreads, d_bases,
             (const uint32_t*)sorted_.as<uint32_t>(), (const uint2*)tasks_.as<uint2>(),
             (const uint32_t*)task_meta_.as<uint32_t>(), (const uint32_t*)order_.as<uint32_t>(),
             (const MsmTotals*)totals_, part, state_.as<uint32_t>(), task_out_.as<uint32_t>());
    }
  }
  // G2 running-sum level on lane pairs (reduce_blocks_pair_kernel)
  void LaunchPairReduce(cudaStream_t st, uint32_t blocks, const uint32_t* bucket0, uint32_t B,
                        uint32_t nb, uint32_t L0, uint32_t wn, uint32_t wide_local, uint32_t* leaves) {
    if constexpr (C::Field::kDegree == 2) {
      constexpr uint32_t kSlots = kPairReduceSlots;
      constexpr int kMin = PairMinBlocks<C>();
      const int roll = options_.reduce_roll < 0 ? C::kReduceRoll : options_.reduce_roll;
      auto kernel = roll ? reduce_blocks_pair_kernel<C, 1, kMin> : reduce_blocks_pair_kernel<C, 0, kMin>;
      LaunchOn(st, kernel, (blocks + kSlots - 1) / kSlots, 4 * kSlots, bucket0, B, nb, L0, wn,
               wide_local, leaves, leaves + kXyzzWords);
    }
  }
  void LaunchLockstep(uint32_t agrid, const uint32_t* d_bases, uint32_t part) {
    if constexpr (C::Field::kDegree == 1) {
      Launch(accumulate_lockstep_kernel<C>, agrid, (uint32_t)kAccThreads, d_bases,
             (const uint32_t*)sorted_.as<uint32_t>(), (const uint2*)tasks_.as<uint2>(),
             (const uint32_t*)task_meta_.as<uint32_t>(), (const uint32_t*)order_.as<uint32_t>(),
             (const MsmTotals*)totals_, part, state_.as<uint32_t>(), task_out_.as<uint32_t>());
    }
  }
  void LaunchPairAccumulate(uint32_t max_tasks, const uint32_t* d_bases, uint32_t part) {
    if constexpr (C::Field::kDegree == 2) {
      constexpr int kFirst = PairMinBlocks<C>(), kSecond = PairMinBlocksAlt<C>();
      if (options_.acc_variant == 2)
        LaunchPairShape<kAccThreads, kSecond>(max_tasks, d_bases, part);
      else
        LaunchPairShape<kAccThreads, kFirst>(max_tasks, d_bases, part);
    }
  }

  // One MSM: bucket values live in `state_` for the whole call; the points are consumed
  // as K consecutive ranges, each range sorted by bucket and added into the bucket values
  // (accumulate_kernel starts from the value the earlier ranges left).  With host inputs
  // the ranges are what gets pipelined: range k+1 crosses PCIe on the copy stream while
  // range k is sorted and accumulated, so the 1.6 GB of a 2^24 BN254 MSM (~29 ms at PCIe
  // gen5 rates) hides behind the ~36 ms of bucket work instead of preceding it.  Device
  // inputs run as one range unless the memory model asks for more.
  //
  // Enqueue() only queues work (copies on the copy stream through a ring of staging slots,
  // kernels and the final D2H on the compute stream); Finish() waits and runs the host
  // epilogue.  All device buffers except the staging ring are shared by consecutive MSMs,
  // which is safe because their kernels are ordered on the one compute stream.
  Pending Enqueue(const void* bases, const void* scalars, size_t n, int slot,
                  bool reserve_only = false, bool gather = false, bool table = false) {
    auto wall0 = std::chrono::steady_clock::now();
    const bool bases_dev = IsDevicePointer(bases), scalars_dev = IsDevicePointer(scalars);
    const bool bases_pageable = !bases_dev && IsPageable(bases);
    const bool scalars_pageable = !scalars_dev && IsPageable(scalars);
    Pending pd;
    pd.any_host = !bases_dev || !scalars_dev;
    pd.host_bytes = (bases_dev ? 0 : n * kAffineBytes) + (scalars_dev ? 0 : n * kScalarBytes);
    pd.slot = slot;

    // ---- how many ranges ----------------------------------------------------------
    size_t K = 1;
    if (options_.ranges > 0) {
      K = options_.ranges;
    } else if (pd.any_host && !in_batch_tail_) {
      K = n >> 18;  // ranges of >= 2^18 points on average, at most 8
      if (K > options_.host_ranges) K = options_.host_ranges;
      if (K < 1) K = 1;
    }
    if (K > kMaxRanges) K = kMaxRanges;
    if (K > n) K = n;
    // the window is chosen knowing the range count: every extra range repeats the per-bucket
    // bookkeeping, so pipelined host inputs prefer a slightly smaller window
    if (reserve_only) skew_hint_ = false;
    if (table) skew_hint_ = false;
    const uint32_t c = table ? table_c_
                             : (reserve_only ? WindowBitsFor(n, K) : WindowBitsFromSample(scalars, n, scalars_dev, K));
    MsmPlan whole = MakePlan(n, c, table);
    bool memory_bound = false;
    {
      // Memory model (icicle_msm_utils.cc:10-68 analogue): more ranges when one range's
      // buffers would not fit.  cudaMemGetInfo costs milliseconds, so the driver is asked
      // only when this call needs more than the engine already owns.
      const size_t state_b = (size_t)whole.TB * kXyzzBytes;
      auto need = [&](size_t k) {
        return state_b + RangeFootprint((n + k - 1) / k, c, !bases_dev, !scalars_dev, table);
      };
      if (need(K) > OwnedBytes()) {
        if (budget_ == 0) {  // refreshed whenever the workspace grows (ReserveAll)
          size_t free_b = 0, total_b = 0;
          TB_CUDA(cudaMemGetInfo(&free_b, &total_b));
          budget_ = (size_t)(0.9 * (double)(free_b + OwnedBytes()));
        }
        while (K < kMaxRanges && K < n && need(K) > budget_) {
          K *= 2;
          memory_bound = true;
        }
        if (K > kMaxRanges) K = kMaxRanges;
      }
    }
    // Range boundaries.  Equal ranges when the split is forced or memory-driven.  For the
    // automatic host-input pipeline the ranges GROW geometrically: nothing can start before
    // range 0 has crossed PCIe, so it is small, and range k+1 may be as much larger than
    // range k as the copy is faster than the bucket work (ratio q), so the copy engine still
    // stays ahead of the kernels.
    std::vector<size_t> bound(K + 1, n);
    bound[0] = 0;
    if (K > 1 && pd.any_host && options_.ranges == 0 && !memory_bound) {
      // kRangeGrowth assumes the ~55 GB/s one GPU gets alone.  When the last host-input call saw
      // less (4-8 GPUs of one box copying at once: 24-30 GB/s each), the copy, not the bucket
      // work, sets the pace and what remains after the last byte has arrived is the processing of
      // the LAST range: ranges stop growing (equal ranges; 2^22 points per GPU on 4 GPUs:
      // 17.6 -> 16.2 ms end to end).
      double q = C::kRangeGrowth * (h2d_gbs_ > 0 ? h2d_gbs_ / 55.0 : 1.0);
      if (q < 1.02) q = 1.02;
      double f = (q - 1.0) / (std::pow(q, (double)K) - 1.0), acc = 0;
      for (size_t r = 1; r < K; ++r) {
        acc += f;
        f *= q;
        size_t b = (size_t)(acc * (double)n) & ~size_t(255);
        bound[r] = b > bound[r - 1] ? b : bound[r - 1];
      }
    } else {
      const size_t eq = (n + K - 1) / K;
      for (size_t r = 1; r < K; ++r) bound[r] = r * eq < n ? r * eq : n;
    }
    size_t m = 0;  // the largest range: what the per-range workspace is sized for
    for (size_t r = 0; r < K; ++r) m = bound[r + 1] - bound[r] > m ? bound[r + 1] - bound[r] : m;
    MsmPlan big = MakePlan(m, c, table);
    const MsmPlan red = ReductionPlan(big);
    pd.plan = big;
    pd.red = red;
    pd.K = K;
    launches_ = 0;
    pd.device_ladder = options_.device_ladder != 0;
    gather = gather && pd.device_ladder;  // host ladder: the partial is gathered after Finish()
    size_t last_len = 0;  // size of the last non-empty range
    for (size_t r = 0; r < K; ++r)
      if (bound[r + 1] > bound[r]) last_len = bound[r + 1] - bound[r];
    pd.low = (reserve_only || !pd.device_ladder || table) ? 0u : ChooseLowWindows(big, last_len);

    // ---- workspace ----------------------------------------------------------------
    // staging ring: kStageSlots slots of one 256-byte-aligned range each, plus one alignment unit
    // per slot, so that the stride derived from the (grow-only) buffer below always holds a range
    const size_t bases_slot_want = (m * kAffineBytes + 255) / 256 * 256 + 256;
    const size_t scalars_slot_want = (m * kScalarBytes + 255) / 256 * 256 + 256;
    uint32_t scan_blocks = (big.TB + kScanItems - 1) / kScanItems;
    if (scan_blocks > (uint32_t)kScanItems)
      throw CudaError{cudaErrorInvalidValue, "too many buckets", __FILE__, __LINE__};
    pd.L0 = ChooseLevelLength(big.B, red.W);
    const uint32_t nb = big.B / pd.L0;  // blocks per window, a power of two
    pd.M = Log2(nb);
    // reduction tree buffers: per window a slice of 2 nb points (the leaves (A_t, P_t)); the
    // scratch and stage-output buffers use the same slicing (reduce_tree_kernel)
    size_t tree_b = 0;
    auto reduction_bytes = [&](uint32_t cc) {
      uint32_t Wc = table ? 1u : WindowsFor(Fr::kBits, cc), Bc = 1u << (cc - 1);
      uint32_t nbc = Bc / ChooseLevelLength(Bc, Wc);
      size_t lv = (size_t)Wc * nbc * 2 * kXyzzBytes;
      if (lv > tree_b) tree_b = lv;
    };
    reduction_bytes(c);
    // Pre-reserving for 2^degree points: smaller MSMs run smaller windows with MORE windows, and
    // these few-MB buffers scale with W — cover every window size a smaller call may choose, so
    // that no later call has to grow the workspace (a growth drains both streams).
    if (reserve_only)
      for (uint32_t cc = kMinWindowBits; cc < c; ++cc) reduction_bytes(cc);
    // the low window group has its own, shorter running-sum blocks (its reduction is what
    // remains exposed at the end of the MSM: latency, not throughput, counts) and its own buffers
    pd.L0_low = LowLevelLength(red, pd.low);
    const size_t tree_lo_b = (size_t)pd.low * (big.B / pd.L0_low) * 2 * kXyzzBytes;
    ReserveAll({{&state_, (size_t)big.TB * kXyzzBytes},
                {&count_, (size_t)(big.TB + 1) * 4},
                {&offset_, (size_t)(big.TB + 1) * 4},
                {&cursor_, (size_t)(big.TB + 1) * 4},
                {&task_base_, (size_t)big.TB * 4},
                {&multi_, (size_t)big.TB * 4},
                {&fold_jobs_, ((size_t)big.max_tasks / kFoldThreads + big.TB + 1) * sizeof(uint2)},
                {&nonzero_slots_, (size_t)kNonzeroSlots * 4},
                {&tasks_, (size_t)big.max_tasks * sizeof(uint2)},
                {&task_meta_, (size_t)big.max_tasks * 4},
                {&order_, (size_t)big.max_tasks * 4},
                {&task_out_, (size_t)big.max_tasks * kXyzzBytes},
                {&sorted_, (size_t)PaddedBound(big) * 4},
                {&mid_, UseTwoLevelSort(big) ? (size_t)m * big.W * 8 : 0},
                {&coarse_, UseTwoLevelSort(big) ? (size_t)(MakeSortPlan(big).regions + 1) * 4 * 4 + 64 : 0},
                {&pair_prefix_, big.R ? (size_t)((PaddedBound(big) >> 1) + PairThreads(big, 0)) *
                                            (kAffineBytes / 2)
                                      : 0},
                {&pair_out_[0], big.R > 0 ? (size_t)(PaddedBound(big) >> 1) * kAffineBytes : 0},
                {&pair_out_[1], big.R > 1 ? (size_t)(PaddedBound(big) >> 2) * kAffineBytes : 0},
                {&pair_out_[2], big.R > 2 ? (size_t)(PaddedBound(big) >> 3) * kAffineBytes : 0},
                {&pair_out_[3], big.R > 3 ? (size_t)(PaddedBound(big) >> 4) * kAffineBytes : 0},
                {&digits_, (size_t)m * big.W * 4},
                {&block_sums_, (size_t)scan_blocks * 8},
                {&len_hist_, (size_t)kOrderBins * 4},
                {&tree_hi_.leaves, tree_b},
                {&tree_hi_.ping, tree_b},
                {&tree_hi_.pong, tree_b},
                {&tree_hi_.out[0], tree_b},
                {&tree_hi_.out[1], tree_b},
                {&tree_lo_.leaves, tree_lo_b},
                {&tree_lo_.ping, tree_lo_b},
                {&tree_lo_.pong, tree_lo_b},
                {&tree_lo_.out[0], tree_lo_b},
                {&tree_lo_.out[1], tree_lo_b},
                {&combine_, (size_t)(2 * kTermSlots + 2 + kMaxWindows) * kXyzzBytes},
                {&bases_stage_, bases_dev ? 0 : bases_slot_want * kStageSlots},
                {&scalars_stage_, scalars_dev ? 0 : scalars_slot_want * kStageSlots}});
    // The slot stride is a property of the BUFFER, not of the call: MSMs of different sizes are in
    // flight together (RunBatch) and the per-slot events only protect a fixed slot geometry; it
    // changes only when the buffer is re-allocated, which drains both streams first.  The
    // reservation above guarantees stride >= this call's range.
    const size_t bases_slot_bytes = bases_stage_.bytes / kStageSlots / 256 * 256;
    const size_t scalars_slot_bytes = scalars_stage_.bytes / kStageSlots / 256 * 256;
    if ((!bases_dev && bases_slot_bytes < m * kAffineBytes) ||
        (!scalars_dev && scalars_slot_bytes < m * kScalarBytes))
      throw CudaError{cudaErrorInvalidValue, "staging slot smaller than a point range", __FILE__,
                      __LINE__};
    static const bool trace = getenv("TACHYON_B200_TRACE") != nullptr;
    auto since = [&] {
      return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    };
    if (trace)
      fprintf(stderr, "[tachyon_b200] enqueue n=%zu c=%u W=%u K=%zu%s: planned + reserved at %.2f ms\n", n,
              c, big.W, K, reserve_only ? " (reserve only)" : "", since());
    if (reserve_only) return pd;

    cudaEvent_t ev_begin = SlotEvent(slot, 0), ev_acc_end = SlotEvent(slot, 1),
                ev_end = SlotEvent(slot, 2), ev_copy_begin = SlotEvent(slot, 3);
    auto ev = [&](size_t r, int which) { return SlotEvent(slot, kSlotEvents + kRangeEvents * r + which); };

    TB_CUDA(cudaEventRecord(ev_begin, stream_));
    TB_CUDA(cudaMemsetAsync(state_.ptr, 0, (size_t)big.TB * kXyzzBytes, stream_));
    if (pd.any_host) TB_CUDA(cudaEventRecord(ev_copy_begin, copy_stream_));

    char* host_out = host_out_ + (size_t)slot * kHostOutBytes;
    // the window groups are formed in the last NON-EMPTY range (a forced range count can leave
    // empty ranges at the end, which queue no kernels at all)
    size_t last_range = 0;
    for (size_t r = 0; r < K; ++r)
      if (bound[r + 1] > bound[r]) last_range = r;
    for (size_t r = 0; r < K; ++r) {
      const size_t lo = bound[r], len = bound[r + 1] - bound[r];
      if (len == 0) {  // degenerate split: keep the event bookkeeping of Finish() simple
        TB_CUDA(cudaEventRecord(ev(r, 0), copy_stream_));
        for (int w = 1; w <= 4; ++w) TB_CUDA(cudaEventRecord(ev(r, w), stream_));
        memset(host_out + kHostPartialBytes + r * sizeof(MsmTotals), 0, sizeof(MsmTotals));
        continue;
      }
      MsmPlan plan = MakePlan(len, c, table);
      // ---- inputs of this range ---------------------------------------------------
      const uint32_t* d_bases;
      const uint32_t* d_scalars;
      size_t stage = 0;
      if (pd.any_host) {
        stage = stage_seq_++ % kStageSlots;
        // the slot's previous tenant must have been consumed
        if (stage_used_[stage]) TB_CUDA(cudaStreamWaitEvent(copy_stream_, StageFreeEvent(stage), 0));
      }
      if (scalars_dev) {
        d_scalars = reinterpret_cast<const uint32_t*>(static_cast<const char*>(scalars) +
                                                      lo * kScalarBytes);
      } else {
        char* dst = scalars_stage_.as<char>() + stage * scalars_slot_bytes;
        CopyToDevice(dst, static_cast<const char*>(scalars) + lo * kScalarBytes,
                     len * kScalarBytes, scalars_pageable);
        d_scalars = reinterpret_cast<const uint32_t*>(dst);
      }
      if (bases_dev) {
        d_bases = reinterpret_cast<const uint32_t*>(static_cast<const char*>(bases) +
                                                    lo * kAffineBytes);
      } else {
        char* dst = bases_stage_.as<char>() + stage * bases_slot_bytes;
        CopyToDevice(dst, static_cast<const char*>(bases) + lo * kAffineBytes, len * kAffineBytes,
                     bases_pageable);
        d_bases = reinterpret_cast<const uint32_t*>(dst);
      }
      if (pd.any_host) {
        TB_CUDA(cudaEventRecord(ev(r, 0), copy_stream_));
        TB_CUDA(cudaStreamWaitEvent(stream_, ev(r, 0), 0));
      }
      TB_CUDA(cudaEventRecord(ev(r, 1), stream_));

      // ---- sort: histogram, scan, tasks, scatter ----------------------------------
      TB_CUDA(cudaMemsetAsync(count_.ptr, 0, (size_t)(plan.TB + 1) * 4, stream_));
      TB_CUDA(cudaMemsetAsync(nonzero_slots_.ptr, 0, (size_t)kNonzeroSlots * 4, stream_));
      uint32_t sgrid = (plan.n + 255) / 256;
      const bool two_level = UseTwoLevelSort(plan);
      SortPlan sp = MakeSortPlan(plan);
      // coarse_: [count | offset | cursor | tile_offset], each regions + 1 words, then totals
      uint32_t* coarse_count = coarse_.as<uint32_t>();
      uint32_t* coarse_offset = coarse_count + (sp.regions + 1);
      uint32_t* coarse_cursor = coarse_offset + (sp.regions + 1);
      uint32_t* tile_offset = coarse_cursor + (sp.regions + 1);
      SortTotals* sort_totals = reinterpret_cast<SortTotals*>(tile_offset + (sp.regions + 1));
      uint32_t tgrid = (plan.n + kSortTile - 1) / kSortTile;
      if (two_level) {
        TB_CUDA(cudaMemsetAsync(coarse_count, 0, (size_t)(sp.regions + 1) * 4, stream_));
        size_t smem = (size_t)sp.regions * 4;
        if (smem > sort_smem_set_) {
          TB_CUDA(cudaFuncSetAttribute(digits_coarse_hist_kernel<C>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          sort_smem_set_ = smem;
        }
        digits_coarse_hist_kernel<C><<<tgrid, kSortThreads, smem, stream_>>>(
            d_scalars, plan, sp, digits_.as<uint32_t>(), coarse_count);
        TB_CUDA(cudaGetLastError());
        ++launches_;
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
        Launch(coarse_scan_kernel, 1, 1024, coarse_count, sp.regions, coarse_offset, coarse_cursor,
               tile_offset, sort_totals, nonzero_slots_.as<uint32_t>());
        LaunchGrid(coarse_scatter_kernel, dim3(tgrid, plan.W), kSortThreads, digits_.as<uint32_t>(),
                   plan, sp, coarse_cursor, mid_.as<uint2>());
        Launch(fine_hist_kernel, sp.max_tiles, kSortThreads, mid_.as<uint2>(), sp, tile_offset,
               coarse_offset, sort_totals, count_.as<uint32_t>());
      } else {
        Launch(digits_hist_kernel<C>, sgrid, 256, d_scalars, plan, digits_.as<uint32_t>(),
               count_.as<uint32_t>(), nonzero_slots_.as<uint32_t>());
      }
      // mean run of the most loaded windows: the narrow (c - 1 bit) windows of a balanced plan
      // spread the same n entries over half as many buckets
      const uint32_t seg_buckets = plan.shared ? plan.B : plan.W * (plan.wide < plan.W ? plan.B / 2 : plan.B);
      Launch(choose_segment_kernel, 1, 64, nonzero_slots_.as<uint32_t>(), seg_buckets, plan.R,
             plan.seg, options_.segment ? plan.seg : 0u, totals_);
      Launch(scan_block_sums_kernel, scan_blocks, kScanThreads, count_.as<uint32_t>(), plan.TB,
             totals_, plan.R, block_sums_.as<uint64_t>());
      Launch(scan_top_kernel, 1, kScanThreads, block_sums_.as<uint64_t>(), scan_blocks, totals_);
      Launch(scan_apply_build_tasks_kernel, scan_blocks, kScanThreads, count_.as<uint32_t>(),
             plan.TB, plan.R, block_sums_.as<uint64_t>(), offset_.as<uint32_t>(),
             cursor_.as<uint32_t>(), task_base_.as<uint32_t>(), tasks_.as<uint2>(),
             task_meta_.as<uint32_t>(), fold_jobs_.as<uint2>(), multi_.as<uint32_t>(),
             sorted_.as<uint32_t>(), totals_);
      if (two_level)
        Launch(fine_scatter_kernel, sp.max_tiles, kSortThreads, mid_.as<uint2>(), sp, tile_offset,
               coarse_offset, sort_totals, cursor_.as<uint32_t>(), sorted_.as<uint32_t>());
      else
        LaunchGrid(digits_scatter_kernel, dim3(sgrid, plan.W), 256, digits_.as<uint32_t>(), plan,
                   cursor_.as<uint32_t>(), sorted_.as<uint32_t>());
      // tasks by window group (last range only: high windows first), then descending length
      const bool split = pd.low > 0 && r == last_range;
      const uint32_t split_key = split ? pd.low * plan.B : 0u;
      TB_CUDA(cudaMemsetAsync(len_hist_.ptr, 0, (size_t)kOrderBins * 4, stream_));
      uint32_t ogrid = (plan.max_tasks + kOrderThreads * kOrderPerThread - 1) /
                       (kOrderThreads * kOrderPerThread);
      Launch(order_hist_kernel, ogrid, kOrderThreads, tasks_.as<uint2>(), task_meta_.as<uint32_t>(),
             totals_, split_key, len_hist_.as<uint32_t>());
      Launch(order_scan_kernel, 1, 1024, len_hist_.as<uint32_t>(), offset_.as<uint32_t>(), split_key,
             totals_);
      Launch(order_scatter_kernel, ogrid, kOrderThreads, tasks_.as<uint2>(),
             task_meta_.as<uint32_t>(), totals_, split_key, len_hist_.as<uint32_t>(),
             order_.as<uint32_t>());
      TB_CUDA(cudaEventRecord(ev(r, 2), stream_));

      // ---- batched-affine pair rounds, then XYZZ accumulation into the bucket values ----
      for (uint32_t pr = 0; pr < plan.R; ++pr) {
        uint32_t grid = PairThreads(plan, pr) / kPairThreads;
        if (pr == 0)
          Launch(pair_round_kernel<C, true>, grid, kPairThreads, d_bases, sorted_.as<uint32_t>(),
                 totals_, pr, pair_prefix_.as<uint32_t>(), pair_out_[0].as<uint32_t>());
        else
          Launch(pair_round_kernel<C, false>, grid, kPairThreads, (const uint32_t*)nullptr,
                 pair_out_[pr - 1].as<uint32_t>(), totals_, pr, pair_prefix_.as<uint32_t>(),
                 pair_out_[pr].as<uint32_t>());
      }
      auto accumulate = [&](uint32_t part, uint32_t max_tasks) {
        uint32_t agrid = (max_tasks + kAccThreads - 1) / kAccThreads;
        if (agrid == 0) agrid = 1;
        if (plan.R)
          Launch(accumulate_kernel<C, true>, agrid, kAccThreads, pair_out_[plan.R - 1].as<uint32_t>(),
                 (const uint32_t*)nullptr, tasks_.as<uint2>(), task_meta_.as<uint32_t>(),
                 order_.as<uint32_t>(), totals_, part, state_.as<uint32_t>(), task_out_.as<uint32_t>());
        else if (options_.stage_points)
          Launch(accumulate_staged_kernel<C>, agrid, kAccThreads, d_bases, sorted_.as<uint32_t>(),
                 tasks_.as<uint2>(), task_meta_.as<uint32_t>(), order_.as<uint32_t>(), totals_, part,
                 state_.as<uint32_t>(), task_out_.as<uint32_t>());
        else if (C::Field::kDegree == 2 && options_.acc_variant != 0)
          LaunchPairAccumulate(max_tasks, d_bases, part);
        else if (C::Field::kDegree == 1 && (uint64_t)plan.n * plan.W >= kAccSmallEntries &&
                 (options_.acc_variant == 3 || (options_.acc_variant < 0 && C::kAccLockstep)))
          LaunchLockstep(agrid, d_bases, part);
        else
          Launch((uint64_t)plan.n * plan.W < kAccSmallEntries
                     ? accumulate_kernel<C, false, AccMinBlocksSmall<C>()>
                     : accumulate_kernel<C, false>,
                 agrid, kAccThreads, d_bases, sorted_.as<uint32_t>(), tasks_.as<uint2>(),
                 task_meta_.as<uint32_t>(), order_.as<uint32_t>(), totals_, part,
                 state_.as<uint32_t>(), task_out_.as<uint32_t>());
      };
      auto fold = [&](uint32_t part) {
        Launch(fold_stage_a_kernel<C>, sm_count_ * 8, kFoldThreads, fold_jobs_.as<uint2>(), totals_,
               offset_.as<uint32_t>(), task_base_.as<uint32_t>(), plan.R, part, split_key,
               task_out_.as<uint32_t>(), state_.as<uint32_t>());
        Launch(fold_stage_b_kernel<C>, sm_count_, kFoldThreads, multi_.as<uint32_t>(), totals_,
               offset_.as<uint32_t>(), task_base_.as<uint32_t>(), plan.R, part, split_key,
               task_out_.as<uint32_t>(), state_.as<uint32_t>());
      };
      if (!split) {
        accumulate(kPartAll, plan.max_tasks);
        TB_CUDA(cudaEventRecord(ev(r, 4), stream_));
        fold(kPartAll);
      } else {
        // high windows first; their reduction and window combination then run on the tail
        // stream while this stream accumulates the low windows
        accumulate(kPartHigh, plan.max_tasks);
        TB_CUDA(cudaEventRecord(ev(r, 4), stream_));
        fold(kPartHigh);
        TB_CUDA(cudaEventRecord(SlotEvent(slot, 4), stream_));
        // the running-sum level of the high windows is throughput-bound like the accumulation:
        // it stays on this stream, ahead of the low windows (on a second stream it shares the
        // SMs with them and ends as late as they do: 2^21 points 1.02 ms instead of 0.48); only
        // the latency-bound merge tree and doubling chain move to the tail stream
        uint32_t* terms_hi = combine_.as<uint32_t>();
        EnqueueReduction(tail_stream_, red, pd.low, red.W - pd.low, pd.L0, tree_hi_, terms_hi, stream_,
                         SlotEvent(slot, 10));
        TB_CUDA(cudaEventRecord(SlotEvent(slot, 6), tail_stream_));
        LaunchOn(tail_stream_, window_combine_kernel<C>, 1, kCombineThreads, terms_hi,
                 TotalBits(red), WindowBitOffset(red, pd.low), (const uint32_t*)nullptr, HiSum(),
                 PerWindow{0, 0, 0});
        TB_CUDA(cudaEventRecord(SlotEvent(slot, 7), tail_stream_));
        TB_CUDA(cudaEventRecord(SlotEvent(slot, 5), tail_stream_));
        // low group: at most low * B buckets, each non-empty one a task, plus the split ones
        uint64_t lo_entries = (uint64_t)plan.n * pd.low;
        uint64_t lo_buckets = (uint64_t)pd.low * plan.B;
        uint32_t seg_floor = options_.segment ? plan.seg : kMinSegment;
        accumulate(kPartLow, (uint32_t)((lo_entries < lo_buckets ? lo_entries : lo_buckets) +
                                        lo_entries / seg_floor));
        fold(kPartLow);
      }
      TB_CUDA(cudaMemcpyAsync(host_out + kHostPartialBytes + r * sizeof(MsmTotals), totals_,
                              sizeof(MsmTotals), cudaMemcpyDeviceToHost, stream_));
      TB_CUDA(cudaEventRecord(ev(r, 3), stream_));
      if (pd.any_host) {
        TB_CUDA(cudaEventRecord(StageFreeEvent(stage), stream_));
        stage_used_[stage] = true;
      }
    }
    TB_CUDA(cudaEventRecord(ev_acc_end, stream_));

    // ---- bucket reduction (running-sum level + fused merge tree) and window combination ----
    // With a window split the high group is already on its way on the tail stream; this stream
    // finishes the low windows and adds the high group's sum.  The result is ONE XYZZ point.
    if (pd.low > 0) {
      uint32_t* terms_lo = combine_.as<uint32_t>() + (size_t)kTermSlots * kXyzzWords;
      EnqueueReduction(stream_, red, 0, pd.low, pd.L0_low, tree_lo_, terms_lo);
      TB_CUDA(cudaEventRecord(SlotEvent(slot, 8), stream_));
      TB_CUDA(cudaStreamWaitEvent(stream_, SlotEvent(slot, 5), 0));
      TB_CUDA(cudaEventRecord(SlotEvent(slot, 9), stream_));
      Launch(window_combine_kernel<C>, 1, kCombineThreads, terms_lo, WindowBitOffset(red, pd.low), 0u,
             (const uint32_t*)HiSum(), Partial(), PerWindow{0, 0, 0});
    } else {
      uint32_t* terms = combine_.as<uint32_t>();
      EnqueueReduction(stream_, red, 0, red.W, pd.L0, tree_hi_, terms);
      TB_CUDA(cudaEventRecord(SlotEvent(slot, 6), stream_));
      if (pd.device_ladder)
        Launch(window_combine_kernel<C>, 1, kCombineThreads, terms, TotalBits(red), 0u,
               (const uint32_t*)nullptr, Partial(), PerWindow{0, 0, 0});
      else  // the parallel part only: one CTA per window, S_w to WindowSums()
        Launch(window_combine_kernel<C>, red.W, kCombineThreads, terms, 0u, 0u, (const uint32_t*)nullptr,
               WindowSums(), PerWindow{red.c, red.wide, 0});
      TB_CUDA(cudaEventRecord(SlotEvent(slot, 7), stream_));
    }
    pd.gathered = gather;
    if (gather) {
      EnqueueGather(Partial());
    } else if (pd.device_ladder) {
      TB_CUDA(cudaMemcpyAsync(host_out, Partial(), kXyzzBytes, cudaMemcpyDeviceToHost, stream_));
    } else {
      TB_CUDA(cudaMemcpyAsync(host_out, WindowSums(), (size_t)red.W * kXyzzBytes, cudaMemcpyDeviceToHost,
                              stream_));
    }
    TB_CUDA(cudaEventRecord(ev_end, stream_));
    pd.launches = launches_;
    if (trace) fprintf(stderr, "[tachyon_b200] enqueue n=%zu: everything queued at %.2f ms\n", n, since());
    timing_.enqueue_ms +=
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return pd;
  }

  static bool bound_nonempty(const MsmTotals& t) { return t.tasks != 0; }

  Point Finish(const Pending& pd) {
    const MsmPlan& plan = pd.plan;
    const int slot = pd.slot;
    auto ev = [&](size_t r, int which) { return SlotEvent(slot, kSlotEvents + kRangeEvents * r + which); };
    auto wait0 = std::chrono::steady_clock::now();
    TB_CUDA(cudaEventSynchronize(SlotEvent(slot, 2)));
    timing_.wait_ms +=
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wait0).count();
    const char* host_out = host_out_ + (size_t)slot * kHostOutBytes;

    // ---- host epilogue: none.  The device left ONE point — bucket reduction, merge tree and
    // window combination (pippenger_base.h:59-77) all ran as kernels. ------------------------
    auto host0 = std::chrono::steady_clock::now();
    Point result;
    if (pd.gathered) {
      result = SumGathered();
    } else if (pd.device_ladder) {
      memcpy(&result, host_out, sizeof(result));
    } else {
      // The ladder of pippenger_base.h:59-77 over the W window sums the device produced: from the
      // top window down, c_w doublings then one addition.  Strictly sequential work on one point.
      const Point* sums = reinterpret_cast<const Point*>(host_out);
      result = Point::Zero();
      const MsmPlan& red = pd.red;  // one bucket set with the precomputed table: no ladder at all
      for (uint32_t w = red.W; w-- > 0;) {
        const uint32_t cw = red.c - (w >= red.wide ? 1u : 0u);
        for (uint32_t k = 0; k < cw && w + 1 < red.W; ++k) result = result.Dbl();
        result = result.Add(sums[w]);
      }
    }
    auto host1 = std::chrono::steady_clock::now();

    float ms;
    for (size_t r = 0; r < pd.K; ++r) {
      MsmTotals tot;
      memcpy(&tot, host_out + kHostPartialBytes + r * sizeof(MsmTotals), sizeof(tot));
      timing_.tasks += tot.tasks;
      timing_.entries += tot.entries;
      TB_CUDA(cudaEventElapsedTime(&ms, ev(r, 1), ev(r, 2)));
      timing_.sort_ms += ms;
      TB_CUDA(cudaEventElapsedTime(&ms, ev(r, 2), ev(r, 3)));
      timing_.accumulate_ms += ms;
      if (bound_nonempty(tot)) {
        // the range's first accumulate launch ran alone on the device
        TB_CUDA(cudaEventElapsedTime(&ms, ev(r, 2), ev(r, 4)));
        timing_.acc_kernel_ms += ms;
        timing_.acc_kernel_entries += tot.entries - tot.entries_lo;  // entries_lo: 0 unless split
      }
    }
    TB_CUDA(cudaEventElapsedTime(&ms, SlotEvent(slot, 6), SlotEvent(slot, 7)));
    timing_.combine_ms += ms;
    timing_.low_windows = pd.low;
    if (pd.low > 0) {
      static const bool trace = getenv("TACHYON_B200_TRACE") != nullptr;
      if (trace) {
        float hi_level, hi_total, lo_tree, lo_wait, lo_rest;
        TB_CUDA(cudaEventElapsedTime(&hi_level, SlotEvent(slot, 4), SlotEvent(slot, 10)));
        TB_CUDA(cudaEventElapsedTime(&hi_total, SlotEvent(slot, 4), SlotEvent(slot, 5)));
        TB_CUDA(cudaEventElapsedTime(&lo_tree, SlotEvent(slot, 1), SlotEvent(slot, 8)));
        TB_CUDA(cudaEventElapsedTime(&lo_wait, SlotEvent(slot, 8), SlotEvent(slot, 9)));
        TB_CUDA(cudaEventElapsedTime(&lo_rest, SlotEvent(slot, 9), SlotEvent(slot, 2)));
        float acc_lo;
        TB_CUDA(cudaEventElapsedTime(&acc_lo, SlotEvent(slot, 4), SlotEvent(slot, 1)));
        fprintf(stderr,
                "[tachyon_b200] window groups: low=%u | after the high group's accumulation: high level "
                "%.3f, high level+tree+combine %.3f, low accumulate %.3f | after the low group's: level+tree "
                "%.3f, wait for high %.3f, combine+copy %.3f ms\n",
                pd.low, hi_level, hi_total, acc_lo, lo_tree, lo_wait, lo_rest);
      }
    }
    if (pd.any_host) {
      // time the copy engine was busy or waiting for a free slot; overlaps the bucket work
      TB_CUDA(cudaEventElapsedTime(&ms, SlotEvent(slot, 3), ev(pd.K - 1, 0)));
      timing_.h2d_ms += ms;
      if (pd.host_bytes >= (size_t(32) << 20) && ms > 0)  // the copy rate this call saw
        h2d_gbs_ = (double)pd.host_bytes / (ms * 1e-3) / 1e9;
    }
    TB_CUDA(cudaEventElapsedTime(&ms, SlotEvent(slot, 1), SlotEvent(slot, 2)));
    timing_.reduce_ms += ms;
    TB_CUDA(cudaEventElapsedTime(&ms, SlotEvent(slot, 0), SlotEvent(slot, 2)));
    timing_.total_ms += ms;
    timing_.host_ms += std::chrono::duration<float, std::milli>(host1 - host0).count();
    timing_.window_bits = plan.c;
    timing_.windows = plan.W;
    timing_.ranges += (uint32_t)pd.K;
    timing_.pair_rounds = plan.R;
    timing_.kernel_launches += pd.launches;
    return result;
  }

  // All-gather of the ranks' partials (device pointer `partial`, one XYZZ point) on the
  // compute stream, then one copy of the `world` points to pinned host memory.
  void EnqueueGather(const void* partial) {
    const NcclApi* nccl = NcclApi::Get(nullptr);
    int rc = nccl->AllGather(partial, gather_dev_, kXyzzBytes, NcclApi::kUint8, comm_, stream_);
    if (rc != 0) throw CudaError{cudaErrorUnknown, nccl->GetErrorString(rc), __FILE__, __LINE__};
    TB_CUDA(cudaMemcpyAsync(gather_host_, gather_dev_, (size_t)world_ * kXyzzBytes,
                            cudaMemcpyDeviceToHost, stream_));
  }
  Point SumGathered() const {  // pippenger_adapter.h:110-113
    Point total;
    for (int g = 0; g < world_; ++g) {
      Point part;
      memcpy(&part, gather_host_ + (size_t)g * kXyzzBytes, sizeof(part));
      total = g == 0 ? part : total.Add(part);
    }
    return total;
  }
  // The rare paths (empty input, several pieces): the local sum is already on the host.
  Point GatherHostPoint(const Point& local) {
    char* up = gather_host_ + (size_t)world_ * kXyzzBytes;
    memcpy(up, &local, sizeof(local));
    TB_CUDA(cudaMemcpyAsync(Scratch(), up, kXyzzBytes, cudaMemcpyHostToDevice, stream_));
    EnqueueGather(Scratch());
    TB_CUDA(cudaStreamSynchronize(stream_));
    return SumGathered();
  }
  uint32_t* Scratch() {
    if (!combine_.ptr) combine_.Reserve((size_t)(2 * kTermSlots + 2 + kMaxWindows) * kXyzzBytes);
    return Partial();
  }

  // What the bucket reduction and the window combination see: W bucket sets, or — with the
  // precomputed table — one set whose sum already is the MSM (no window weights left).
  static MsmPlan ReductionPlan(MsmPlan p) {
    if (p.shared) {
      p.W = 1;
      p.wide = 1;
    }
    return p;
  }

  // Bit position of window w's lowest bit / of the end of the top window (balanced windows).
  static uint32_t WindowBitOffset(const MsmPlan& p, uint32_t w) {
    return w * p.c - (w > p.wide ? w - p.wide : 0u);
  }
  static uint32_t TotalBits(const MsmPlan& p) { return WindowBitOffset(p, p.W); }
  uint32_t* HiSum() { return combine_.as<uint32_t>() + (size_t)2 * kTermSlots * kXyzzWords; }
  uint32_t* Partial() { return HiSum() + kXyzzWords; }
  uint32_t* WindowSums() { return Partial() + kXyzzWords; }  // kMaxWindows points

  // Bucket reduction of windows [w0, w0 + wn) on stream `st`: one blocked running-sum level,
  // then the merge tree in stages of <= kTreeStageLevels levels per launch; the last stage
  // scatters every window's (A, P, D_j) to the bit positions of `terms`.
  // `level_stream` runs the (throughput-bound) running-sum level, `st` the (latency-bound) tree;
  // when they differ, `after_level` orders the two.
  void EnqueueReduction(cudaStream_t st, const MsmPlan& plan, uint32_t w0, uint32_t wn, uint32_t L0,
                        TreeBuffers& set, uint32_t* terms, cudaStream_t level_stream = nullptr,
                        cudaEvent_t after_level = nullptr) {
    if (!level_stream) level_stream = st;
    const uint32_t nb = plan.B / L0, M = Log2(nb), l0 = Log2(L0);
    const size_t slice_words = (size_t)nb * 2 * kXyzzWords;
    uint32_t* leaves = set.leaves.template as<uint32_t>() + (size_t)w0 * slice_words;
    const uint32_t* bucket0 = state_.as<uint32_t>() + (size_t)w0 * plan.B * kXyzzWords;
    const uint32_t wide_local = plan.wide > w0 ? plan.wide - w0 : 0u;
    const uint32_t blocks = wn * nb;
    if (options_.reduce_mode == 0) {
      LaunchOn(level_stream, reduce_level_kernel<C, true>, (blocks + kReduceThreads - 1) / kReduceThreads,
               kReduceThreads, bucket0, (const uint32_t*)nullptr, plan.B, nb, L0, 0u, wn, leaves,
               leaves + kXyzzWords);
    } else if (C::Field::kDegree == 2 && options_.reduce_mode == 1) {
      LaunchPairReduce(level_stream, blocks, bucket0, plan.B, nb, L0, wn, wide_local, leaves);
    } else {
      constexpr uint32_t kSlots = ReduceSlots<C>();
      using K0 = typename C::Field;
      const int roll = options_.reduce_roll < 0 ? C::kReduceRoll : options_.reduce_roll;
      auto kernel = roll == 2   ? reduce_blocks_kernel<C, typename K0::template WithRoll<2>>
                    : roll == 1 ? reduce_blocks_kernel<C, typename K0::template WithRoll<1>>
                                : reduce_blocks_kernel<C, K0>;
      if constexpr (C::Field::kDegree == 1) {  // one inlined call site of the addition
        const bool inl = options_.reduce_inline < 0 ? C::kReduceInline : options_.reduce_inline != 0;
        if (inl && roll == 0) kernel = reduce_blocks_kernel<C, K0, true>;
        if (inl && roll == 1) kernel = reduce_blocks_kernel<C, typename K0::template WithRoll<1>, true>;
      }
      LaunchOn(level_stream, kernel, (blocks + kSlots - 1) / kSlots, 2 * kSlots, bucket0,
               plan.B, nb, L0, wn, wide_local, leaves, leaves + kXyzzWords);
    }
    if (after_level) {
      TB_CUDA(cudaEventRecord(after_level, level_stream));
      if (level_stream != st) TB_CUDA(cudaStreamWaitEvent(st, after_level, 0));
    }
    // Stages of the merge tree.  The LAST stage is one CTA per window, so it is kept to
    // kTreeLastLevels levels (its first level has 2^(levels-1) x (values per node) additions for
    // 64 lane groups); the levels below it are cut into equal stages of <= kTreeStageLevels.
    const uint32_t last_levels = M < kTreeLastLevels ? M : kTreeLastLevels;
    const uint32_t lower = M - last_levels;
    const uint32_t lower_stages = (lower + kTreeStageLevels - 1) / kTreeStageLevels;
    const uint32_t* in = leaves;
    uint32_t vin = 2, nodes = nb, remaining = M, stage = 0;
    do {
      uint32_t levels = last_levels;
      if (stage < lower_stages) {  // spread `lower` levels over the lower stages, larger ones first
        const uint32_t left = remaining - last_levels, stages_left = lower_stages - stage;
        levels = (left + stages_left - 1) / stages_left;
      }
      const bool last = levels == remaining;
      const uint32_t ctas = nodes >> levels;
      uint32_t* out = set.out[stage & 1].template as<uint32_t>() + (size_t)w0 * slice_words;
      TreeFinal fin{last ? 1u : 0u, w0, plan.c, plan.wide, l0, terms};
      // more than two CTAs per resident slot: the stage is throughput-bound, its wide levels run
      // one thread per addition; otherwise every level uses the four-lane latency form
      const uint32_t thread_items =
          (uint64_t)wn * ctas > (uint64_t)4 * sm_count_ ? (uint32_t)kTreeThreads / 2 : 0xffffffffu;
      LaunchOn(st, reduce_tree_kernel<C>, wn * ctas, (uint32_t)kTreeThreads, in, vin, levels, ctas,
               slice_words, set.ping.template as<uint32_t>() + (size_t)w0 * slice_words,
               set.pong.template as<uint32_t>() + (size_t)w0 * slice_words, out, fin, thread_items);
      in = out;
      vin += levels;
      nodes >>= levels;
      remaining -= levels;
      ++stage;
    } while (remaining > 0);
  }

  // Running-sum block length of the low group: a block costs L0 sequential additions (~7 us
  // each when latency-bound), every halving adds a tree level (~3 us) and doubles the tree's
  // work; 4 up to 2^18 bucket slots, 8 above.
  static uint32_t LowLevelLength(const MsmPlan& p, uint32_t low) {
    uint32_t L = (uint64_t)low * p.B <= (1u << 18) ? 4u : 8u;
    while (L > 2 && L > p.B) L >>= 1;
    return L;
  }

  // Nanoseconds per mixed addition of the accumulation kernel at full occupancy (measured, B200).
  static constexpr double MaddNanos() {
    // (BLS12-381 G1 with its warps in step: 0.313; G2 on lane pairs: 0.475 / 1.08)
    return C::Field::kWords <= 8 ? 0.138 : (C::Field::kWords <= 12 ? 0.313 : (C::Field::kWords <= 16 ? 0.475 : 1.08));
  }

  // How many low windows to accumulate last.  The window combination of the high group is a
  // chain of ~(bits above the split) point doublings, latency-bound on one lane; it is free as
  // long as the low group's accumulation (n mixed additions per window at the pipe rate) lasts
  // longer.  What remains exposed after the low group is its own reduction plus the doublings
  // below the split.  Picks the split with the smallest exposed time; 0 = no split.
  uint32_t ChooseLowWindows(const MsmPlan& p, size_t n_last) const {
    if (options_.low_windows >= 0)
      return (uint32_t)options_.low_windows < p.W ? (uint32_t)options_.low_windows : p.W - 1;
    if (p.R) return 0;  // pair rounds: experimental path, single group
    // Below ~2^24 entries the accumulation is latency-bound (it ends when its longest task ends,
    // whatever the number of windows), so a second accumulate launch adds its own critical path
    // instead of hiding anything: measured, 2^18 points 2.33 -> 2.62 ms, 2^16 1.08 -> 1.44 ms.
    if ((uint64_t)n_last * p.W < (uint64_t(1) << 24)) return 0;
    // measured on B200: microseconds per doubling of the combine chain (four-lane form,
    // tools/probe/chain_probe.cu), nanoseconds per mixed addition of the accumulation kernel at
    // full occupancy
    constexpr int kW = C::Field::kWords;
    const double dbl_us = kW <= 8 ? 1.64 : (kW <= 12 ? 3.28 : (kW <= 16 ? 4.41 : 9.6));
    const double madd_ns = MaddNanos();
    // timeline after the high group's accumulation and running-sum level (t = 0), tail stream:
    // merge tree (latency), doubling chain; compute stream: accumulation of the low windows, then
    // their level + tree, then their chain, which needs the high group's sum.  Minimise what is
    // NOT accumulation work.
    const double kTreeUs = 130.0;     // merge tree of a window group (latency-bound)
    const double kLowFixedUs = 110.0; // low group: running-sum level + tree
    const double kSplitUs = 30.0;     // second accumulate launch and its partial last wave
    const uint32_t total_bits = TotalBits(p);
    auto chain = [&](uint32_t bits) { return bits ? dbl_us * (bits + 8) + 20.0 : 0.0; };
    double best = kTreeUs + chain(total_bits);  // no split: everything after the level is exposed
    uint32_t best_low = 0;
    for (uint32_t low = 1; low < p.W; ++low) {
      const uint32_t below = WindowBitOffset(p, low);
      const double acc_low = (double)low * (double)n_last * madd_ns * 1e-3;
      const double hi_ready = kTreeUs + chain(total_bits - below);
      const double lo_ready = acc_low + kLowFixedUs;
      const double exposed = (hi_ready > lo_ready ? hi_ready : lo_ready) + chain(below) - acc_low + kSplitUs;
      if (exposed < best) {
        best = exposed;
        best_low = low;
      }
    }
    return best_low;
  }

  // Threads of pair round r: kPairBatch pairs each, but never fewer than a full grid.
  uint32_t PairThreads(const MsmPlan& p, uint32_t r) const {
    uint64_t pairs = (PaddedBound(p) >> r) >> 1;
    uint64_t t = (pairs + kPairBatch - 1) / kPairBatch;
    uint64_t full = (uint64_t)sm_count_ * AccMinBlocks<C>() * kPairThreads;
    if (t < full) t = full;
    if (t > pairs) t = pairs ? pairs : 1;
    return (uint32_t)((t + kPairThreads - 1) / kPairThreads * kPairThreads);
  }

  size_t OwnedBytes() const {
    size_t b = 0;
    for (const DeviceBuffer* d : AllBuffers()) b += d->bytes;
    return b;
  }
  std::vector<const DeviceBuffer*> AllBuffers() const {
    return {&registered_, &bases_stage_, &scalars_stage_, &state_, &count_, &offset_, &cursor_, &task_base_,
            &tasks_, &task_meta_, &multi_, &sorted_, &digits_, &task_out_, &block_sums_, &order_,
            &len_hist_, &tree_hi_.leaves, &tree_hi_.ping, &tree_hi_.pong, &tree_hi_.out[0],
            &tree_hi_.out[1], &tree_lo_.leaves, &tree_lo_.ping, &tree_lo_.pong, &tree_lo_.out[0],
            &tree_lo_.out[1], &combine_,
            &pair_prefix_, &pair_out_[0], &pair_out_[1], &pair_out_[2], &pair_out_[3], &mid_,
            &coarse_, &fold_jobs_, &nonzero_slots_};
  }

  static uint32_t Log2(uint32_t x) {
    uint32_t r = 0;
    while ((1u << r) < x) ++r;
    return r;
  }

  // Blocks of the running-sum level: as long as possible while the grid still fills the chip
  // (a block costs 2 L full additions, the merge tree ~3 per block).  Measured sweep of the
  // blocks wanted per SM (option "level_fill"); with the round-1 kernels 384 was best up to ~1.5 M
  // bucket slots and 768 above (2^24 points 3.34 -> 3.25 ms).
  uint32_t ChooseLevelLength(uint32_t buckets_per_window, uint32_t windows) const {
    uint64_t items = (uint64_t)buckets_per_window * windows;
    // Re-measured with the final running-sum kernels (profiles/r2_zz_level_fill_sweep.txt): 384 blocks
    // per SM everywhere (the 768 of the earlier kernels now loses: BN254 2^23 1.36 -> 1.26 ms,
    // BLS12-381 2^22 2.16 -> 1.87 ms), 96 for the 8-limb curve below ~400 K bucket slots (c <= 15:
    // 0.31 -> 0.27 ms), 48 for G2, whose lane-pair kernel holds 96 / 64 blocks per SM at a time
    // (BN254 G2 2^20 0.96 -> 0.85 ms, 2^16 1.06 -> 0.90 ms).
    uint32_t fill = options_.level_fill ? options_.level_fill
                    : (C::Field::kDegree == 2 ? 48u : (C::Field::kWords == 8 && items < 400000 ? 96u : 384u));
    uint64_t want_blocks = (uint64_t)sm_count_ * fill;
    uint32_t L = 64;
    while (L > 4 && items / L < want_blocks) L >>= 1;
    if (L > buckets_per_window) L = buckets_per_window;
    return L;
  }

  // pinned result buffer per Pending slot: the MSM's point, then one MsmTotals per range
  static constexpr size_t kHostPartialBytes = (kMaxWindows * kXyzzBytes + 255) / 256 * 256;
  static constexpr size_t kHostOutBytes = kHostPartialBytes + kMaxRanges * sizeof(MsmTotals);
  // bit positions of the window-combination term arrays: W * c < 256 + c
  static constexpr uint32_t kTermSlots = 320;

  int device_;
  int sm_count_ = 148;
  cudaStream_t own_stream_ = nullptr;
  cudaStream_t stream_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  cudaStream_t sample_stream_ = nullptr;
  cudaStream_t tail_stream_ = nullptr;
  std::vector<cudaEvent_t> events_;
  size_t stage_seq_ = 0;
  bool in_batch_tail_ = false;
  bool skew_hint_ = false;  // the last sample of the scalars was skewed (WindowBitsFromSample)
  char* bounce_ = nullptr;
  size_t bounce_seq_ = 0;
  bool bounce_used_[kBounceSlots] = {};
  std::unique_ptr<ParallelMemcpy> copier_;
  size_t budget_ = 0;  // device bytes this engine may use; 0 = ask the driver
  double h2d_gbs_ = 0;  // host-to-device rate of the last large host-input call (0: none yet)
  bool stage_used_[kStageSlots] = {};
  MsmOptions options_;
  MsmTiming timing_;
  uint32_t launches_ = 0;
  MsmTotals* totals_ = nullptr;
  void* arena_ = nullptr;  // backing store of the workspace buffers (ReserveAll)
  char* host_out_ = nullptr;
  size_t registered_n_ = 0;
  uint32_t table_c_ = 0, table_wide_ = 0;  // window size of the precomputed table (0: none)
  DeviceBuffer registered_, bases_stage_, scalars_stage_, state_, count_, offset_, cursor_, task_base_, tasks_,
      task_meta_, multi_, sorted_, digits_, task_out_, block_sums_, order_, len_hist_, combine_,
      pair_prefix_, pair_out_[4], mid_, coarse_,
      fold_jobs_, nonzero_slots_;
  TreeBuffers tree_hi_, tree_lo_;
  size_t sort_smem_set_ = 0;
  // multi-process point-range sharding (JoinRanks)
  NcclApi::Comm comm_ = nullptr;
  int rank_ = 0, world_ = 1;
  char* gather_dev_ = nullptr;
  char* gather_host_ = nullptr;
};

}  // namespace tb200
