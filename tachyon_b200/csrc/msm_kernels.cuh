// sm_100a kernels of the variable-base MSM pipeline (bucket method).
//
// This is the from-scratch replacement of the single external call the
// reference makes at
//   tachyon/math/elliptic_curves/msm/algorithms/icicle/icicle_msm_bn254_g1.cc:12-17
// (and the bls12_381 twin), and computes the same group element as the CPU
// path tachyon/math/elliptic_curves/msm/algorithms/pippenger/pippenger.h:68-169.
//
//   digits_hist     scalar de-Montgomery + signed-window recode -> bucket sizes
//   scan_*          exclusive scan of sizes (bucket offsets) and task counts
//   build_tasks     bucket -> fixed-max-length segments ("tasks")
//   digits_scatter  counting-sort scatter of (point index | sign) by bucket
//   accumulate      per task: sum of +-P over its segment, XYZZ mixed adds  (HOT)
//   fold_partials   buckets that were split into several tasks
//   reduce_level    sum_k (k+1) B_k per window, blocked running sums
//
// Layouts (all in HBM, SoA where a kernel streams, AoS where it gathers):
//   bases    n x {x, y}           Montgomery u32 limbs, 64 B (BN254) / 96 B (BLS)
//   scalars  n x 8 u32            Montgomery Fr
//   count    TB+1 u32             TB = W * 2^(c-1) buckets, key = w * 2^(c-1) + |d| - 1
//   sorted   <= n*W u32           point index | sign << 31, grouped by key
//   tasks    uint2 {start, len}   len <= seg
//   task_out T x XYZZ             128 B / 192 B
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xyzz.cuh"

namespace tb200 {

struct Bn254Curve {
  using Fq = Bn254FqParams;
  using Field = FpField<Bn254FqParams>;  // coordinate field of the group
  using Fr = Bn254FrParams;
  using Gen = Bn254G1Generator;
  static constexpr const char* kName = "bn254";
  // running-sum kernel: both roles share ONE inlined call site of the addition (2^24: reduce
  // 3.45 -> 2.75 ms; every product a fused IMAD.WIDE, no operands through the stack)
  static constexpr bool kReduceInline = true;
  static constexpr int kReduceRoll = 0;  // running-sum kernel: unrolled multiplications (looped: 3.38 -> 3.67 ms at 2^24)
  // accumulation warps free-running: the 68 KB loop body streams through the instruction cache
  // at a 98 % hit rate (one barrier per step: 29.96 -> 30.43 ms at 2^24)
  static constexpr bool kAccLockstep = false;
  // how much larger the next pipelined host range may be: bucket work per point / PCIe time
  // per point (2.4 ns vs 1.75 ns for 96 B at ~55 GB/s), with a margin
  static constexpr double kRangeGrowth = 1.3;
};
struct Bls381Curve {
  using Fq = Bls381FqParams;
  using Field = FpField<Bls381FqParams>;
  using Fr = Bls381FrParams;
  using Gen = Bls381G1Generator;
  static constexpr const char* kName = "bls12_381";
  static constexpr bool kReduceInline = false;  // running-sum kernel: addition out of line (inlined: 2.13 -> 2.21 ms at 2^22)
  static constexpr int kReduceRoll = 1;  // running-sum kernel: looped multiplications (2^22: 2.24 -> 2.12 ms)
  // accumulation warps of a CTA in step (accumulate_lockstep_kernel): 2^22 points 21.64 ->
  // 19.71 ms, FMA-heavy pipe 82 -> 90 %, instruction-cache hit rate 79 -> 90 %
  static constexpr bool kAccLockstep = true;
  static constexpr double kRangeGrowth = 2.0;  // 6.3 ns of bucket work vs 2.3 ns of PCIe per point
};

// G2: the same curves over Fq2 (bn254/BUILD.bazel:150-200, bls12_381/BUILD.bazel:153-200);
// SURVEY 8f-2, the B2 query of a Groth16 proof (zk/r1cs/groth16/prove.h:129-131).
struct Bn254G2Curve {
  using Fq = Bn254FqParams;
  using Field = Fp2Field<Bn254FqParams>;
  using Fr = Bn254FrParams;
  using Gen = Bn254G2Generator;
  static constexpr const char* kName = "bn254_g2";
  static constexpr bool kAccLockstep = true;  // pair kernel: warps of a CTA in step
  static constexpr int kReduceRoll = 1;  // running-sum kernel: looped multiplications (2^20: 1.11 -> 1.06 ms)
  static constexpr double kRangeGrowth = 2.0;
};
struct Bls381G2Curve {
  using Fq = Bls381FqParams;
  using Field = Fp2Field<Bls381FqParams>;
  using Fr = Bls381FrParams;
  using Gen = Bls381G2Generator;
  static constexpr const char* kName = "bls12_381_g2";
  static constexpr bool kAccLockstep = true;  // pair kernel: warps of a CTA in step
  static constexpr int kReduceRoll = 1;  // running-sum kernel: looped multiplications (2^20: 4.09 -> 3.93 ms)
  static constexpr double kRangeGrowth = 2.5;
};

struct MsmPlan {
  uint32_t n;        // points
  uint32_t c;        // window bits
  uint32_t W;        // windows, W * c >= scalar bits + 1
  uint32_t B;        // buckets per window = 2^(c-1)
  uint32_t TB;       // W * B
  uint32_t seg;      // max entries per task
  uint32_t max_tasks;
  uint32_t aggregate;  // warp-aggregate the bucket atomics (pays off for repeated digits)
  uint32_t R;          // pair rounds of the batched-affine pre-reduction; bucket runs in
                       // `sorted` start at multiples of 2^R and are padded with kNoEntry
  uint32_t shared;     // 1: all windows share ONE set of B buckets — the bases are a table of
                       // precomputed multiples T[w][i] = 2^(bit offset of window w) * P_i
                       // (`stride` points per window), so the digit of window w selects
                       // T[w][i] and no window weights remain (TB = B)
  uint32_t stride;     // points per window slice of the precomputed table
  uint32_t wide;       // windows [0, wide) take c bits, windows [wide, W) take c - 1 (balanced
                       // windows: the slack W * c - (bits + 1) is spread over the top windows
                       // instead of leaving one nearly empty, heavily loaded top window)
};

constexpr uint32_t kNoEntry = 0xffffffffu;  // padding slot in `sorted`: the identity

// Device-side totals of one point range.  entries / tasks / multi are written by the scan,
// seg by choose_segment_kernel before it, fold_jobs / multi2 by the task builder.
struct MsmTotals {
  uint32_t entries;    // (padded) entries in `sorted`
  uint32_t tasks;      // accumulation tasks
  uint32_t multi;      // buckets split into several tasks
  uint32_t seg;        // task length limit chosen from the measured bucket occupancy
  uint32_t fold_jobs;  // chunks of kFoldThreads partial sums to fold (stage A)
  uint32_t multi2;     // buckets with more than one chunk (stage B)
  uint32_t nonzero;    // non-zero digits of the range
  uint32_t tasks_hi;   // tasks of the HIGH window group (bucket keys >= split_key); they come
                       // first in `order`, the low group's follow
  uint32_t entries_lo; // entries of the low window group (= offset[split_key])
  uint32_t pad;
};

// Which tasks a launch of the accumulation / fold kernels covers.  The windows of the last point
// range are processed as two groups: the high windows first, so that their bucket reduction and
// the long doubling chain of the window combination (window_combine_kernel) run on a second
// stream WHILE the low windows are still being accumulated.
enum : uint32_t { kPartAll = 0, kPartHigh = 1, kPartLow = 2 };

constexpr uint32_t kMinSegment = 16;
constexpr int kNonzeroSlots = 64;  // partial counters of non-zero digits (spread the atomics)

// Task length limit: the smallest power of two >= 4x the MEASURED mean bucket run, within
// [kMinSegment, seg_max].  The host cannot know the occupancy: witness-like scalars (mostly
// 0 / 1 / small) fill a fraction of the buckets a uniform draw would, and with the limit sized
// for uniform scalars their few heavy buckets became a handful of 128-entry tasks that ran
// alone for 0.9 ms (BN254 2^20: accumulate 1.32 ms for 0.35 ms of work).
static __global__ void choose_segment_kernel(const uint32_t* __restrict__ nonzero_slots,
                                             uint32_t total_buckets, uint32_t R, uint32_t seg_max,
                                             uint32_t seg_forced, MsmTotals* __restrict__ totals) {
  uint32_t v = threadIdx.x < kNonzeroSlots ? nonzero_slots[threadIdx.x] : 0u;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ uint32_t part[2];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t nonzero = part[0] + part[1];
    uint32_t tb = total_buckets ? total_buckets : 1u;
    uint32_t mean_run = ((nonzero + tb - 1) / tb) >> R;  // rounded up, as the host's n / B + 1
    uint32_t seg = kMinSegment;
    while (seg < seg_max && seg < 4u * mean_run + 4u) seg <<= 1;
    // small MSMs end when their longest task ends (a 128-entry task runs ~0.9 ms): keep the
    // limit near (entries / 32 K), but not below twice the mean run
    uint32_t par = kMinSegment;
    while (par < seg && par < (nonzero >> (15 + R))) par <<= 1;
    while (par < seg && par < 2u * mean_run) par <<= 1;
    seg = par;
    totals->seg = seg_forced ? seg_forced : seg;
    totals->nonzero = nonzero;
    totals->fold_jobs = 0;
    totals->multi2 = 0;
  }
}

// ---------------------------------------------------------------------------
// Scalar recoding.  Same digits as pippenger.h:27-51 FillDigits up to the
// representation of +-2^(c-1) (taken as +2^(c-1) here, -2^(c-1) there; both
// name the same multiple) and with W chosen so the top digit never carries
// out, hence every window has exactly 2^(c-1) buckets.
// ---------------------------------------------------------------------------
template <class Fr>
TB_DEV void load_scalar_canonical(Fp<Fr>& s, const uint32_t* scalars, uint32_t i) {
  Fp<Fr> m;
  fp_load<Fr>(m, scalars + (size_t)i * Fp<Fr>::N);
  fp_from_mont<Fr>(s, m);
}

// Pops the low c bits and shifts the scalar right by c.
template <int N>
TB_DEV uint32_t pop_window(uint32_t (&s)[N], uint32_t c) {
  uint32_t bits = s[0] & ((1u << c) - 1u);
#pragma unroll
  for (int i = 0; i < N - 1; ++i) s[i] = __funnelshift_r(s[i], s[i + 1], c);
  s[N - 1] >>= c;
  return bits;
}

// Calls f(w, bucket_key, negative) for every non-zero digit.  Window w is cw = c or c - 1 bits
// wide (plan.wide); its digits lie in [-2^(cw-1), 2^(cw-1)] and use the first 2^(cw-1) of the
// window's B bucket slots.
template <class Fr, class Fn>
TB_DEV void for_each_digit(Fp<Fr>& s, const MsmPlan& plan, Fn f) {
  uint32_t carry = 0;
  for (uint32_t w = 0; w < plan.W; ++w) {
    const uint32_t cw = plan.c - (w >= plan.wide ? 1u : 0u);
    const uint32_t half = 1u << (cw - 1);
    uint32_t d = pop_window(s.l, cw) + carry;
    bool neg = d > half;
    carry = neg ? 1u : 0u;
    uint32_t mag = neg ? (2u * half - d) : d;
    f(w, mag, neg);
  }
}

// Bucket counter increment for one lane's key; returns the previous value.  Repeated keys
// inside a warp (small scalars, the reference's "non_uniform" test set, witness vectors full
// of 0/1) would serialise at one L2 address, so when a cheap neighbour test sees a duplicate
// the warp switches to one atomic per distinct key (match.any).  The test is a single
// shuffle + vote; match.any itself saturates the ADU pipe (ncu: 95 %) and is not worth
// paying on uniformly random digits.
TB_DEV uint32_t bucket_inc(uint32_t* counter_base, uint32_t key, bool valid, bool aggregate) {
  uint32_t result = 0;
  bool dup = false;
  if (aggregate) {
    uint32_t other = __shfl_xor_sync(0xffffffffu, valid ? key : 0xffffffffu, 1);
    dup = __any_sync(0xffffffffu, valid && other == key);
  }
  if (!dup) {
    if (valid) result = atomicAdd(counter_base + key, 1u);
    return result;
  }
  uint32_t vote = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    uint32_t peers = __match_any_sync(vote, key);
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(peers) - 1;
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter_base + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    result = base + rank;
  }
  return result;
}

// Pass 1 over the scalars: de-Montgomery, recode, write the digit words window-major
// (digits[w * n + i] = |d| | sign << 31, 0 for a zero digit; coalesced per window) and count
// bucket sizes.
template <class C>
__global__ void __launch_bounds__(256, 6) digits_hist_kernel(const uint32_t* __restrict__ scalars,
                                                          MsmPlan plan,
                                                          uint32_t* __restrict__ digits,
                                                          uint32_t* __restrict__ count,
                                                          uint32_t* __restrict__ nonzero_slots) {
  using Fr = typename C::Fr;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool in = i < plan.n;
  Fp<Fr> s;
  if (in) {
    load_scalar_canonical<Fr>(s, scalars, i);
  } else {
    fp_set_zero<Fr>(s);
  }
  uint32_t nz = 0;
  for_each_digit<Fr>(s, plan, [&](uint32_t w, uint32_t mag, bool neg) {
    bool valid = in && mag != 0;
    if (in) digits[(size_t)w * plan.n + i] = mag | (neg ? 0x80000000u : 0u);
    bucket_inc(count, (plan.shared ? 0u : w * plan.B) + mag - 1, valid, plan.aggregate != 0);
    nz += valid;
  });
  // non-zero digits of the CTA -> one of kNonzeroSlots counters (input of choose_segment)
  __shared__ uint32_t warp_nz[8];
  nz = __reduce_add_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0) warp_nz[threadIdx.x >> 5] = nz;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) total += warp_nz[k];
    if (total) atomicAdd(nonzero_slots + (blockIdx.x % kNonzeroSlots), total);
  }
}

// Pass 2, window-major (blockIdx.y = window): all CTAs of one window run before the next
// window's, so the slice of `sorted` being filled (4 n bytes) and the window's cursors
// (2^(c+1) bytes) stay resident in the 126 MB L2 and every 32-byte sector goes to HBM once.
// The scalar-major form wrote 4-byte words at random over the whole n*W array: 14.7 GB of
// DRAM traffic and L2-missing atomics at n = 2^24 (profiles/r1_b_*).
// cursor[key] starts at offset[key]; the value returned by the atomic is the slot.
static __global__ void __launch_bounds__(256, 8) digits_scatter_kernel(const uint32_t* __restrict__ digits,
                                                             MsmPlan plan,
                                                             uint32_t* __restrict__ cursor,
                                                             uint32_t* __restrict__ sorted) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w = blockIdx.y;
  bool in = i < plan.n;
  uint32_t d = in ? digits[(size_t)w * plan.n + i] : 0u;
  uint32_t mag = d & 0x7fffffffu;
  bool valid = mag != 0;
  uint32_t pos = bucket_inc(cursor, (plan.shared ? 0u : w * plan.B) + mag - 1, valid, plan.aggregate != 0);
  // shared buckets: the entry names the window's slice of the precomputed table
  if (valid) sorted[pos] = (plan.shared ? w * plan.stride + i : i) | (d & 0x80000000u);
}

// ---------------------------------------------------------------------------
// Exclusive scan over TB bucket sizes, carrying two sums at once in a u64:
// low word = entries (bucket offsets), high word = tasks (ceil(size / seg)).
// Three passes; kScanItems per block.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanItems = kScanThreads * kScanPerThread;

// cnt entries occupy cntp = cnt rounded up to 2^R slots; after the R pair rounds the bucket
// is a run of cntp >> R affine points, cut into tasks of <= seg.
TB_DEV uint32_t padded_count(uint32_t cnt, uint32_t R) {
  uint32_t a = (1u << R) - 1u;
  return (cnt + a) & ~a;
}
TB_DEV uint64_t scan_item(uint32_t cnt, uint32_t seg, uint32_t R) {
  uint32_t cntp = padded_count(cnt, R);
  uint32_t t = ((cntp >> R) + seg - 1) / seg;
  return ((uint64_t)t << 32) | cntp;
}

TB_DEV uint64_t block_exclusive_scan(uint64_t v, uint64_t* total, uint64_t* smem) {
  // smem: kScanThreads/32 words
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint64_t s = (lane < kScanThreads / 32) ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= (uint32_t)o) s += y;
    }
    if (lane < kScanThreads / 32) smem[lane] = s;
  }
  __syncthreads();
  uint64_t warp_prefix = warp ? smem[warp - 1] : 0;
  *total = smem[kScanThreads / 32 - 1];
  __syncthreads();
  return warp_prefix + x - v;
}

static __global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(
    const uint32_t* __restrict__ count, uint32_t n, const MsmTotals* __restrict__ totals,
    uint32_t R, uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t smem[kScanThreads / 32];
  const uint32_t seg = totals->seg;
  uint32_t base = blockIdx.x * kScanItems + threadIdx.x;  // a sum: any order, so read coalesced
  uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    uint32_t idx = base + k * kScanThreads;
    if (idx < n) sum += scan_item(count[idx], seg, R);
  }
  uint64_t total;
  block_exclusive_scan(sum, &total, smem);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block; nblocks <= kScanItems
static __global__ void __launch_bounds__(kScanThreads) scan_top_kernel(uint64_t* __restrict__ block_sums,
                                                                uint32_t nblocks,
                                                                MsmTotals* __restrict__ totals) {
  __shared__ uint64_t smem[kScanThreads / 32];
  uint64_t v[kScanPerThread];
  uint64_t sum = 0;
  uint32_t base = threadIdx.x * kScanPerThread;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    v[k] = (base + k < nblocks) ? block_sums[base + k] : 0;
    sum += v[k];
  }
  uint64_t total;
  uint64_t prefix = block_exclusive_scan(sum, &total, smem);
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    if (base + k < nblocks) block_sums[base + k] = prefix;
    prefix += v[k];
  }
  if (threadIdx.x == 0) {
    totals->entries = (uint32_t)total;
    totals->tasks = (uint32_t)(total >> 32);
    totals->multi = 0;
  }
}

// task_meta word: bucket key in the low 24 bits; kTaskFirst = first task of its bucket (its
// accumulator starts from the bucket's running value in `state`), kTaskSingle = the only
// task of its bucket (its result is the bucket's new value and goes straight to `state`).
constexpr uint32_t kTaskFirst = 0x80000000u;
constexpr uint32_t kTaskSingle = 0x40000000u;
constexpr uint32_t kTaskKeyMask = 0x00ffffffu;

constexpr int kFoldThreads = 128;  // partial sums folded by one CTA of stage A

// Writes offset[] (TB+1 entries), cursor[] (= offset, consumed by the scatter)
// and the tasks of every bucket.  A bucket split into t > 1 tasks gets ceil(t / kFoldThreads)
// fold jobs {bucket, chunk} (stage A); with more than one chunk it is also listed in multi2
// (stage B).
static __global__ void __launch_bounds__(kScanThreads) scan_apply_build_tasks_kernel(
    const uint32_t* __restrict__ count, uint32_t n, uint32_t R,
    const uint64_t* __restrict__ block_prefix, uint32_t* __restrict__ offset,
    uint32_t* __restrict__ cursor, uint32_t* __restrict__ task_base, uint2* __restrict__ tasks,
    uint32_t* __restrict__ task_meta, uint2* __restrict__ fold_jobs,
    uint32_t* __restrict__ multi2_keys, uint32_t* __restrict__ sorted,
    MsmTotals* __restrict__ totals) {
  // A warp owns 32 * kScanPerThread consecutive buckets and walks them in rows of 32: lane l
  // handles bucket row * 32 + l, so the loads of count[] and the stores of offset / cursor /
  // task_base / tasks are coalesced (neighbouring lanes <-> neighbouring buckets <-> neighbouring
  // task slots).  The exclusive prefix of a bucket = block prefix + earlier warps + earlier rows
  // of this warp + a shuffle scan inside the row.
  __shared__ uint64_t warp_total[kScanThreads / 32];
  const uint32_t seg = totals->seg;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t wbase = blockIdx.x * kScanItems + warp * (32 * kScanPerThread) + lane;
  // pass 1: this warp's total; pass 2 re-reads the counts (L2 hits) row by row with a running
  // prefix, so nothing is kept in per-thread arrays (the one-pass form needed 254 registers)
  uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    uint32_t idx = wbase + k * 32;
    if (idx < n) sum += scan_item(count[idx], seg, R);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) warp_total[warp] = sum;
  __syncthreads();
  uint64_t run = block_prefix[blockIdx.x];
  for (uint32_t v = 0; v < warp; ++v) run += warp_total[v];
#pragma unroll 2
  for (int k = 0; k < kScanPerThread; ++k) {
    uint32_t idx = wbase + k * 32;
    uint32_t cnt_k = (idx < n) ? count[idx] : 0;
    uint64_t item = scan_item(cnt_k, seg, R);
    uint64_t x = item;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (uint32_t)o) x += y;
    }
    uint64_t prefix = run + x - item;
    run += __shfl_sync(0xffffffffu, x, 31);
    if (idx < n) {
      uint32_t off = (uint32_t)prefix;
      uint32_t tb = (uint32_t)(prefix >> 32);
      offset[idx] = off;
      cursor[idx] = off;
      task_base[idx] = tb;
      uint32_t cntp = padded_count(cnt_k, R);
      uint32_t run_len = cntp >> R;  // points left after the pair rounds
      uint32_t t = (run_len + seg - 1) / seg;
      for (uint32_t s = 0; s < t; ++s) {
        uint32_t len = min(seg, run_len - s * seg);
        tasks[tb + s] = make_uint2((off >> R) + s * seg, len);
        task_meta[tb + s] = idx | (s == 0 ? kTaskFirst : 0u) | (t == 1 ? kTaskSingle : 0u);
      }
      for (uint32_t q = cnt_k; q < cntp; ++q) sorted[off + q] = kNoEntry;
      if (t > 1) {
        atomicAdd(&totals->multi, 1u);
        uint32_t chunks = (t + kFoldThreads - 1) / kFoldThreads;
        uint32_t jb = atomicAdd(&totals->fold_jobs, chunks);
        for (uint32_t j = 0; j < chunks; ++j) fold_jobs[jb + j] = make_uint2(idx, j);
        if (chunks > 1) multi2_keys[atomicAdd(&totals->multi2, 1u)] = idx;
      }
      if (idx == n - 1) offset[n] = off + cntp;
    }
  }
}

// ---------------------------------------------------------------------------
// Task ordering: a counting sort of task ids by length, longest first, so that
// the 32 lanes of a warp of the accumulation kernel run (almost) the same
// number of mixed additions.  Lengths are <= kMaxSegment.
//   order_hist     per-CTA shared histogram -> global histogram
//   order_scan     one CTA: descending exclusive scan -> first slot of every length
//   order_scatter  per-CTA reservation of a slot range per length, then local ranks
// ---------------------------------------------------------------------------
constexpr int kMaxSegment = 1024;
constexpr int kOrderThreads = 256;
constexpr int kOrderPerThread = 8;
// sort key of a task: window group (0 = high, 1 = low) major, then DEscending length
constexpr int kOrderBins = 2 * (kMaxSegment + 1);
TB_DEV uint32_t order_bin(uint32_t len, uint32_t meta, uint32_t split_key) {
  uint32_t low = (meta & kTaskKeyMask) < split_key ? 1u : 0u;
  return low * (kMaxSegment + 1) + (kMaxSegment - len);
}

static __global__ void __launch_bounds__(kOrderThreads, 8) order_hist_kernel(
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const MsmTotals* __restrict__ totals, uint32_t split_key, uint32_t* __restrict__ len_hist) {
  __shared__ uint32_t sh[kOrderBins];
  for (int i = threadIdx.x; i < kOrderBins; i += kOrderThreads) sh[i] = 0;
  __syncthreads();
  uint32_t T = totals->tasks;
  uint32_t base = blockIdx.x * (kOrderThreads * kOrderPerThread);
#pragma unroll
  for (int k = 0; k < kOrderPerThread; ++k) {
    uint32_t g = base + k * kOrderThreads + threadIdx.x;
    if (g < T) atomicAdd(&sh[order_bin(tasks[g].y, task_meta[g], split_key)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kOrderBins; i += kOrderThreads)
    if (sh[i]) atomicAdd(&len_hist[i], sh[i]);
}

// in: len_hist[bin] counts; out: len_hist[bin] = first slot of the bin (exclusive scan in bin
// order); totals->tasks_hi = first slot of the low group; totals->entries_lo from the offsets.
static __global__ void __launch_bounds__(1024) order_scan_kernel(
    uint32_t* __restrict__ len_hist, const uint32_t* __restrict__ offset, uint32_t split_key,
    MsmTotals* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  constexpr int kPer = (kOrderBins + 1023) / 1024;  // bins per thread, consecutive
  const uint32_t first = threadIdx.x * kPer;
  uint32_t v[kPer], sum = 0;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    v[k] = first + k < (uint32_t)kOrderBins ? len_hist[first + k] : 0u;
    sum += v[k];
  }
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= (uint32_t)o) t += y;
    }
    warp_sums[lane] = t;
  }
  __syncthreads();
  uint32_t run = (warp ? warp_sums[warp - 1] : 0) + x - sum;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    if (first + k < (uint32_t)kOrderBins) {
      len_hist[first + k] = run;
      if (first + k == (uint32_t)(kMaxSegment + 1)) totals->tasks_hi = run;
    }
    run += v[k];
  }
  if (threadIdx.x == 0) totals->entries_lo = split_key ? offset[split_key] : 0u;
}

static __global__ void __launch_bounds__(kOrderThreads, 8) order_scatter_kernel(
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const MsmTotals* __restrict__ totals, uint32_t split_key, uint32_t* __restrict__ len_cursor,
    uint32_t* __restrict__ order) {
  __shared__ uint32_t cnt[kOrderBins];
  __shared__ uint32_t start[kOrderBins];
  for (int i = threadIdx.x; i < kOrderBins; i += kOrderThreads) cnt[i] = 0;
  __syncthreads();
  uint32_t T = totals->tasks;
  uint32_t base = blockIdx.x * (kOrderThreads * kOrderPerThread);
  uint32_t bin[kOrderPerThread], rank[kOrderPerThread];
#pragma unroll
  for (int k = 0; k < kOrderPerThread; ++k) {
    uint32_t g = base + k * kOrderThreads + threadIdx.x;
    bin[k] = (g < T) ? order_bin(tasks[g].y, task_meta[g], split_key) : 0xffffffffu;
    rank[k] = (g < T) ? atomicAdd(&cnt[bin[k]], 1u) : 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kOrderBins; i += kOrderThreads)
    if (cnt[i]) start[i] = atomicAdd(&len_cursor[i], cnt[i]);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kOrderPerThread; ++k) {
    uint32_t g = base + k * kOrderThreads + threadIdx.x;
    if (bin[k] != 0xffffffffu) order[start[bin[k]] + rank[k]] = g;
  }
}

// ---------------------------------------------------------------------------
// Bucket accumulation — the hot loop (pippenger.h:112-135).  One thread per
// task; the next point is fetched while the current one is being added.
// ---------------------------------------------------------------------------
constexpr int kAccThreads = 128;
// resident CTAs per SM the register budget is held to: 4 x 128 threads x 128 registers for
// both G1 curves (BLS12-381 at 3 x 168 registers had no spills but fewer warps to hide the
// multiply chains: accumulate 21.8 -> 21.4 ms at 2^22 with 152 B of spills), 2 x 128 x 255
// (BN254 G2; 3 x 168 measured slower, 10.15 -> 10.69 ms at 2^20), 1 (BLS12-381 G2)
template <class C>
constexpr int AccMinBlocks() {
  return C::Field::kWords <= 12 ? 4 : (C::Field::kWords <= 16 ? 2 : 1);
}

// Small ranges of the 12-limb curve are latency-bound rather than pipe-bound and run faster with
// the spill-free 3 CTAs/SM build (BLS12-381 2^19 points: 3.71 ms against 3.96 at 4 CTAs/SM;
// from 2^20 points the 4-CTA build wins): a second instantiation, picked by entry count.
template <class C>
constexpr int AccMinBlocksSmall() {
  return C::Field::kWords == 12 ? 3 : AccMinBlocks<C>();
}
constexpr uint64_t kAccSmallEntries = 12u << 20;  // below: the AccMinBlocksSmall build

// kReduced: the task's points are a run of affine points left by the pair rounds (read in
// order, no index or sign); otherwise they are gathered from `bases` through `sorted`.
template <class C, bool kReduced, int kMinBlocks = AccMinBlocks<C>()>
__global__ void __launch_bounds__(kAccThreads, kMinBlocks) accumulate_kernel(
    const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const uint32_t* __restrict__ order, const MsmTotals* __restrict__ totals, uint32_t part,
    uint32_t* __restrict__ state, uint32_t* __restrict__ task_out) {
  using K = typename C::Field;
  constexpr int kAffineWords = 2 * K::kWords;
  constexpr int kXyzzWords = 4 * K::kWords;
  // slots [0, tasks_hi) are the high window group, [tasks_hi, tasks) the low one
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x + (part == kPartLow ? totals->tasks_hi : 0u);
  if (slot >= (part == kPartHigh ? totals->tasks_hi : totals->tasks)) return;
  uint32_t g = order[slot];  // tasks in descending length: warps stay convergent
  uint2 task = tasks[g];
  uint32_t meta = task_meta[g];
  uint32_t* bucket = state + (size_t)(meta & kTaskKeyMask) * kXyzzWords;
  const uint32_t* ent = sorted + task.x;
  const uint32_t* run = bases + (size_t)task.x * kAffineWords;
  uint32_t e = kReduced ? 0u : ent[0];
  Affine<K> nxt;
  if (kReduced) {
    affine_load<K>(nxt, run);
  } else {
    affine_load<K>(nxt, bases + (size_t)(e & 0x7fffffffu) * kAffineWords);
  }
  XYZZ<K> acc;
  if (meta & kTaskFirst) {
    xyzz_load<K>(acc, bucket);  // value left by the earlier point ranges (zz == 0: none)
  } else {
    xyzz_set_zero<K>(acc);
  }
  for (uint32_t j = 0; j < task.y; ++j) {
    Affine<K> cur = nxt;
    bool neg = e >> 31;
    if (j + 1 < task.y) {
      if (kReduced) {
        affine_load<K>(nxt, run + (size_t)(j + 1) * kAffineWords);
      } else {
        e = ent[j + 1];
        affine_load<K>(nxt, bases + (size_t)(e & 0x7fffffffu) * kAffineWords);
      }
    }
    xyzz_madd<K>(acc, cur, neg);
  }
  xyzz_store<K>((meta & kTaskSingle) ? bucket : task_out + (size_t)g * kXyzzWords, acc);
}

// G2 accumulation with one task per LANE PAIR: lane 2k holds the c0 components of the bucket
// value and of the points, lane 2k + 1 the c1 components (fp.cuh Fp2Lanes, xyzz.cuh PairOps).
// Half the registers per lane of accumulate_kernel<G2> — twice the warps per SM for the same
// multiply count — and warps that stay converged: every lane pair of a warp walks
// max(task lengths of the warp) steps (tasks are ordered by length, so the lengths of one warp
// are nearly equal) with finished pairs predicated off, which is what a diverged warp costs
// anyway.  Memory layout of points and bucket values is unchanged (c0 | c1 per coordinate): a
// lane reads and writes its half of every coordinate.
// Resident CTAs per SM (of kAccThreads lanes) the pair kernel's register budget is held to.
// Measured (B200, 2^20 points, accumulation only; one thread per task: 11.01 / 22.34 ms):
//   BN254      3 x 168 registers (12 B of spills)  8.48 ms     4 x 128 (144 B)  8.55 ms
//   BLS12-381  2 x 242 registers (no spills)      18.14 ms     3 x 168 (240 B) 19.57 ms
template <class C>
constexpr int PairMinBlocks() {
  return C::Field::kWords <= 16 ? 3 : 2;
}
template <class C>
constexpr int PairMinBlocksAlt() {
  return C::Field::kWords <= 16 ? 4 : 3;
}

// kLockstep: the warps of a CTA also stay within one addition of each other (one barrier per
// step, see accumulate_lockstep_kernel) so that they share the instruction stream.
template <class C, int kThreads, int kMinBlocks, bool kLockstep = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks) accumulate_pair_kernel(
    const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const uint32_t* __restrict__ order, const MsmTotals* __restrict__ totals, uint32_t part,
    uint32_t* __restrict__ state, uint32_t* __restrict__ task_out) {
  using F = typename C::Fq;
  static_assert(C::Field::kDegree == 2, "lane pairs split a quadratic extension");
  constexpr int N = Fp<F>::N;
  constexpr int kAffineWords = 4 * N;
  constexpr int kXyzzWords = 8 * N;
  const uint32_t role = threadIdx.x & 1;
  const uint32_t slot = ((blockIdx.x * kThreads + threadIdx.x) >> 1) +
                        (part == kPartLow ? totals->tasks_hi : 0u);
  const bool live = slot < (part == kPartHigh ? totals->tasks_hi : totals->tasks);
  uint32_t g = 0, meta = 0;
  uint2 task = make_uint2(0u, 0u);
  if (live) {
    g = order[slot];
    task = tasks[g];
    meta = task_meta[g];
  }
  const uint32_t len = live ? task.y : 0u;
  uint32_t steps = __reduce_max_sync(0xffffffffu, len);  // warp-uniform trip count
  if (kLockstep) {
    __shared__ uint32_t cta_steps;
    if (threadIdx.x == 0) cta_steps = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&cta_steps, steps);
    __syncthreads();
    steps = cta_steps;
  }
  const uint32_t* ent = sorted + task.x;
  uint32_t* bucket = state + (size_t)(meta & kTaskKeyMask) * kXyzzWords + role * N;
  auto fetch = [&](PairAffine<F>& p, uint32_t e) {
    const uint32_t* src = bases + (size_t)(e & 0x7fffffffu) * kAffineWords + role * N;
    fp_load<F>(p.x, src);
    fp_load<F>(p.y, src + 2 * N);
  };
  uint32_t e = 0;
  PairAffine<F> nxt;
  fp_set_zero<F>(nxt.x);
  fp_set_zero<F>(nxt.y);
  if (len > 0) {
    e = ent[0];
    fetch(nxt, e);
  }
  PairPoint<F> acc;
  if (live && (meta & kTaskFirst)) {
    fp_load_rw<F>(acc.x, bucket);
    fp_load_rw<F>(acc.y, bucket + 2 * N);
    fp_load_rw<F>(acc.zz, bucket + 4 * N);
    fp_load_rw<F>(acc.zzz, bucket + 6 * N);
  } else {
    Fp2Lanes<F>::set_one(acc.x, role);
    Fp2Lanes<F>::set_one(acc.y, role);
    fp_set_zero<F>(acc.zz);
    fp_set_zero<F>(acc.zzz);
  }
  for (uint32_t j = 0; j < steps; ++j) {
    PairAffine<F> cur = nxt;
    const bool neg = e >> 31;
    if (j + 1 < len) {
      e = ent[j + 1];
      fetch(nxt, e);
    }
    PairOps<F>::madd(acc, cur, neg, j < len, role);
    if (kLockstep) __syncthreads();
  }
  if (live) {
    uint32_t* dst = (meta & kTaskSingle) ? bucket : task_out + (size_t)g * kXyzzWords + role * N;
    fp_store<F>(dst, acc.x);
    fp_store<F>(dst + 2 * N, acc.y);
    fp_store<F>(dst + 4 * N, acc.zz);
    fp_store<F>(dst + 6 * N, acc.zzz);
  }
}

// Experiment: accumulate_kernel with the warps of a CTA kept within one addition of each other
// (one barrier per step; every thread walks the CTA's longest task length with its own additions
// predicated), so that the four warps fetch the same stretch of the loop body at about the same
// time.  The 12-limb loop body is 90 KB of SASS and the instruction cache serves only 79 % of
// its requests with 16 warps at unrelated places in it (ncu: no_instruction is the second
// largest stall).
template <class C, int kMinBlocks = AccMinBlocks<C>()>
__global__ void __launch_bounds__(kAccThreads, kMinBlocks) accumulate_lockstep_kernel(
    const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const uint32_t* __restrict__ order, const MsmTotals* __restrict__ totals, uint32_t part,
    uint32_t* __restrict__ state, uint32_t* __restrict__ task_out) {
  using K = typename C::Field;
  constexpr int kAffineWords = 2 * K::kWords;
  constexpr int kXyzzWords = 4 * K::kWords;
  __shared__ uint32_t cta_steps;
  if (threadIdx.x == 0) cta_steps = 0;
  __syncthreads();
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x + (part == kPartLow ? totals->tasks_hi : 0u);
  const bool live = slot < (part == kPartHigh ? totals->tasks_hi : totals->tasks);
  uint32_t g = 0, meta = 0;
  uint2 task = make_uint2(0u, 0u);
  if (live) {
    g = order[slot];
    task = tasks[g];
    meta = task_meta[g];
  }
  const uint32_t len = live ? task.y : 0u;
  const uint32_t warp_steps = __reduce_max_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0) atomicMax(&cta_steps, warp_steps);
  __syncthreads();
  const uint32_t steps = cta_steps;
  uint32_t* bucket = state + (size_t)(meta & kTaskKeyMask) * kXyzzWords;
  const uint32_t* ent = sorted + task.x;
  uint32_t e = 0;
  Affine<K> nxt;
  K::set_zero(nxt.x);
  K::set_zero(nxt.y);
  if (len > 0) {
    e = ent[0];
    affine_load<K>(nxt, bases + (size_t)(e & 0x7fffffffu) * kAffineWords);
  }
  XYZZ<K> acc;
  if (live && (meta & kTaskFirst)) {
    xyzz_load<K>(acc, bucket);
  } else {
    xyzz_set_zero<K>(acc);
  }
  for (uint32_t j = 0; j < steps; ++j) {
    if (j < len) {
      Affine<K> cur = nxt;
      bool neg = e >> 31;
      if (j + 1 < len) {
        e = ent[j + 1];
        affine_load<K>(nxt, bases + (size_t)(e & 0x7fffffffu) * kAffineWords);
      }
      xyzz_madd<K>(acc, cur, neg);
    }
    __syncthreads();
  }
  if (live) xyzz_store<K>((meta & kTaskSingle) ? bucket : task_out + (size_t)g * kXyzzWords, acc);
}

// The same accumulation with the NEXT point staged through shared memory instead of registers:
// cp.async (LDGSTS) copies the gathered point straight from L2 into the thread's private slot
// while the current mixed addition runs, so nothing of the point in flight is live in registers
// across the addition (the register form holds 16 / 24 / 32 / 48 words there — with the 128
// register budget of 4 CTAs/SM that is what spills on the 12-limb curve).  Two slots per thread
// (the slot being read and the slot being filled), laid out [slot][16-byte chunk][thread] so
// that the 128-bit shared-memory accesses of a warp are conflict-free.
TB_DEV void cp_async16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
TB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
TB_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class C, int kMinBlocks = AccMinBlocks<C>()>
__global__ void __launch_bounds__(kAccThreads, kMinBlocks) accumulate_staged_kernel(
    const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
    const uint2* __restrict__ tasks, const uint32_t* __restrict__ task_meta,
    const uint32_t* __restrict__ order, const MsmTotals* __restrict__ totals, uint32_t part,
    uint32_t* __restrict__ state, uint32_t* __restrict__ task_out) {
  using K = typename C::Field;
  constexpr int kAffineWords = 2 * K::kWords;
  constexpr int kXyzzWords = 4 * K::kWords;
  constexpr int kChunks = kAffineWords / 4;  // 16-byte chunks of one affine point
  __shared__ uint4 stage[2][kChunks][kAccThreads];
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x + (part == kPartLow ? totals->tasks_hi : 0u);
  if (slot >= (part == kPartHigh ? totals->tasks_hi : totals->tasks)) return;
  uint32_t g = order[slot];
  uint2 task = tasks[g];
  uint32_t meta = task_meta[g];
  uint32_t* bucket = state + (size_t)(meta & kTaskKeyMask) * kXyzzWords;
  const uint32_t* ent = sorted + task.x;
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(&stage[0][0][threadIdx.x]);
  constexpr uint32_t kChunkStride = kAccThreads * 16, kSlotStride = kChunks * kChunkStride;
  auto fetch = [&](uint32_t entry, uint32_t buf) {
    const uint4* src = reinterpret_cast<const uint4*>(bases + (size_t)(entry & 0x7fffffffu) * kAffineWords);
#pragma unroll
    for (int q = 0; q < kChunks; ++q) cp_async16(my + buf * kSlotStride + q * kChunkStride, src + q);
    cp_async_commit();
  };
  uint32_t e = ent[0];
  fetch(e, 0);
  uint32_t e_next = task.y > 1 ? ent[1] : 0u;
  XYZZ<K> acc;
  if (meta & kTaskFirst) {
    xyzz_load<K>(acc, bucket);
  } else {
    xyzz_set_zero<K>(acc);
  }
  for (uint32_t j = 0; j < task.y; ++j) {
    cp_async_wait_all();  // point j has landed in slot j & 1
    Affine<K> cur;
    {
      uint32_t* w = reinterpret_cast<uint32_t*>(&cur);
#pragma unroll
      for (int q = 0; q < kChunks; ++q) {
        uint4 v = stage[j & 1][q][threadIdx.x];
        w[4 * q] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
      }
    }
    const bool neg = e >> 31;
    if (j + 1 < task.y) {
      e = e_next;
      fetch(e, (j + 1) & 1);
      if (j + 2 < task.y) e_next = ent[j + 2];
    }
    xyzz_madd<K>(acc, cur, neg);
  }
  xyzz_store<K>((meta & kTaskSingle) ? bucket : task_out + (size_t)g * kXyzzWords, acc);
}

// ---------------------------------------------------------------------------
// Batched-affine pre-reduction ("pair rounds").  A mixed XYZZ addition costs 10 field
// multiplications; an affine + affine addition costs 3 (lambda, lambda^2, y3) plus one
// inversion, and Montgomery's trick shares ONE inversion among a whole batch at 3 more
// multiplications per addition — 6 + (inversion / batch) instead of 10.  Round r replaces
// slots (2p, 2p+1) of its input array by their sum in slot p of its output array.  Bucket
// runs start at multiples of 2^R and are padded with the identity, so after R rounds every
// bucket is a run of (padded count >> R) affine points at (offset >> R), which
// accumulate_kernel<kReduced> then finishes in XYZZ.  No bucket structure is needed inside
// a round: pairs never straddle buckets by construction.
//
// One thread owns the batch {t, t + T, t + 2T, ...} (T = threads of the grid; neighbouring
// lanes touch neighbouring slots, so all streaming accesses are coalesced):
//   forward   prefix_j = d_0 ... d_(j-1) stored to scratch, d_j = the pair's denominator
//   invert    1 / (d_0 ... d_(B-1)) by Fermat, once per thread
//   backward  1/d_j = inv * prefix_j, inv *= d_j, then the addition itself
// Denominators: x2 - x1 (addition), 2 y1 (doubling, x1 == x2 and y1 == y2 != 0), 1 when
// there is nothing to divide (an operand is the identity, or P + (-P)).  Same group results
// as the XYZZ path's case analysis (point_xyzz_impl.h:114-176); the affine identity is (0, 0)
// (affine_point.h:125).
// ---------------------------------------------------------------------------
constexpr int kPairThreads = 128;
constexpr uint32_t kPairBatch = 512;  // pairs per thread the grid is sized for

enum : uint32_t { kPairAdd = 0, kPairDouble = 1, kPairCopy1 = 2, kPairCopy2 = 3, kPairZero = 4 };

// Loads pair p.  kFirst: slots hold point index | sign << 31 into `bases` (kNoEntry: identity).
template <class C, bool kFirst>
TB_DEV void pair_load(Affine<typename C::Field>& p1, Affine<typename C::Field>& p2, uint32_t p,
                      const uint32_t* __restrict__ bases, const uint32_t* __restrict__ in) {
  using K = typename C::Field;
  constexpr int kAffineWords = 2 * K::kWords;
  if (kFirst) {
    uint2 e = reinterpret_cast<const uint2*>(in)[p];
    if (e.x == kNoEntry) {
      K::set_zero(p1.x);
      K::set_zero(p1.y);
    } else {
      affine_load<K>(p1, bases + (size_t)(e.x & 0x7fffffffu) * kAffineWords);
      K::cneg(p1.y, p1.y, e.x >> 31);
    }
    if (e.y == kNoEntry) {
      K::set_zero(p2.x);
      K::set_zero(p2.y);
    } else {
      affine_load<K>(p2, bases + (size_t)(e.y & 0x7fffffffu) * kAffineWords);
      K::cneg(p2.y, p2.y, e.y >> 31);
    }
  } else {
    affine_load<K>(p1, in + (size_t)(2 * (size_t)p) * kAffineWords);
    affine_load<K>(p2, in + (size_t)(2 * (size_t)p + 1) * kAffineWords);
  }
}

// Case of the pair and its denominator d (never zero).
template <class K>
TB_DEV uint32_t pair_denominator(typename K::El& d, const Affine<K>& p1, const Affine<K>& p2) {
  bool z1 = affine_is_zero<K>(p1), z2 = affine_is_zero<K>(p2);
  K::sub(d, p2.x, p1.x);
  uint32_t kind = kPairAdd;
  if (z1 || z2) {
    kind = z1 ? (z2 ? kPairZero : kPairCopy2) : kPairCopy1;
  } else if (K::is_zero(d)) {
    if (K::eq(p1.y, p2.y) && !K::is_zero(p1.y)) {
      kind = kPairDouble;
      K::dbl(d, p1.y);
    } else {
      kind = kPairZero;
    }
  }
  if (kind >= kPairCopy1) K::set_one(d);
  return kind;
}

template <class C, bool kFirst>
__global__ void __launch_bounds__(kPairThreads, AccMinBlocks<C>()) pair_round_kernel(
    const uint32_t* __restrict__ bases, const uint32_t* __restrict__ in,
    const MsmTotals* __restrict__ totals, uint32_t round, uint32_t* __restrict__ prefix,
    uint32_t* __restrict__ out) {
  using K = typename C::Field;
  constexpr int N = K::kWords;
  constexpr int kAffineWords = 2 * N;
  const uint32_t pairs = (totals->entries >> round) >> 1;  // entries is a multiple of 2^R
  const uint32_t T = gridDim.x * blockDim.x;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pairs) return;
  const uint32_t B = (pairs - t + T - 1) / T;  // this thread's batch

  // ---- forward: prefix products of the denominators ---------------------------------
  typename K::El acc;
  K::set_one(acc);
  {
    Affine<K> n1, n2;
    pair_load<C, kFirst>(n1, n2, t, bases, in);
    for (uint32_t j = 0; j < B; ++j) {
      Affine<K> p1 = n1, p2 = n2;
      if (j + 1 < B) pair_load<C, kFirst>(n1, n2, t + (j + 1) * T, bases, in);
      typename K::El d;
      pair_denominator<K>(d, p1, p2);
      K::store(prefix + ((size_t)j * T + t) * N, acc);
      K::mul(acc, acc, d);
    }
  }
  typename K::El inv;
  K::inv(inv, acc);

  // ---- backward: the additions --------------------------------------------------------
  Affine<K> n1, n2;
  typename K::El npre;
  pair_load<C, kFirst>(n1, n2, t + (B - 1) * T, bases, in);
  K::load_rw(npre, prefix + ((size_t)(B - 1) * T + t) * N);
  for (uint32_t j = B; j-- > 0;) {
    Affine<K> p1 = n1, p2 = n2;
    typename K::El pre = npre;
    if (j > 0) {
      pair_load<C, kFirst>(n1, n2, t + (j - 1) * T, bases, in);
      K::load_rw(npre, prefix + ((size_t)(j - 1) * T + t) * N);
    }
    typename K::El d, dinv, num, lam, x3, y3, tt;
    uint32_t kind = pair_denominator<K>(d, p1, p2);
    K::mul(dinv, inv, pre);  // 1 / d_j
    K::mul(inv, inv, d);     // inverse of the remaining prefix
    K::sub(num, p2.y, p1.y);
    if (kind == kPairDouble) {   // lambda = 3 x1^2 / (2 y1)
      K::sqr(tt, p1.x);
      K::dbl(num, tt);
      K::add(num, num, tt);
    }
    K::mul(lam, num, dinv);
    K::sqr(x3, lam);         // x3 = lambda^2 - x1 - x2
    K::sub(x3, x3, p1.x);
    K::sub(x3, x3, p2.x);
    K::sub(tt, p1.x, x3);    // y3 = lambda (x1 - x3) - y1
    K::mul(y3, lam, tt);
    K::sub(y3, y3, p1.y);
    Affine<K> r, cp;
    typename K::El zero;
    K::set_zero(zero);
    K::select(cp.x, kind == kPairCopy1, p1.x, p2.x);  // the operand that is not the identity
    K::select(cp.y, kind == kPairCopy1, p1.y, p2.y);
    K::select(cp.x, kind == kPairZero, zero, cp.x);
    K::select(cp.y, kind == kPairZero, zero, cp.y);
    K::select(r.x, kind <= kPairDouble, x3, cp.x);
    K::select(r.y, kind <= kPairDouble, y3, cp.y);
    uint32_t* dst = out + (size_t)(t + j * T) * kAffineWords;
    K::store(dst, r.x);
    K::store(dst + N, r.y);
  }
}

// Buckets split over several tasks: their partial sums (the first of which already carries
// the bucket's previous value) are folded into `state` in two stages, so that a bucket holding
// a large share of all entries (witness vectors: every scalar equal to 1 lands in one bucket)
// is folded by many CTAs instead of one.
//   stage A  one CTA per job {bucket, chunk}: tree-sum of <= kFoldThreads partials; the result
//            goes to `state` when the bucket has a single chunk, else in place to the chunk's
//            first slot
//   stage B  one CTA per bucket with several chunks: sums the chunk results into `state`
template <class K>
TB_DEV void fold_tree(XYZZ<K>& acc, uint32_t live_count, uint32_t* sh) {
  constexpr int kXyzzWords = 4 * K::kWords;
  XYZZ<K> tmp;
  int live = live_count < (uint32_t)kFoldThreads ? (int)live_count : kFoldThreads;
  int top = 1;
  while (top < live) top <<= 1;
  for (int stride = top / 2; stride >= 1; stride >>= 1) {
    if ((int)threadIdx.x >= stride && (int)threadIdx.x < 2 * stride)
      xyzz_store<K>(sh + (threadIdx.x - stride) * kXyzzWords, acc);
    __syncthreads();
    if ((int)threadIdx.x < stride) {
      xyzz_load<K>(tmp, sh + threadIdx.x * kXyzzWords);
      xyzz_add<K>(acc, tmp);
    }
    __syncthreads();
  }
}

template <class C>
__global__ void __launch_bounds__(kFoldThreads) fold_stage_a_kernel(
    const uint2* __restrict__ fold_jobs, const MsmTotals* __restrict__ totals,
    const uint32_t* __restrict__ offset, const uint32_t* __restrict__ task_base, uint32_t R,
    uint32_t part, uint32_t split_key, uint32_t* __restrict__ task_out,
    uint32_t* __restrict__ state) {
  using K = typename C::Field;
  constexpr int kXyzzWords = 4 * K::kWords;
  __shared__ uint32_t sh[kFoldThreads / 2 * kXyzzWords];
  const uint32_t seg = totals->seg;
  for (uint32_t job = blockIdx.x; job < totals->fold_jobs; job += gridDim.x) {
    uint2 jb = fold_jobs[job];
    uint32_t key = jb.x;
    if (part != kPartAll && (key < split_key) != (part == kPartLow)) continue;  // other group
    uint32_t cnt = (offset[key + 1] - offset[key]) >> R;  // padded run after the pair rounds
    uint32_t t = (cnt + seg - 1) / seg;
    uint32_t first = jb.y * kFoldThreads;
    uint32_t here = min((uint32_t)kFoldThreads, t - first);
    uint32_t* slots = task_out + (size_t)(task_base[key] + first) * kXyzzWords;
    XYZZ<K> acc;
    xyzz_set_zero<K>(acc);
    if (threadIdx.x < here) xyzz_load<K>(acc, slots + (size_t)threadIdx.x * kXyzzWords);
    __syncthreads();  // every partial of the chunk is in registers before slot 0 is rewritten
    fold_tree<K>(acc, here, sh);
    if (threadIdx.x == 0)
      xyzz_store<K>(t <= (uint32_t)kFoldThreads ? state + (size_t)key * kXyzzWords : slots, acc);
    __syncthreads();
  }
}

template <class C>
__global__ void __launch_bounds__(kFoldThreads) fold_stage_b_kernel(
    const uint32_t* __restrict__ multi2_keys, const MsmTotals* __restrict__ totals,
    const uint32_t* __restrict__ offset, const uint32_t* __restrict__ task_base, uint32_t R,
    uint32_t part, uint32_t split_key, const uint32_t* __restrict__ task_out,
    uint32_t* __restrict__ state) {
  using K = typename C::Field;
  constexpr int kXyzzWords = 4 * K::kWords;
  __shared__ uint32_t sh[kFoldThreads / 2 * kXyzzWords];
  const uint32_t seg = totals->seg;
  for (uint32_t m = blockIdx.x; m < totals->multi2; m += gridDim.x) {
    uint32_t key = multi2_keys[m];
    if (part != kPartAll && (key < split_key) != (part == kPartLow)) continue;  // other group
    uint32_t cnt = (offset[key + 1] - offset[key]) >> R;
    uint32_t t = (cnt + seg - 1) / seg;
    uint32_t chunks = (t + kFoldThreads - 1) / kFoldThreads;
    const uint32_t* slots = task_out + (size_t)task_base[key] * kXyzzWords;
    XYZZ<K> acc, tmp;
    xyzz_set_zero<K>(acc);
    for (uint32_t j = threadIdx.x; j < chunks; j += kFoldThreads) {
      xyzz_load<K>(tmp, slots + (size_t)j * kFoldThreads * kXyzzWords);
      xyzz_add<K>(acc, tmp);
    }
    fold_tree<K>(acc, chunks, sh);
    if (threadIdx.x == 0) xyzz_store<K>(state + (size_t)key * kXyzzWords, acc);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Bucket reduction (pippenger_base.h:36-57): S_w = sum_k (k+1) B_k.
// Blocked running sums: with F(x) = sum_k k x_k and A(x) = sum_k x_k over
// blocks t of L items,  F(x) = sum_t Wt_t + L * F({A_t}).  Every level maps
// n items to ceil(n / L) items and carries the already-weighted part in a
// second series C (scaled by 2^shift = product of the previous Ls), so after
// the last level S_w = A + C.
//   level 0: items are the bucket values in `state` (zz == 0: empty), no C input
//   level>0: items are the previous level's A and C arrays
// ---------------------------------------------------------------------------
constexpr int kReduceThreads = 128;

template <class C, bool kFirst>
__global__ void __launch_bounds__(kReduceThreads) reduce_level_kernel(
    const uint32_t* __restrict__ in_a, const uint32_t* __restrict__ in_c, uint32_t n_in,
    uint32_t n_out, uint32_t L, uint32_t shift, uint32_t windows, uint32_t* __restrict__ out_a,
    uint32_t* __restrict__ out_c) {
  using K = typename C::Field;
  constexpr int kXyzzWords = 4 * K::kWords;
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_out * windows) return;
  uint32_t w = g / n_out, t = g % n_out;
  uint32_t lo = t * L, hi = min(n_in, lo + L);
  XYZZ<K> run, wt, csum, item;
  xyzz_set_zero<K>(run);
  xyzz_set_zero<K>(wt);
  xyzz_set_zero<K>(csum);
  for (uint32_t k = hi; k-- > lo;) {
    uint32_t idx = w * n_in + k;
    if (kFirst) {
      xyzz_load<K>(item, in_a + (size_t)idx * kXyzzWords);
      xyzz_add<K>(run, item);
    } else {
      xyzz_load<K>(item, in_a + (size_t)idx * kXyzzWords);
      xyzz_add<K>(run, item);
      xyzz_load<K>(item, in_c + (size_t)idx * kXyzzWords);
      xyzz_add<K>(csum, item);
    }
    if (k > lo) xyzz_add<K>(wt, run);
  }
  for (uint32_t s = 0; s < shift; ++s) xyzz_dbl<K>(wt);
  xyzz_add<K>(csum, wt);
  // leaf node g of the reduction tree: (A, P) side by side (out_c = out_a + one point)
  xyzz_store<K>(out_a + (size_t)g * 2 * kXyzzWords, run);
  xyzz_store<K>(out_c + (size_t)g * 2 * kXyzzWords, csum);
}

// Level 0 of the bucket reduction with the two dependent additions of a running sum split over
// TWO threads.  For a block of L bucket values (walked from the top) one thread keeps
// run_k = run_(k+1) + B_k, the other wt += run_k; wt lags one step behind, so both additions of
// a step proceed at the same time: the chain per block is L additions deep instead of 2 L - 1,
// twice as many warps are in flight for the same work, and each thread holds two points instead
// of three (the single-thread form spilled: 128 registers + 384 B of stack).  Roles are per
// warp (no divergence); run values cross through a double-buffered shared-memory slot per
// block, one __syncthreads per step.  Writes A = sum B_k to out_a and Wt = sum_k (k - lo) B_k
// to out_c, like reduce_level_kernel<C, true> with shift = 0: block g is leaf node g of the
// reduction tree, its two values (A, P) stored side by side (out_c = out_a + one point).  The
// caller passes pointers offset to the first window of the group it reduces and `wide` relative
// to that window.
//   wide / windows: blocks in the upper half of a narrow (c - 1 bit) window hold no entries by
//   construction (balanced windows); a CTA that lies wholly inside one writes zeros and leaves.
template <class C>
constexpr int ReduceSlots() {
  return 4 * C::Field::kWords > 64 ? 32 : 64;
}

// kInline: both roles share ONE call site of the addition, inlined there (no out-of-line call, no
// operands through the stack).
template <class C, class K = typename C::Field, bool kInline = false>
__global__ void __launch_bounds__(2 * ReduceSlots<C>(), K::kWords <= 12 ? 4 : 0) reduce_blocks_kernel(
    const uint32_t* __restrict__ state, uint32_t n_in, uint32_t n_out, uint32_t L,
    uint32_t windows, uint32_t wide, uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_c) {
  constexpr int kXyzzWords = 4 * K::kWords;
  constexpr int kSlots = kXyzzWords > 64 ? 32 : 64;  // = ReduceSlots<C>()
  constexpr int kStride = kXyzzWords + 4;  // 16-byte accesses of neighbouring slots hit distinct banks
  __shared__ __align__(16) uint32_t sh[2 * kSlots * kStride];
  const uint32_t slot = threadIdx.x % kSlots;
  const bool summing = threadIdx.x >= (uint32_t)kSlots;  // role B: wt += run
  const uint32_t total = n_out * windows;
  const uint32_t g0 = blockIdx.x * kSlots;
  const uint32_t g = g0 + slot;
  {
    // known-empty CTA (uniform decision: first and last block of the CTA)
    uint32_t g1 = min(g0 + kSlots, total) - 1;
    uint32_t w0 = g0 / n_out, w1 = g1 / n_out;
    if (w0 == w1 && w0 >= wide && (g0 % n_out) * L >= n_in / 2) {
      if (g < total) {
        XYZZ<K> z;
        xyzz_set_zero<K>(z);
        xyzz_store<K>((summing ? out_c : out_a) + (size_t)g * 2 * kXyzzWords, z);
      }
      return;
    }
  }
  const bool live = g < total;
  const uint32_t w = live ? g / n_out : 0, t = live ? g % n_out : 0;
  const uint32_t lo = t * L, hi = min(n_in, lo + L);
  const uint32_t steps = live ? hi - lo : 0;
  XYZZ<K> acc, in;
  xyzz_set_zero<K>(acc);
  uint32_t* mine = sh + slot * kStride;
  // step i: role A adds bucket hi - 1 - i and publishes run in buffer i & 1; role B adds what
  // was published in step i - 1 (all but the run that includes the block's lowest bucket)
  for (uint32_t i = 0; i < L; ++i) {
    if (kInline) {
      const bool have = summing ? (i >= 1 && i < steps) : (i < steps);
      if (have) {
        const uint32_t* src = summing ? mine + ((i - 1) & 1) * (kSlots * kStride)
                                      : state + (size_t)(w * n_in + (hi - 1 - i)) * kXyzzWords;
        xyzz_load<K>(in, src);
        xyzz_add_inlined<K>(acc, in);
        if (!summing && i + 1 < steps) xyzz_store<K>(mine + (i & 1) * (kSlots * kStride), acc);
      }
    } else if (!summing) {
      if (i < steps) {
        uint32_t idx = w * n_in + (hi - 1 - i);
        xyzz_load<K>(in, state + (size_t)idx * kXyzzWords);
        xyzz_add<K>(acc, in);
        if (i + 1 < steps) xyzz_store<K>(mine + (i & 1) * (kSlots * kStride), acc);
      }
    } else {
      if (i >= 1 && i < steps) {
        xyzz_load<K>(in, mine + ((i - 1) & 1) * (kSlots * kStride));
        xyzz_add<K>(acc, in);
      }
    }
    __syncthreads();
  }
  if (live) xyzz_store<K>((summing ? out_c : out_a) + (size_t)g * 2 * kXyzzWords, acc);
}

// The same running-sum level for G2 with every block handled by a LANE PAIR per role (a lane per
// Fq2 component, xyzz.cuh PairOps::add): kSlots blocks per CTA, 2 x 2 x kSlots threads — warps
// [0, kSlots / 16) keep the running sums, the others the weighted sums.  Warps are role-uniform
// and in step by the per-step barrier, so every shuffle is a full-mask one.  Same values, same
// memory layout as reduce_blocks_kernel.
constexpr int kPairReduceSlots = 32;

template <class C, int kRoll, int kMinBlocks>
__global__ void __launch_bounds__(4 * kPairReduceSlots, kMinBlocks) reduce_blocks_pair_kernel(
    const uint32_t* __restrict__ state, uint32_t n_in, uint32_t n_out, uint32_t L,
    uint32_t windows, uint32_t wide, uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_c) {
  using F = typename C::Fq;
  static_assert(C::Field::kDegree == 2, "lane pairs split a quadratic extension");
  using Ops = PairOps<F, kRoll>;
  constexpr int N = Fp<F>::N;
  constexpr int kXyzzWords = 8 * N;
  constexpr int kSlots = kPairReduceSlots;
  constexpr int kStride = kXyzzWords + 4;
  __shared__ __align__(16) uint32_t sh[2 * kSlots * kStride];
  const uint32_t role = threadIdx.x & 1;
  const uint32_t pair = threadIdx.x >> 1;
  const uint32_t slot = pair % kSlots;
  const bool summing = pair >= (uint32_t)kSlots;  // second half of the CTA: wt += run
  const uint32_t total = n_out * windows;
  const uint32_t g0 = blockIdx.x * kSlots;
  const uint32_t g = g0 + slot;
  auto load = [&](PairPoint<F>& p, const uint32_t* src) {
    fp_load_rw<F>(p.x, src + role * N);
    fp_load_rw<F>(p.y, src + 2 * N + role * N);
    fp_load_rw<F>(p.zz, src + 4 * N + role * N);
    fp_load_rw<F>(p.zzz, src + 6 * N + role * N);
  };
  auto store = [&](uint32_t* dst, const PairPoint<F>& p) {
    fp_store<F>(dst + role * N, p.x);
    fp_store<F>(dst + 2 * N + role * N, p.y);
    fp_store<F>(dst + 4 * N + role * N, p.zz);
    fp_store<F>(dst + 6 * N + role * N, p.zzz);
  };
  PairPoint<F> acc, in;
  Fp2Lanes<F>::set_one(acc.x, role);
  Fp2Lanes<F>::set_one(acc.y, role);
  fp_set_zero<F>(acc.zz);
  fp_set_zero<F>(acc.zzz);
  {
    // known-empty CTA (uniform decision: first and last block of the CTA)
    uint32_t g1 = min(g0 + kSlots, total) - 1;
    uint32_t w0 = g0 / n_out, w1 = g1 / n_out;
    if (w0 == w1 && w0 >= wide && (g0 % n_out) * L >= n_in / 2) {
      if (g < total) store((summing ? out_c : out_a) + (size_t)g * 2 * kXyzzWords, acc);
      return;
    }
  }
  const bool live = g < total;
  const uint32_t w = live ? g / n_out : 0, t = live ? g % n_out : 0;
  const uint32_t lo = t * L, hi = min(n_in, lo + L);
  const uint32_t steps = live ? hi - lo : 0;
  uint32_t* mine = sh + slot * kStride;
  in = acc;
  for (uint32_t i = 0; i < L; ++i) {
    if (!summing) {  // warp-uniform
      const bool active = i < steps;
      if (active) load(in, state + (size_t)(w * n_in + (hi - 1 - i)) * kXyzzWords);
      Ops::add(acc, in, active, role);
      if (i + 1 < steps) store(mine + (i & 1) * (kSlots * kStride), acc);
    } else {
      const bool active = i >= 1 && i < steps;
      if (active) load(in, mine + ((i - 1) & 1) * (kSlots * kStride));
      Ops::add(acc, in, active, role);
    }
    __syncthreads();
  }
  if (live) store((summing ? out_c : out_a) + (size_t)g * 2 * kXyzzWords, acc);
}

// Tail of the bucket reduction.  After level 0 every window has nb = 2^M blocks t with
// (A_t, P_t) and  S_w = sum_t A_t + sum_t P_t + L * sum_t t A_t.  Writing t in binary,
// sum_t t A_t = sum_j 2^j D_j with D_j = sum of A_t over the t whose bit j is set, so only
// PLAIN sums remain and the powers of two are applied once, by window_combine_kernel, which
// doubles anyway.  A binary tree carries per node the vector (A, P, D_0 .. D_(s-1)) of its 2^s
// blocks; a merge is  A = A_l + A_r, P = P_l + P_r, D_i = D_i,l + D_i,r, D_s = A_r — all
// independent, one thread each — so every level is ONE addition deep.
//
// reduce_tree_kernel runs `levels` levels of that tree in ONE launch: a CTA owns 2^levels
// consecutive input nodes of one window (vin values each) and ping-pongs between two private
// scratch regions in global memory (L1/L2 resident; a __syncthreads per level), so a 12-14
// level tree is two launches instead of 12-14 latency-bound ones.  Levels with at least
// `thread_items` work items use one thread per addition (the throughput form, chosen by the host
// when the launch has several CTAs per SM slot), the others four lanes per addition.  The stage that ends with one
// node per window (`fin.enabled`) scatters that node's values to the window's bit positions of
// the term array Y consumed by window_combine_kernel:
//   Y[off_w]          = A + P
//   Y[off_w + l0 + j] = D_j          (weight L = 2^l0 of the blocked running sums)
//   other positions of the window    = identity
struct TreeFinal {
  uint32_t enabled;
  uint32_t w_begin;  // global index of the launch's first window
  uint32_t c, wide;  // window widths: c bits for w < wide, c - 1 above (balanced windows)
  uint32_t l0;       // log2 of the level-0 block length
  uint32_t* terms;   // Y, one XYZZ per bit position
};
constexpr int kTreeThreads = 256;
constexpr uint32_t kTreeStageLevels = 7;  // most levels per launch: 128 input nodes per CTA
constexpr uint32_t kTreeLastLevels = 5;   // levels of the final, one-CTA-per-window stage

TB_DEV uint32_t window_bit_offset(uint32_t w, uint32_t c, uint32_t wide) {
  return w * c - (w > wide ? w - wide : 0u);
}

// (8-limb curve: held to 2 CTAs per SM — left alone, ptxas has chosen anything from 113 to 170
// registers for this kernel from build to build)
template <class C>
__global__ void __launch_bounds__(kTreeThreads, C::Field::kWords == 8 ? 2 : 0) reduce_tree_kernel(
    const uint32_t* in, uint32_t vin, uint32_t levels, uint32_t ctas_per_window,
    size_t slice_words, uint32_t* ping, uint32_t* pong, uint32_t* out, TreeFinal fin,
    uint32_t thread_items) {
  using K = typename C::Field;
  constexpr int kXyzzWords = 4 * K::kWords;
  // every buffer is cut into per-window slices of slice_words; inside its window's slice a CTA
  // owns the nodes [chunk * 2^levels, (chunk + 1) * 2^levels) of vin values each
  const uint32_t n0 = 1u << levels;
  const uint32_t lw = blockIdx.x / ctas_per_window, chunk = blockIdx.x % ctas_per_window;
  const size_t region = (size_t)lw * slice_words + (size_t)chunk * n0 * vin * kXyzzWords;
  const uint32_t* src = in + region;
  uint32_t* bufs[2] = {ping + region, pong + region};
  const uint32_t vfinal = vin + levels;
  const uint32_t group = threadIdx.x >> 2, lane = threadIdx.x & 3, mask = 0xfu << (threadIdx.x & 28);
  for (uint32_t l = 1; l <= levels; ++l) {
    const uint32_t m = n0 >> l, vout = vin + l, vprev = vout - 1;
    // the last level of a stage that is not the final one writes the compact output array
    uint32_t* dst = (l == levels && !fin.enabled)
                        ? out + (size_t)lw * slice_words + (size_t)chunk * vfinal * kXyzzWords
                        : bufs[l & 1];
    if (m * vout >= thread_items) {
      // throughput form (many CTAs per SM, wide level): one work item per thread, plain additions
      // — 14 multiplications per thread instead of 16 lane-multiplications plus shuffles
      for (uint32_t it = threadIdx.x; it < m * vout; it += kTreeThreads) {
        const uint32_t node = it / vout, v = it % vout;
        const uint32_t* left = src + (size_t)(2 * node) * vprev * kXyzzWords;
        const uint32_t* right = left + (size_t)vprev * kXyzzWords;
        XYZZ<K> a;
        if (v == vout - 1) {
          xyzz_load<K>(a, right);
        } else {
          XYZZ<K> b;
          xyzz_load<K>(a, left + (size_t)v * kXyzzWords);
          xyzz_load<K>(b, right + (size_t)v * kXyzzWords);
          xyzz_add<K>(a, b);
        }
        xyzz_store<K>(dst + (size_t)it * kXyzzWords, a);
      }
      __syncthreads();
      src = dst;
      continue;
    }
    // latency form: one work item (node, value) per group of four lanes (Coop4: an addition is 4
    // multiplication levels deep instead of 14 multiplications)
    for (uint32_t it = group; it < m * vout; it += kTreeThreads / 4) {
      const uint32_t node = it / vout, v = it % vout;
      const uint32_t* left = src + (size_t)(2 * node) * vprev * kXyzzWords;
      const uint32_t* right = left + (size_t)vprev * kXyzzWords;
      XYZZ<K> a;
      if (v == vout - 1) {
        xyzz_load<K>(a, right);  // D_new = A_r
      } else {
        XYZZ<K> b;
        xyzz_load<K>(a, left + (size_t)v * kXyzzWords);
        xyzz_load<K>(b, right + (size_t)v * kXyzzWords);
        Coop4<K>::add(a, b, lane, mask);
      }
      if (lane == 0) xyzz_store<K>(dst + (size_t)it * kXyzzWords, a);
    }
    __syncthreads();
    src = dst;
  }
  if (!fin.enabled) return;
  // one node per window left (ctas_per_window == 1): values (A, P, D_0 .. D_(M-1)) at src
  const uint32_t w = fin.w_begin + lw;
  const uint32_t cw = fin.c - (w >= fin.wide ? 1u : 0u);
  const uint32_t base = window_bit_offset(w, fin.c, fin.wide);
  const uint32_t M = vfinal - 2;
  for (uint32_t bit = group; bit < cw; bit += kTreeThreads / 4) {
    XYZZ<K> a;
    xyzz_set_zero<K>(a);
    if (bit == 0) {
      XYZZ<K> b;
      xyzz_load<K>(a, src);
      xyzz_load<K>(b, src + kXyzzWords);
      Coop4<K>::add(a, b, lane, mask);
    }
    if (bit >= fin.l0 && bit - fin.l0 < M) {
      XYZZ<K> d;
      xyzz_load<K>(d, src + (size_t)(2 + bit - fin.l0) * kXyzzWords);
      Coop4<K>::add(a, d, lane, mask);  // (l0 >= 1, so bit 0 never carries a D term; kept general)
    }
    if (lane == 0) xyzz_store<K>(fin.terms + (size_t)(base + bit) * kXyzzWords, a);
  }
}

// Window combination (pippenger_base.h:59-77 AccumulateWindowSums, the Horner over the window
// sums) on the device: the MSM value is  sum_b 2^b Y_b  over the bit-positioned terms Y the
// reduction tree left.  A pairing tree computes it in place: a segment's sum is its left half
// plus 2^(length of the left half) times its right half, so the doublings of different
// subtrees proceed in parallel and the term at bit b is doubled exactly b times in total — the
// c doublings per window of the reference's Horner, (W - 1) c in all on the longest path, plus
// log2 additions instead of one per term.  One CTA.  Positions [0, clear_below) are reset to the identity first (the low
// window group writes its own term array).  The result, plus *add_in when given (the high
// group's sum), goes to out.
//
// The doubling chain is latency-bound and is the one serial part of an MSM: 3.8 us per 254-bit
// doubling on one lane, 1.6 us with the four-lane form used here (tools/probe/chain_probe.cu;
// 381-bit: 8.0 -> 3.3 us, Fq2: 12.6 -> 4.4 and 46 -> 9.6 us), i.e. ~0.45 ms for the ~255
// doublings of a BN254 MSM.  The engine hides it behind the accumulation of the low windows.
constexpr int kCombineThreads = 128;

// per_window.c != 0: one CTA per window — CTA b combines only the bit positions of window
// per_window.w_begin + b and writes that window's sum S_w = sum_k (k + 1) B_k to out[b]
// (`count`, `clear_below`, `add_in` unused).  This is the parallel part of the window
// combination; what is left — sum_w 2^(offset of w) S_w, ~255 strictly sequential doublings of
// one point — is the ladder the engine runs on the host by default (msm_engine.cuh, Finish).
struct PerWindow {
  uint32_t c, wide, w_begin;
};

template <class C>
__global__ void __launch_bounds__(kCombineThreads) window_combine_kernel(
    uint32_t* terms, uint32_t count, uint32_t clear_below, const uint32_t* add_in,
    uint32_t* out, PerWindow per_window) {
  using K = typename C::Field;
  constexpr int kXyzzWords = 4 * K::kWords;
  constexpr uint32_t kGroups = kCombineThreads / 4;
  if (per_window.c) {
    const uint32_t w = per_window.w_begin + blockIdx.x;
    terms += (size_t)window_bit_offset(w, per_window.c, per_window.wide) * kXyzzWords;
    count = per_window.c - (w >= per_window.wide ? 1u : 0u);
    clear_below = 0;
    add_in = nullptr;
    out += (size_t)blockIdx.x * kXyzzWords;
  }
  // groups of four lanes share one (left, right) pair: Coop4 doublings and additions
  const uint32_t group = threadIdx.x >> 2, lane = threadIdx.x & 3, mask = 0xfu << (threadIdx.x & 28);
  for (uint32_t b = threadIdx.x; b < clear_below; b += kCombineThreads) {
    XYZZ<K> z;
    xyzz_set_zero<K>(z);
    xyzz_store<K>(terms + (size_t)b * kXyzzWords, z);
  }
  __syncthreads();
  // Balanced pairing: at depth d the positions are cut into 2^d segments with boundaries
  // b(d, i) = ceil(i * count / 2^d); segment (d, i) = its left child + 2^(shift) * its right
  // child, shift = length of the left child.  Bottom-up, in place (a segment's sum lives at its
  // first position), so the longest chain is ~count + log2(count) doublings for any count.
  uint32_t depth = 0;
  while ((1u << depth) < count) ++depth;
  for (uint32_t d = depth; d-- > 0;) {
    for (uint32_t i = group; i < (1u << d); i += kGroups) {
      const uint32_t lo = (uint32_t)(((uint64_t)i * count + (1u << d) - 1) >> d);
      const uint32_t mid = (uint32_t)(((uint64_t)(2 * i + 1) * count + (2u << d) - 1) >> (d + 1));
      const uint32_t hi = (uint32_t)(((uint64_t)(i + 1) * count + (1u << d) - 1) >> d);
      if (mid <= lo || mid >= hi) continue;  // a child is empty: nothing to merge
      XYZZ<K> t, a;
      xyzz_load<K>(t, terms + (size_t)mid * kXyzzWords);
      if (!xyzz_is_zero<K>(t)) {  // uniform inside the group
        for (uint32_t k = lo; k < mid; ++k) Coop4<K>::dbl_nz(t, lane, mask);
        xyzz_load<K>(a, terms + (size_t)lo * kXyzzWords);
        Coop4<K>::add(a, t, lane, mask);
        if (lane == 0) xyzz_store<K>(terms + (size_t)lo * kXyzzWords, a);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) {
    XYZZ<K> a;
    if (count) {
      xyzz_load<K>(a, terms);
    } else {
      xyzz_set_zero<K>(a);
    }
    if (add_in) {
      XYZZ<K> b;
      xyzz_load<K>(b, add_in);
      Coop4<K>::add(a, b, lane, mask);
    }
    if (lane == 0) xyzz_store<K>(out, a);
  }
}

// ---------------------------------------------------------------------------
// Element-wise hooks used by the parity tests (the role of
// tachyon/math/finite_fields/kernels/prime_field_ops.cu.h:13-51 and
// short_weierstrass/kernels/elliptic_curve_ops.cu.h:15-79 in the reference's
// own GPU correctness tests).
// ---------------------------------------------------------------------------
template <class F>
__global__ void field_op_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out,
                                uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int N = Fp<F>::N;
  Fp<F> x, y, r;
  fp_load<F>(x, a + (size_t)i * N);
  fp_load<F>(y, b + (size_t)i * N);
  switch (op) {
    case 0: fp_add<F>(r, x, y); break;
    case 1: fp_sub<F>(r, x, y); break;
    case 2: fp_mul<F>(r, x, y); break;
    case 3: fp_sqr<F>(r, x); break;
    case 4: fp_neg<F>(r, x); break;
    case 5: fp_dbl<F>(r, x); break;
    case 6: fp_inv<F>(r, x); break;
    case 7: fp_from_mont<F>(r, x); break;
    default: fp_to_mont<F>(r, x); break;
  }
  fp_store<F>(out + (size_t)i * N, r);
}

// The same hook for an extension field kind (Fq2): ops 0..6 as above.
template <class K>
__global__ void ext_field_op_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out,
                                    uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int N = K::kWords;
  typename K::El x, y, r;
  K::load(x, a + (size_t)i * N);
  K::load(y, b + (size_t)i * N);
  switch (op) {
    case 0: K::add(r, x, y); break;
    case 1: K::sub(r, x, y); break;
    case 2: K::mul(r, x, y); break;
    case 3: K::sqr(r, x); break;
    case 4: K::neg(r, x); break;
    case 5: K::dbl(r, x); break;
    default: K::inv(r, x); break;
  }
  K::store(out + (size_t)i * N, r);
}

// op 0: out = a + b (XYZZ + XYZZ); 1: out = a + affine b; 2: out = a - affine b;
// 3: out = 2a.  a, out: XYZZ arrays; b: XYZZ (op 0) or affine (op 1, 2).
template <class C>
__global__ void point_op_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out,
                                uint32_t n) {
  using K = typename C::Field;
  constexpr int N = K::kWords;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<K> p;
  xyzz_load<K>(p, a + (size_t)i * 4 * N);
  if (op == 0) {
    XYZZ<K> q;
    xyzz_load<K>(q, b + (size_t)i * 4 * N);
    xyzz_add<K>(p, q);
  } else if (op == 1 || op == 2) {
    Affine<K> q;
    affine_load<K>(q, b + (size_t)i * 2 * N);
    xyzz_madd<K>(p, q, op == 2);
  } else {
    xyzz_dbl<K>(p);
  }
  xyzz_store<K>(out + (size_t)i * 4 * N, p);
}

// Table of precomputed multiples for REGISTERED bases (the role of precompute_factor in
// algorithms/icicle/icicle_msm.h:21): T[w][i] = 2^(bit offset of window w) * P_i, affine.  One
// thread per point walks up the windows (c_w doublings each) and normalises every multiple
// (one field inversion each; this runs once per registration).  T[0] = the bases themselves.
template <class C>
__global__ void __launch_bounds__(128) precompute_table_kernel(const uint32_t* __restrict__ bases,
                                                               uint32_t n, uint32_t stride, uint32_t W,
                                                               uint32_t c, uint32_t wide,
                                                               uint32_t* __restrict__ table) {
  using K = typename C::Field;
  constexpr int kAffineWords = 2 * K::kWords;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<K> a;
  affine_load<K>(a, bases + (size_t)i * kAffineWords);
  affine_store<K>(table + (size_t)i * kAffineWords, a);
  XYZZ<K> p;
  xyzz_set_zero<K>(p);
  xyzz_madd<K>(p, a, false);
  for (uint32_t w = 1; w < W; ++w) {
    const uint32_t cw = c - (w - 1 >= wide ? 1u : 0u);  // width of the window below
    for (uint32_t k = 0; k < cw; ++k) xyzz_dbl<K>(p);
    if (xyzz_is_zero<K>(p)) {
      K::set_zero(a.x);
      K::set_zero(a.y);
    } else {
      typename K::El zi3, zi2;
      K::inv(zi3, p.zzz);
      K::mul(zi2, zi3, p.zz);
      K::sqr(zi2, zi2);
      K::mul(a.x, p.x, zi2);
      K::mul(a.y, p.y, zi3);
    }
    affine_store<K>(table + ((size_t)w * stride + i) * kAffineWords, a);
  }
}

// ---------------------------------------------------------------------------
// Synthetic inputs (mirrors elliptic_curves/test/random.h:11-29
// and big_int.h:107-116 with a fixed-seed counter-based SplitMix64): chains of
// 2^12 successive doublings, chain j starting at [h_j] G.
// ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64_at(uint64_t seed, uint64_t index) {
  uint64_t z = seed + (index + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr uint32_t kChainLog = 12;

// One thread per chain: writes the chain's points as XYZZ.
template <class C>
__global__ void generate_chains_kernel(uint64_t seed, uint32_t first_chain, uint32_t n_chains,
                                       uint32_t n_points, uint32_t* __restrict__ out_xyzz) {
  using K = typename C::Field;
  constexpr int N = K::kWords;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chains) return;
  uint64_t h = splitmix64_at(seed ^ 0x7074ull, first_chain + t) | 1ull;
  Affine<K> gen;
  K::template set_words<typename C::Gen>(gen.x, 0);
  K::template set_words<typename C::Gen>(gen.y, N);
  XYZZ<K> p;
  xyzz_set_zero<K>(p);
  for (int bit = 63; bit >= 0; --bit) {
    xyzz_dbl<K>(p);
    if ((h >> bit) & 1) xyzz_madd<K>(p, gen, false);
  }
  uint32_t chain_len = 1u << kChainLog;
  for (uint32_t d = 0; d < chain_len; ++d) {
    uint64_t i = (uint64_t)t * chain_len + d;
    if (i >= n_points) break;
    xyzz_store<K>(out_xyzz + i * 4 * N, p);
    xyzz_dbl<K>(p);
  }
}

// XYZZ -> affine (point_xyzz.h:199-213), one thread per point.
template <class C>
__global__ void normalize_kernel(const uint32_t* __restrict__ in_xyzz, uint32_t n,
                                 uint32_t* __restrict__ out_affine) {
  using K = typename C::Field;
  constexpr int N = K::kWords;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<K> p;
  xyzz_load<K>(p, in_xyzz + (size_t)i * 4 * N);
  Affine<K> a;
  if (xyzz_is_zero<K>(p)) {
    K::set_zero(a.x);
    K::set_zero(a.y);
  } else {
    typename K::El zi3, zi2;
    K::inv(zi3, p.zzz);
    K::mul(zi2, zi3, p.zz);
    K::sqr(zi2, zi2);
    K::mul(a.x, p.x, zi2);
    K::mul(a.y, p.y, zi3);
  }
  K::store(out_affine + (size_t)i * 2 * N, a.x);
  K::store(out_affine + (size_t)i * 2 * N + N, a.y);
}

// dist: 0 uniform, 1 non_uniform (one scalar repeated), 2 witness
template <class C>
__global__ void generate_scalars_kernel(uint64_t seed, int dist, uint64_t first, uint32_t n,
                                        uint32_t* __restrict__ out) {
  using Fr = typename C::Fr;
  constexpr int N = Fp<Fr>::N;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t i = first + k;
  Fp<Fr> x;
  bool random = true;
  uint64_t src = i;
  if (dist == 1) src = 0;
  if (dist == 2) {
    uint64_t sel = splitmix64_at(seed ^ 0x7769ull, i) % 10;
    if (sel < 7) {
      random = false;
      fp_set_zero<Fr>(x);
      if (sel >= 4) x.l[0] = 1;
    } else if (sel < 9) {
      random = false;
      fp_set_zero<Fr>(x);
      x.l[0] = (uint32_t)splitmix64_at(seed ^ 0x7363ull, i * (N / 2));
    }
  }
  if (random) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      uint64_t v = splitmix64_at(seed ^ 0x7363ull, src * (N / 2) + j);
      x.l[2 * j] = (uint32_t)v;
      x.l[2 * j + 1] = (uint32_t)(v >> 32);
    }
    // halve until < r   (big_int.h:107-116)
    for (;;) {
      uint32_t t = sub_cc(x.l[0], Fr::mod(0));
#pragma unroll
      for (int j = 1; j < N; ++j) t = subc_cc(x.l[j], Fr::mod(j));
      uint32_t borrow = subc(0u, 0u);
      (void)t;
      if (borrow) break;
#pragma unroll
      for (int j = 0; j < N - 1; ++j) x.l[j] = __funnelshift_r(x.l[j], x.l[j + 1], 1);
      x.l[N - 1] >>= 1;
    }
  }
  Fp<Fr> m;
  fp_to_mont<Fr>(m, x);
  fp_store<Fr>(out + (size_t)k * N, m);
}

// ---------------------------------------------------------------------------
// INT32 multiply-pipe peak (the roofline denominator of SURVEY.md §8d): rate of
// 32x32->64-bit multiply-adds.  Multiplicands are taken from neighbouring
// accumulators so that ptxas cannot hoist or strength-reduce the products (an
// earlier version with loop-invariant operands was turned into IADD3s and
// over-reported the peak by 1.8x; see DESIGN.md "pipe rates").
//   variant 0: mad.lo.cc / madc.hi.cc chains  -> IMAD.WIDE.U32(.X), carry in predicate
//   variant 1: mad.wide.u32 with 64-bit addend -> IMAD.WIDE.U32 (+ IADD3 as ptxas sees fit)
//   variant 2: mad.lo + mad.hi pairs           -> IMAD + IMAD.HI
// Each thread issues iters * 16 products.
// ---------------------------------------------------------------------------
template <int kVariant>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t iters, uint32_t seed,
                                                        uint32_t* __restrict__ out) {
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint32_t lo[16], hi[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo[k] = k + a;
    hi[k] = k ^ b;
  }
  for (uint32_t it = 0; it < iters; ++it) {
    if (kVariant == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        lo[4 * g] = mad_lo_cc(lo[(4 * g + 7) & 15], b, lo[4 * g]);
        hi[4 * g] = madc_hi_cc(lo[(4 * g + 7) & 15], b, hi[4 * g]);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          uint32_t m = lo[(4 * g + k + 7) & 15];
          lo[4 * g + k] = madc_lo_cc(m, b, lo[4 * g + k]);
          hi[4 * g + k] = madc_hi_cc(m, b, hi[4 * g + k]);
        }
      }
    } else if (kVariant == 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        asm volatile(
            "{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
            : "+r"(lo[k]), "+r"(hi[k])
            : "r"(lo[(k + 7) & 15]), "r"(b));
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        uint32_t m = lo[(k + 7) & 15];
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[k]) : "r"(m), "r"(b));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(m), "r"(b));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s ^= lo[k] ^ hi[k];
  if (s == 0x12345678u) out[0] = s;
}

}  // namespace tb200
