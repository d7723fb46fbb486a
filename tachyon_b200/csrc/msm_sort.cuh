// Two-level bucket sort of the (bucket, point) entries of one point range.
//
// The one-level counting sort of msm_kernels.cuh (digits_hist / digits_scatter) spends one
// global atomic per entry twice — 2 x 218 M at BN254 2^24, the second kind returning a value
// and followed by a 4-byte store somewhere in a 64 MB slice — and is bound by the L2 atomic
// unit (ncu: 137 long-scoreboard stalls per issued instruction, 1.9 GB of DRAM writes for
// 0.87 GB of output).  Here the bucket key is split into a coarse part (key >> 10) and a fine
// part (10 bits), and both passes count in SHARED memory per tile, touching global counters
// once per (tile, non-empty bin):
//
//   digits_coarse_hist   scalar -> canonical -> signed digits (written window-major) and, per
//                        tile of 4096 scalars, a shared histogram over (window, coarse bin)
//   coarse_scan          exclusive scan over the W * Cw coarse regions; tiles per region
//   coarse_scatter       per (tile, window): reserve a run per coarse bin with ONE atomic,
//                        place {index | sign, key} entries (8 B) by shared-memory rank
//   fine_hist            per tile of 4096 entries of ONE region: shared histogram over the 1024
//                        fine bins -> bucket sizes (global reduction per non-empty bin)
//   (scan + task building of msm_kernels.cuh, unchanged: bucket offsets, cursors, tasks)
//   fine_scatter         per tile: reserve a run per fine bin, write index | sign (4 B) into
//                        the region's 128 KB slice of `sorted`
//
// Skewed scalars (one bucket holding everything) make tiles with a single non-empty bin: one
// atomic per tile instead of 4096 serialised ones.  Same output contract as the one-level
// sort: `count`, then `sorted` grouped by bucket key through `cursor`.
#pragma once
#include "msm_kernels.cuh"

namespace tb200 {

constexpr int kSortThreads = 256;
constexpr int kSortPerThread = 16;
constexpr int kSortTile = kSortThreads * kSortPerThread;  // 4096
constexpr uint32_t kFineBits = 10;
constexpr uint32_t kFineBins = 1u << kFineBits;
constexpr uint32_t kMaxCoarsePerWindow = 1024;  // c <= 21; the tile staging keeps the CTA under 48 KB

struct SortPlan {
  uint32_t Cw;       // coarse bins per window = B >> kFineBits
  uint32_t regions;  // W * Cw
  uint32_t max_tiles;
};

struct SortTotals {
  uint32_t tiles;
  uint32_t pad[3];
};

// Shared-memory counter increment returning the previous value; lanes of a warp that hit the
// same counter (repeated scalars) are combined into one atomic when a cheap neighbour test
// sees a duplicate, as bucket_inc does for global counters.
TB_DEV uint32_t shared_inc(uint32_t* counters, uint32_t idx, bool valid) {
  uint32_t other = __shfl_xor_sync(0xffffffffu, valid ? idx : 0xffffffffu, 1);
  bool dup = __any_sync(0xffffffffu, valid && other == idx);
  uint32_t result = 0;
  if (!dup) {
    if (valid) result = atomicAdd(counters + idx, 1u);
    return result;
  }
  uint32_t vote = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    uint32_t peers = __match_any_sync(vote, idx);
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(peers) - 1;
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counters + idx, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    result = base + rank;
  }
  return result;
}

// In-place exclusive scan of `nbins` (<= 8 * kSortThreads) shared counters by the CTA.
TB_DEV void block_exclusive_scan_bins(uint32_t* bins, uint32_t nbins, uint32_t* warp_sums) {
  const uint32_t per = (nbins + kSortThreads - 1) / kSortThreads;
  const uint32_t lo = threadIdx.x * per, hi = min(nbins, lo + per);
  uint32_t sum = 0;
  for (uint32_t k = lo; k < hi; ++k) sum += bins[k];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = lane < kSortThreads / 32 ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= (uint32_t)o) v += y;
    }
    if (lane < kSortThreads / 32) warp_sums[lane] = v;
  }
  __syncthreads();
  uint32_t run = (warp ? warp_sums[warp - 1] : 0u) + x - sum;
  for (uint32_t k = lo; k < hi; ++k) {
    uint32_t c = bins[k];
    bins[k] = run;
    run += c;
  }
  __syncthreads();
}

// dynamic shared memory: W * Cw counters
template <class C>
__global__ void __launch_bounds__(kSortThreads, 6) digits_coarse_hist_kernel(
    const uint32_t* __restrict__ scalars, MsmPlan plan, SortPlan sp,
    uint32_t* __restrict__ digits, uint32_t* __restrict__ coarse_count) {
  using Fr = typename C::Fr;
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t k = threadIdx.x; k < sp.regions; k += kSortThreads) sh_hist[k] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t i = base + k * kSortThreads + threadIdx.x;
    bool in = i < plan.n;
    Fp<Fr> s;
    if (in) {
      load_scalar_canonical<Fr>(s, scalars, i);
    } else {
      fp_set_zero<Fr>(s);
    }
    for_each_digit<Fr>(s, plan, [&](uint32_t w, uint32_t mag, bool neg) {
      if (in) digits[(size_t)w * plan.n + i] = mag | (neg ? 0x80000000u : 0u);
      shared_inc(sh_hist, w * sp.Cw + ((mag - 1) >> kFineBits), in && mag != 0);
    });
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < sp.regions; k += kSortThreads)
    if (sh_hist[k]) atomicAdd(coarse_count + k, sh_hist[k]);
}

// One CTA.  coarse_offset[r], coarse_cursor[r] = entries before region r; tile_offset[r] = tiles
// before region r; totals->tiles.
__global__ void __launch_bounds__(1024) coarse_scan_kernel(
    const uint32_t* __restrict__ coarse_count, uint32_t regions,
    uint32_t* __restrict__ coarse_offset, uint32_t* __restrict__ coarse_cursor,
    uint32_t* __restrict__ tile_offset, SortTotals* __restrict__ totals,
    uint32_t* __restrict__ nonzero_slots);
#ifdef TB200_DEFINE_SORT_KERNELS
__global__ void __launch_bounds__(1024) coarse_scan_kernel(
    const uint32_t* __restrict__ coarse_count, uint32_t regions,
    uint32_t* __restrict__ coarse_offset, uint32_t* __restrict__ coarse_cursor,
    uint32_t* __restrict__ tile_offset, SortTotals* __restrict__ totals,
    uint32_t* __restrict__ nonzero_slots){
  __shared__ uint32_t warp_e[32], warp_t[32];
  const uint32_t per = (regions + 1023) / 1024;
  const uint32_t lo = threadIdx.x * per, hi = min(regions, lo + per);
  uint32_t se = 0, st = 0;
  for (uint32_t r = lo; r < hi; ++r) {
    uint32_t c = coarse_count[r];
    se += c;
    st += (c + kSortTile - 1) / kSortTile;
  }
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t xe = se, xt = st;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t ye = __shfl_up_sync(0xffffffffu, xe, o), yt = __shfl_up_sync(0xffffffffu, xt, o);
    if (lane >= (uint32_t)o) {
      xe += ye;
      xt += yt;
    }
  }
  if (lane == 31) {
    warp_e[warp] = xe;
    warp_t[warp] = xt;
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t a = warp_e[lane], b = warp_t[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t ya = __shfl_up_sync(0xffffffffu, a, o), yb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= (uint32_t)o) {
        a += ya;
        b += yb;
      }
    }
    warp_e[lane] = a;
    warp_t[lane] = b;
  }
  __syncthreads();
  uint32_t pe = (warp ? warp_e[warp - 1] : 0) + xe - se;
  uint32_t pt = (warp ? warp_t[warp - 1] : 0) + xt - st;
  for (uint32_t r = lo; r < hi; ++r) {
    uint32_t c = coarse_count[r];
    coarse_offset[r] = pe;
    coarse_cursor[r] = pe;
    tile_offset[r] = pt;
    pe += c;
    pt += (c + kSortTile - 1) / kSortTile;
  }
  if (threadIdx.x == 1023) {
    coarse_offset[regions] = warp_e[31];
    tile_offset[regions] = warp_t[31];
    totals->tiles = warp_t[31];
    nonzero_slots[0] = warp_e[31];  // the other slots were zeroed by the host
  }
}
#endif

// The four non-template kernels are DEFINED in one translation unit of their own
// (msm_sort_kernels.cu, which sets TB200_DEFINE_SORT_KERNELS and is compiled without nvcc's
// --split-compile) and only declared everywhere else: as `static` kernels of this header every
// curve's translation unit carried its own copy, and the copies came out of the split compile
// with different code from the same source — 64 registers in one unit, 128 (or, once capped,
// 64 with the per-thread arrays demoted to local memory) in another.
//
// Every kernel of this file states its resident CTAs per SM in __launch_bounds__: left to its
// own heuristics ptxas gave the SAME source 64 registers in one translation unit and 128 in
// another (and differently from build to build), which halved the occupancy of the two scatter
// kernels and cost 0.8 ms of a 2^24-point sort (3.43 -> 4.24 ms) in "unlucky" builds.
//
// grid (tiles of kSortTile points, W).  The tile is counting-sorted by coarse bin in shared
// memory first, so that the entries of one bin leave as one contiguous run: a warp then
// stores a handful of 64-byte segments instead of 32 scattered 8-byte words.
__global__ void __launch_bounds__(kSortThreads, 4) coarse_scatter_kernel(
    const uint32_t* __restrict__ digits, MsmPlan plan, SortPlan sp,
    uint32_t* __restrict__ coarse_cursor, uint2* __restrict__ mid);
#ifdef TB200_DEFINE_SORT_KERNELS
__global__ void __launch_bounds__(kSortThreads, 4) coarse_scatter_kernel(
    const uint32_t* __restrict__ digits, MsmPlan plan, SortPlan sp,
    uint32_t* __restrict__ coarse_cursor, uint2* __restrict__ mid){
  __shared__ uint32_t hist[kMaxCoarsePerWindow];   // counts, then local offsets
  __shared__ uint32_t delta[kMaxCoarsePerWindow];  // global run start - local offset
  __shared__ uint2 staged[kSortTile];
  __shared__ uint32_t warp_sums[kSortThreads / 32];
  const uint32_t w = blockIdx.y;
  for (uint32_t k = threadIdx.x; k < sp.Cw; k += kSortThreads) hist[k] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
  uint32_t d[kSortPerThread], rank[kSortPerThread];
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t i = base + k * kSortThreads + threadIdx.x;
    d[k] = i < plan.n ? digits[(size_t)w * plan.n + i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t mag = d[k] & 0x7fffffffu;
    rank[k] = shared_inc(hist, (mag - 1) >> kFineBits, mag != 0);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < sp.Cw; k += kSortThreads)
    delta[k] = hist[k] ? atomicAdd(coarse_cursor + w * sp.Cw + k, hist[k]) : 0u;
  __syncthreads();
  block_exclusive_scan_bins(hist, sp.Cw, warp_sums);
  uint32_t total = 0;
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t mag = d[k] & 0x7fffffffu;
    if (mag) {
      uint32_t i = base + k * kSortThreads + threadIdx.x;
      uint32_t key = mag - 1;
      staged[hist[key >> kFineBits] + rank[k]] = make_uint2(i | (d[k] & 0x80000000u), key);
    }
    total += mag != 0;
  }
  // entries in the tile = sum over threads
  total = __reduce_add_sync(0xffffffffu, total);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = total;
  __syncthreads();
  uint32_t count = 0;
#pragma unroll
  for (int k = 0; k < kSortThreads / 32; ++k) count += warp_sums[k];
  for (uint32_t k = threadIdx.x; k < sp.Cw; k += kSortThreads) delta[k] -= hist[k];
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < count; j += kSortThreads) {
    uint2 e = staged[j];
    mid[delta[e.y >> kFineBits] + j] = e;
  }
}
#endif

// Region and entry span of tile t.
TB_DEV bool tile_span(uint32_t t, const uint32_t* __restrict__ tile_offset,
                      const uint32_t* __restrict__ coarse_offset, uint32_t regions,
                      uint32_t& region, uint32_t& start, uint32_t& end) {
  uint32_t lo = 0, hi = regions;  // last r with tile_offset[r] <= t
  while (hi - lo > 1) {
    uint32_t mid_r = (lo + hi) >> 1;
    if (tile_offset[mid_r] <= t) lo = mid_r; else hi = mid_r;
  }
  region = lo;
  start = coarse_offset[lo] + (t - tile_offset[lo]) * kSortTile;
  end = min(coarse_offset[lo + 1], start + kSortTile);
  return start < end;
}

__global__ void __launch_bounds__(kSortThreads, 8) fine_hist_kernel(
    const uint2* __restrict__ mid, SortPlan sp, const uint32_t* __restrict__ tile_offset,
    const uint32_t* __restrict__ coarse_offset, const SortTotals* __restrict__ totals,
    uint32_t* __restrict__ count);
#ifdef TB200_DEFINE_SORT_KERNELS
__global__ void __launch_bounds__(kSortThreads, 8) fine_hist_kernel(
    const uint2* __restrict__ mid, SortPlan sp, const uint32_t* __restrict__ tile_offset,
    const uint32_t* __restrict__ coarse_offset, const SortTotals* __restrict__ totals,
    uint32_t* __restrict__ count){
  __shared__ uint32_t hist[kFineBins];
  if (blockIdx.x >= totals->tiles) return;
  uint32_t region, start, end;
  tile_span(blockIdx.x, tile_offset, coarse_offset, sp.regions, region, start, end);
  for (uint32_t k = threadIdx.x; k < kFineBins; k += kSortThreads) hist[k] = 0;
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t p = start + k * kSortThreads + threadIdx.x;
    bool in = p < end;
    uint32_t key = in ? mid[p].y : 0u;
    shared_inc(hist, key & (kFineBins - 1u), in);
  }
  __syncthreads();
  uint32_t* dst = count + (size_t)region * kFineBins;
  for (uint32_t k = threadIdx.x; k < kFineBins; k += kSortThreads)
    if (hist[k]) atomicAdd(dst + k, hist[k]);
}
#endif

__global__ void __launch_bounds__(kSortThreads, 4) fine_scatter_kernel(
    const uint2* __restrict__ mid, SortPlan sp, const uint32_t* __restrict__ tile_offset,
    const uint32_t* __restrict__ coarse_offset, const SortTotals* __restrict__ totals,
    uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted);
#ifdef TB200_DEFINE_SORT_KERNELS
__global__ void __launch_bounds__(kSortThreads, 4) fine_scatter_kernel(
    const uint2* __restrict__ mid, SortPlan sp, const uint32_t* __restrict__ tile_offset,
    const uint32_t* __restrict__ coarse_offset, const SortTotals* __restrict__ totals,
    uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted){
  __shared__ uint32_t hist[kFineBins];   // counts, then local offsets
  __shared__ uint32_t delta[kFineBins];  // global run start - local offset
  __shared__ uint32_t staged[kSortTile];
  __shared__ uint16_t staged_bin[kSortTile];
  __shared__ uint32_t warp_sums[kSortThreads / 32];
  if (blockIdx.x >= totals->tiles) return;
  uint32_t region, start, end;
  tile_span(blockIdx.x, tile_offset, coarse_offset, sp.regions, region, start, end);
  for (uint32_t k = threadIdx.x; k < kFineBins; k += kSortThreads) hist[k] = 0;
  __syncthreads();
  uint2 e[kSortPerThread];
  uint32_t rank[kSortPerThread];
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k) {
    uint32_t p = start + k * kSortThreads + threadIdx.x;
    e[k] = p < end ? mid[p] : make_uint2(0u, 0xffffffffu);
  }
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k)
    rank[k] = shared_inc(hist, e[k].y & (kFineBins - 1u), e[k].y != 0xffffffffu);
  __syncthreads();
  uint32_t* cur = cursor + (size_t)region * kFineBins;
  for (uint32_t k = threadIdx.x; k < kFineBins; k += kSortThreads)
    delta[k] = hist[k] ? atomicAdd(cur + k, hist[k]) : 0u;
  __syncthreads();
  block_exclusive_scan_bins(hist, kFineBins, warp_sums);
#pragma unroll
  for (int k = 0; k < kSortPerThread; ++k) {
    if (e[k].y != 0xffffffffu) {
      uint32_t bin = e[k].y & (kFineBins - 1u);
      uint32_t pos = hist[bin] + rank[k];
      staged[pos] = e[k].x;
      staged_bin[pos] = (uint16_t)bin;
    }
  }
  for (uint32_t k = threadIdx.x; k < kFineBins; k += kSortThreads) delta[k] -= hist[k];
  __syncthreads();
  const uint32_t count = end - start;
  for (uint32_t j = threadIdx.x; j < count; j += kSortThreads)
    sorted[delta[staged_bin[j]] + j] = staged[j];
}
#endif

}  // namespace tb200
