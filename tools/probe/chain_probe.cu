// Latency of a chain of point doublings / additions on ONE lane versus the four-lane cooperative
// forms (xyzz.cuh Coop4) — the serial part of the window combination.  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tachyon_b200/csrc \
//        -o tools/probe/chain_probe tools/probe/chain_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>

#include "msm_kernels.cuh"

using namespace tb200;

template <class C>
__device__ void make_points(XYZZ<typename C::Field>& p, XYZZ<typename C::Field>& q) {
  using K = typename C::Field;
  Affine<K> gen;
  K::template set_words<typename C::Gen>(gen.x, 0);
  K::template set_words<typename C::Gen>(gen.y, K::kWords);
  xyzz_set_zero<K>(p);
  xyzz_madd<K>(p, gen, false);
  xyzz_dbl<K>(p);
  xyzz_madd<K>(p, gen, false);  // 3 G
  q = p;
  xyzz_dbl<K>(q);
  xyzz_dbl<K>(q);
  xyzz_madd<K>(q, gen, false);  // 13 G
}

// mode 0: doublings, one lane, out-of-line call; 1: doublings, 4 lanes; 2: additions acc += q
// (q doubled now and then so operands vary), one lane; 3: additions, 4 lanes
template <class C>
__global__ void chain_kernel(int mode, int iters, uint32_t* out) {
  using K = typename C::Field;
  XYZZ<K> p, q;
  make_points<C>(p, q);
  const uint32_t lane = threadIdx.x & 3, mask = 0xfu << (threadIdx.x & 28);
  if (mode == 0) {
    if (threadIdx.x == 0)
      for (int i = 0; i < iters; ++i) xyzz_dbl_nz<K>(p);
  } else if (mode == 1) {
    for (int i = 0; i < iters; ++i) Coop4<K>::dbl_nz(p, lane, mask);
  } else if (mode == 2) {
    if (threadIdx.x == 0)
      for (int i = 0; i < iters; ++i) {
        xyzz_add<K>(p, q);
        if ((i & 7) == 7) xyzz_add<K>(q, p);
      }
  } else {
    for (int i = 0; i < iters; ++i) {
      Coop4<K>::add(p, q, lane, mask);
      if ((i & 7) == 7) Coop4<K>::add(q, p, lane, mask);
    }
  }
  if (threadIdx.x == 0) xyzz_store<K>(out, p);
  if (threadIdx.x == 3 && (mode & 1)) xyzz_store<K>(out + 4 * K::kWords, p);
}

template <class C>
void run(const char* name) {
  using K = typename C::Field;
  const int words = 4 * K::kWords;
  uint32_t* d;
  cudaMalloc(&d, 2 * words * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<uint32_t> ref(words), got(2 * words);
  for (int mode = 0; mode < 4; ++mode) {
    const int iters = mode < 2 ? 256 : 64;
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      chain_kernel<C><<<1, (mode & 1) ? 4 : 32>>>(mode, iters, d);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaMemcpy(got.data(), d, 2 * words * 4, cudaMemcpyDeviceToHost);
    bool same = true;
    if (mode & 1) {
      same = memcmp(got.data(), ref.data(), words * 4) == 0 &&
             memcmp(got.data() + words, ref.data(), words * 4) == 0;
    } else {
      memcpy(ref.data(), got.data(), words * 4);
    }
    printf("%-14s %s x%d, %s: %8.1f us  (%.2f us each)%s%s\n", name, mode < 2 ? "dbl" : "add", iters,
           (mode & 1) ? "4 lanes" : "1 lane ", best * 1e3, best * 1e3 / (mode < 2 ? iters : iters + iters / 8),
           (mode & 1) ? (same ? "  == 1-lane result" : "  MISMATCH") : "", err ? cudaGetErrorString(err) : "");
  }
  cudaFree(d);
}

int main() {
  run<Bn254Curve>("bn254 g1");
  run<Bls381Curve>("bls12_381 g1");
  run<Bn254G2Curve>("bn254 g2");
  run<Bls381G2Curve>("bls12_381 g2");
  return 0;
}
