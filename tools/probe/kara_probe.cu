// Probe: Karatsuba (3 x 4x4 limbs) product + separate Montgomery reduction against the
// row-interleaved fp_mul of fp.cuh, BN254 Fq.  Development aid.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../tachyon_b200/csrc \
//        -o kara_probe kara_probe.cu && ./kara_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fp.cuh"

using namespace tb200;
using F = Bn254FqParams;

// r[0..7] = u[0..3] * v[0..3], parity-split accumulators so every product is one IMAD.WIDE.
TB_DEV void mul4x4(uint32_t (&r)[8], const uint32_t* u, const uint32_t* v) {
  uint32_t EV[8], OD[8];  // EV[k] = column k; OD[k] = column k + 1
#pragma unroll
  for (int k = 0; k < 8; ++k) EV[k] = OD[k] = 0;
  mul_wide(EV[0], EV[1], u[0], v[0]);
  mul_wide(EV[2], EV[3], u[2], v[0]);
  mul_wide(OD[0], OD[1], u[1], v[0]);
  mul_wide(OD[2], OD[3], u[3], v[0]);
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    {  // even columns: j = i mod 2, +2
      const int j0 = i & 1, c = i + j0;
      EV[c] = mad_lo_cc(u[j0], v[i], EV[c]);
      EV[c + 1] = madc_hi_cc(u[j0], v[i], EV[c + 1]);
      EV[c + 2] = madc_lo_cc(u[j0 + 2], v[i], EV[c + 2]);
      EV[c + 3] = madc_hi_cc(u[j0 + 2], v[i], EV[c + 3]);
      if (c + 4 < 8) EV[c + 4] = addc(EV[c + 4], 0u);  // that word is still empty: no cascade
    }
    {  // odd columns: j = (i + 1) mod 2, +2
      const int j1 = (i + 1) & 1, c = i + j1;  // odd
      OD[c - 1] = mad_lo_cc(u[j1], v[i], OD[c - 1]);
      OD[c] = madc_hi_cc(u[j1], v[i], OD[c]);
      OD[c + 1] = madc_lo_cc(u[j1 + 2], v[i], OD[c + 1]);
      OD[c + 2] = madc_hi_cc(u[j1 + 2], v[i], OD[c + 2]);
      if (c + 3 < 8) OD[c + 3] = addc(OD[c + 3], 0u);
    }
  }
  r[0] = EV[0];
  r[1] = add_cc(EV[1], OD[0]);
#pragma unroll
  for (int k = 2; k < 7; ++k) r[k] = addc_cc(EV[k], OD[k - 1]);
  r[7] = addc(EV[7], OD[6]);
}

// |x - y| for 4 limbs; returns all-ones if x < y.
TB_DEV uint32_t absdiff4(uint32_t (&d)[4], const uint32_t* x, const uint32_t* y) {
  d[0] = sub_cc(x[0], y[0]);
  d[1] = subc_cc(x[1], y[1]);
  d[2] = subc_cc(x[2], y[2]);
  d[3] = subc_cc(x[3], y[3]);
  uint32_t neg = subc(0u, 0u);  // 0xffffffff when x < y
  // two's complement negate when neg: (d ^ neg) - neg
  d[0] = sub_cc(d[0] ^ neg, neg);
  d[1] = subc_cc(d[1] ^ neg, neg);
  d[2] = subc_cc(d[2] ^ neg, neg);
  d[3] = subc(d[3] ^ neg, neg);
  return neg;
}

// On entry Y is the previous row's aligned array (Y[0] == 0); reduction row only.
template <class FF, int N>
TB_DEV void mont_reduce_only_row(uint32_t (&X)[N], uint32_t (&Y)[N]) {
  X[0] = add_cc(X[0], Y[1]);
  uint32_t m = X[0] * FF::kInv32;
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    uint32_t c_lo = (j + 2 < N) ? Y[j + 2] : 0u;
    uint32_t c_hi = (j + 3 < N) ? Y[j + 3] : 0u;
    Y[j] = madc_lo_cc(FF::mod(j + 1), m, c_lo);
    Y[j + 1] = madc_hi_cc(FF::mod(j + 1), m, c_hi);
  }
  X[0] = mad_lo_cc(FF::mod(0), m, X[0]);
  X[1] = madc_hi_cc(FF::mod(0), m, X[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    X[j] = madc_lo_cc(FF::mod(j), m, X[j]);
    X[j + 1] = madc_hi_cc(FF::mod(j), m, X[j + 1]);
  }
  Y[N - 1] = addc(Y[N - 1], 0u);
}

TB_DEV void fp_mul_kara(Fp<F>& r, const Fp<F>& a, const Fp<F>& b) {
  constexpr int N = 8;
  uint32_t z0[8], z2[8], zm[8], da[4], db[4];
  mul4x4(z0, a.l, b.l);
  mul4x4(z2, a.l + 4, b.l + 4);
  uint32_t na = absdiff4(da, a.l, a.l + 4);   // a0 - a1
  uint32_t nb = absdiff4(db, b.l + 4, b.l);   // b1 - b0
  mul4x4(zm, da, db);
  uint32_t neg = na ^ nb;                      // (a0-a1)(b1-b0) < 0
  // s = z0 + z2 + sign * zm   (9 limbs: s[0..7], top)
  uint32_t s[8];
  s[0] = add_cc(z0[0], z2[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k) s[k] = addc_cc(z0[k], z2[k]);
  uint32_t top = addc(0u, 0u);
  // + (zm ^ neg) + (neg & 1): carry-in via an add.cc that produces it
  add_cc(neg, neg & 1u);  // carry = 1 exactly when neg (0xffffffff + 1)
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = addc_cc(s[k], zm[k] ^ neg);
  top = addc(top, neg);   // + carry - (neg ? 1 : 0)   (neg = -1 as a limb)
  // T = z0 + s << 128 + z2 << 256
  uint32_t E[N], O[N], hi[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) E[k] = z0[k];
  E[4] = add_cc(z0[4], s[0]);
  E[5] = addc_cc(z0[5], s[1]);
  E[6] = addc_cc(z0[6], s[2]);
  E[7] = addc_cc(z0[7], s[3]);
  hi[0] = addc_cc(z2[0], s[4]);
  hi[1] = addc_cc(z2[1], s[5]);
  hi[2] = addc_cc(z2[2], s[6]);
  hi[3] = addc_cc(z2[3], s[7]);
  hi[4] = addc_cc(z2[4], top);
  hi[5] = addc_cc(z2[5], 0u);
  hi[6] = addc_cc(z2[6], 0u);
  hi[7] = addc(z2[7], 0u);
  // Montgomery reduction of the low half, eight rows, then + high half
#pragma unroll
  for (int k = 0; k < N; ++k) O[k] = 0;
  mont_reduce_row<F, N>(E, O);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_reduce_only_row<F, N>(O, E);
    if (i + 1 < N) mont_reduce_only_row<F, N>(E, O);
  }
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  t[0] = add_cc(t[0], hi[0]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(t[j], hi[j]);
  t[N - 1] = addc(t[N - 1], hi[N - 1]);
  fp_reduce_once<F>(r, t);
}

template <int V>
__global__ void __launch_bounds__(128, 4) chain(uint32_t iters, const uint32_t* in, uint32_t* out) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  Fp<F> x, y;
  fp_load<F>(x, in + (size_t)(g % 1024) * 16);
  fp_load<F>(y, in + (size_t)(g % 1024) * 16 + 8);
  for (uint32_t it = 0; it < iters; ++it) {
    Fp<F> t;
    if (V == 0) fp_mul<F>(t, x, y); else fp_mul_kara(t, x, y);
    x = y;
    y = t;
  }
  fp_store<F>(out + (size_t)g * 8, y);
}

int main() {
  const int threads = 148 * 4 * 128 * 4, iters = 2000;
  uint32_t *in, *o0, *o1;
  cudaMalloc(&in, 1024 * 16 * 4);
  cudaMalloc(&o0, (size_t)threads * 8 * 4);
  cudaMalloc(&o1, (size_t)threads * 8 * 4);
  uint32_t* h = (uint32_t*)malloc(1024 * 16 * 4);
  uint64_t st = 88172645463325252ull;
  for (int i = 0; i < 1024 * 16; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    h[i] = (uint32_t)st;
    if ((i & 7) == 7) h[i] &= 0x0fffffffu;  // < p
  }
  cudaMemcpy(in, h, 1024 * 16 * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms[2];
  for (int v = 0; v < 2; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) chain<0><<<threads / 128, 128>>>(iters, in, o0);
      else chain<1><<<threads / 128, 128>>>(iters, in, o1);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
    }
    printf("variant %d: %.3f ms, %.1f G modmul/s  (%s)\n", v, ms[v], (double)threads * iters / ms[v] / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  uint32_t* a = (uint32_t*)malloc((size_t)threads * 32);
  uint32_t* b = (uint32_t*)malloc((size_t)threads * 32);
  cudaMemcpy(a, o0, (size_t)threads * 32, cudaMemcpyDeviceToHost);
  cudaMemcpy(b, o1, (size_t)threads * 32, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t i = 0; i < (size_t)threads * 8; ++i) bad += a[i] != b[i];
  printf("mismatching limbs: %zu of %zu; speed-up %.3f\n", bad, (size_t)threads * 8, ms[0] / ms[1]);
  return bad != 0;
}
