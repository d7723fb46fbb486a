// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces, for the device side of the MSM path, what the reference reaches
// through tachyon/math/finite_fields/prime_field_gpu.h:327-429 (MulLimbs /
// MadNRedc / Clamp) and kernels/prime_field_ops_internal.cu.h:13-131.  Values
// are little-endian u32 limbs, Montgomery form with R = 2^(32*N) (identical
// bytes to the reference's u64-limb R = 2^(64*N/2)), always fully reduced to
// [0, p) on function exit so that zero/equality tests are plain limb compares.
//
// The multiplier is written so that every 32x32->64 product is one
// IMAD.WIDE.U32(.X) with the carry riding the predicate: products of one row
// are split by column parity into two accumulators whose 64-bit words never
// straddle (ptxas fuses each mad.lo.cc/madc.hi.cc pair below into a single
// IMAD.WIDE.U32.X — checked with cuobjdump -sass).  Per N-limb multiply:
// 2*N*N + N products (136 for BN254, 300 for BLS12-381), which is exactly the
// per-modmul figure SURVEY.md §8(d) uses for the IMAD roofline.
#pragma once
#include <stdint.h>

#include "field_constants.h"

namespace tb200 {

#define TB_DEV __device__ __forceinline__

// ---- carry-chain primitives (CC flag lives across consecutive statements) --
TB_DEV uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
TB_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
TB_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
TB_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
TB_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
               : "=r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}

// F is one of the *Params structs of field_constants.h.
template <class F>
struct Fp {
  static constexpr int N = F::kLimbs32;
  uint32_t l[N];
};

template <class F>
TB_DEV void fp_set_zero(Fp<F>& r) {
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) r.l[i] = 0;
}
template <class F>
TB_DEV void fp_set_one(Fp<F>& r) {
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) r.l[i] = F::one(i);
}
template <class F>
TB_DEV bool fp_is_zero(const Fp<F>& a) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) acc |= a.l[i];
  return acc == 0;
}
template <class F>
TB_DEV bool fp_eq(const Fp<F>& a, const Fp<F>& b) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) acc |= a.l[i] ^ b.l[i];
  return acc == 0;
}

// r = t - p if t >= p else t, for t < 2p  (the Clamp of big_int.h:279-291).
template <class F>
TB_DEV void fp_reduce_once(Fp<F>& r, const uint32_t (&t)[Fp<F>::N]) {
  constexpr int N = Fp<F>::N;
  uint32_t u[N];
  u[0] = sub_cc(t[0], F::mod(0));
#pragma unroll
  for (int i = 1; i < N; ++i) u[i] = subc_cc(t[i], F::mod(i));
  uint32_t borrow = subc(0u, 0u);  // 0xffffffff when t < p
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = borrow ? t[i] : u[i];
}

// r = a + b mod p   (prime_field_fallback.h:199-206)
template <class F>
TB_DEV void fp_add(Fp<F>& r, const Fp<F>& a, const Fp<F>& b) {
  constexpr int N = Fp<F>::N;
  uint32_t t[N];
  t[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; ++i) t[i] = addc_cc(a.l[i], b.l[i]);
  t[N - 1] = addc(a.l[N - 1], b.l[N - 1]);  // p < 2^(32N-1): no carry out
  fp_reduce_once<F>(r, t);
}
template <class F>
TB_DEV void fp_dbl(Fp<F>& r, const Fp<F>& a) {
  fp_add<F>(r, a, a);
}

// r = a - b mod p   (prime_field_fallback.h:234-243)
template <class F>
TB_DEV void fp_sub(Fp<F>& r, const Fp<F>& a, const Fp<F>& b) {
  constexpr int N = Fp<F>::N;
  uint32_t t[N];
  t[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N; ++i) t[i] = subc_cc(a.l[i], b.l[i]);
  uint32_t borrow = subc(0u, 0u);  // all-ones when a < b
  r.l[0] = add_cc(t[0], F::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(t[i], F::mod(i) & borrow);
  r.l[N - 1] = addc(t[N - 1], F::mod(N - 1) & borrow);
}

// r = -a mod p, with -0 = 0   (prime_field_fallback.h:253-260)
template <class F>
TB_DEV void fp_neg(Fp<F>& r, const Fp<F>& a) {
  constexpr int N = Fp<F>::N;
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) nz |= a.l[i];
  uint32_t mask = nz ? 0xffffffffu : 0u;
  r.l[0] = sub_cc(F::mod(0) & mask, a.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; ++i) r.l[i] = subc_cc(F::mod(i) & mask, a.l[i]);
  r.l[N - 1] = subc(F::mod(N - 1) & mask, a.l[N - 1]);
}
// r = neg ? -a : a
template <class F>
TB_DEV void fp_cneg(Fp<F>& r, const Fp<F>& a, bool neg) {
  Fp<F> n;
  fp_neg<F>(n, a);
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) r.l[i] = neg ? n.l[i] : a.l[i];
}

// ---------------------------------------------------------------------------
// Montgomery multiplication, row-interleaved, parity-split accumulators.
//
// The running value V (N+1 limbs) is held as  V = X + (Y << 32)  where X and Y
// are N-limb arrays read as N/2 64-bit words, so a product a_j*b_i always
// lands on a whole word of one of them (even j -> X, odd j -> Y).  After the
// Montgomery row (adding m*p with m = X[0]*inv) limb X[0] is zero and the
// implicit division by 2^32 swaps the roles: Y becomes the aligned array and
// X, moved down one word, the offset one.  That move is folded into the next
// row's first chain (the addend is read two limbs ahead), so it costs no
// instruction; only X[1], the half word left behind, needs one add.
// Carries out of the aligned chain go to Y's top limb; carries out of the
// offset chain cannot occur because V < 2^(32(N+1)) throughout
// (V < 2p + 2^33 p and p < 2^(32N-2) for both curves).
// ---------------------------------------------------------------------------
template <class F, int N>
TB_DEV void mont_reduce_row(uint32_t (&X)[N], uint32_t (&Y)[N]) {
  uint32_t m = X[0] * F::kInv32;
  Y[0] = mad_lo_cc(F::mod(1), m, Y[0]);
  Y[1] = madc_hi_cc(F::mod(1), m, Y[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    Y[j] = madc_lo_cc(F::mod(j + 1), m, Y[j]);
    Y[j + 1] = madc_hi_cc(F::mod(j + 1), m, Y[j + 1]);
  }
  X[0] = mad_lo_cc(F::mod(0), m, X[0]);
  X[1] = madc_hi_cc(F::mod(0), m, X[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    X[j] = madc_lo_cc(F::mod(j), m, X[j]);
    X[j + 1] = madc_hi_cc(F::mod(j), m, X[j + 1]);
  }
  Y[N - 1] = addc(Y[N - 1], 0u);
}

// One row i >= 1.  On entry Y is the previous row's aligned array (Y[0] == 0).
template <class F, int N>
TB_DEV void mont_mul_row(uint32_t (&X)[N], uint32_t (&Y)[N], const uint32_t (&a)[N],
                         uint32_t bi) {
  X[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    uint32_t c_lo = (j + 2 < N) ? Y[j + 2] : 0u;
    uint32_t c_hi = (j + 3 < N) ? Y[j + 3] : 0u;
    Y[j] = madc_lo_cc(a[j + 1], bi, c_lo);
    Y[j + 1] = madc_hi_cc(a[j + 1], bi, c_hi);
  }
  X[0] = mad_lo_cc(a[0], bi, X[0]);
  X[1] = madc_hi_cc(a[0], bi, X[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    X[j] = madc_lo_cc(a[j], bi, X[j]);
    X[j + 1] = madc_hi_cc(a[j], bi, X[j + 1]);
  }
  Y[N - 1] = addc(Y[N - 1], 0u);
  mont_reduce_row<F, N>(X, Y);
}

// r = a * b * R^-1 mod p, a, b < p  (value of prime_field_fallback.h:331-355)
template <class F>
TB_DEV void fp_mul(Fp<F>& r, const Fp<F>& a, const Fp<F>& b) {
  constexpr int N = Fp<F>::N;
  static_assert(N % 2 == 0, "limb count must be even");
  uint32_t E[N], O[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    mul_wide(E[j], E[j + 1], a.l[j], b.l[0]);
    mul_wide(O[j], O[j + 1], a.l[j + 1], b.l[0]);
  }
  mont_reduce_row<F, N>(E, O);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_mul_row<F, N>(O, E, a.l, b.l[i]);
    if (i + 1 < N) mont_mul_row<F, N>(E, O, a.l, b.l[i + 1]);
  }
  // last row had X = O (O[0] == 0), Y = E:  result = E + (O >> 32)
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  fp_reduce_once<F>(r, t);
}

// r = (a * b + a2 * b2) * R^-1 mod p with ONE Montgomery reduction: each row adds two
// product rows before its reduction row, saving N*N + N of the 2*(2*N*N + N) products of
// two separate multiplications.  Bounds: V < 4p + 3 * 2^32 p < 2^(32(N+1)) and the result
// is < p (2p/R + 1) < 2p, so one conditional subtraction still canonicalises it.
template <class F, int N>
TB_DEV void mont_mul2_row(uint32_t (&X)[N], uint32_t (&Y)[N], const uint32_t (&a)[N], uint32_t bi,
                          const uint32_t (&a2)[N], uint32_t b2i) {
  X[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    uint32_t c_lo = (j + 2 < N) ? Y[j + 2] : 0u;
    uint32_t c_hi = (j + 3 < N) ? Y[j + 3] : 0u;
    Y[j] = madc_lo_cc(a[j + 1], bi, c_lo);
    Y[j + 1] = madc_hi_cc(a[j + 1], bi, c_hi);
  }
  Y[0] = mad_lo_cc(a2[1], b2i, Y[0]);
  Y[1] = madc_hi_cc(a2[1], b2i, Y[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    Y[j] = madc_lo_cc(a2[j + 1], b2i, Y[j]);
    Y[j + 1] = madc_hi_cc(a2[j + 1], b2i, Y[j + 1]);
  }
  X[0] = mad_lo_cc(a[0], bi, X[0]);
  X[1] = madc_hi_cc(a[0], bi, X[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    X[j] = madc_lo_cc(a[j], bi, X[j]);
    X[j + 1] = madc_hi_cc(a[j], bi, X[j + 1]);
  }
  Y[N - 1] = addc(Y[N - 1], 0u);
  X[0] = mad_lo_cc(a2[0], b2i, X[0]);
  X[1] = madc_hi_cc(a2[0], b2i, X[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    X[j] = madc_lo_cc(a2[j], b2i, X[j]);
    X[j + 1] = madc_hi_cc(a2[j], b2i, X[j + 1]);
  }
  Y[N - 1] = addc(Y[N - 1], 0u);
  mont_reduce_row<F, N>(X, Y);
}

template <class F>
TB_DEV void fp_mul2(Fp<F>& r, const Fp<F>& a, const Fp<F>& b, const Fp<F>& a2, const Fp<F>& b2) {
  constexpr int N = Fp<F>::N;
  uint32_t E[N], O[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    mul_wide(E[j], E[j + 1], a.l[j], b.l[0]);
    mul_wide(O[j], O[j + 1], a.l[j + 1], b.l[0]);
  }
  // second product of row 0; carries out of the aligned chain go to O's top limb
  O[0] = mad_lo_cc(a2.l[1], b2.l[0], O[0]);
  O[1] = madc_hi_cc(a2.l[1], b2.l[0], O[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    O[j] = madc_lo_cc(a2.l[j + 1], b2.l[0], O[j]);
    O[j + 1] = madc_hi_cc(a2.l[j + 1], b2.l[0], O[j + 1]);
  }
  E[0] = mad_lo_cc(a2.l[0], b2.l[0], E[0]);
  E[1] = madc_hi_cc(a2.l[0], b2.l[0], E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    E[j] = madc_lo_cc(a2.l[j], b2.l[0], E[j]);
    E[j + 1] = madc_hi_cc(a2.l[j], b2.l[0], E[j + 1]);
  }
  O[N - 1] = addc(O[N - 1], 0u);
  mont_reduce_row<F, N>(E, O);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_mul2_row<F, N>(O, E, a.l, b.l[i], a2.l, b2.l[i]);
    if (i + 1 < N) mont_mul2_row<F, N>(E, O, a.l, b.l[i + 1], a2.l, b2.l[i + 1]);
  }
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  fp_reduce_once<F>(r, t);
}

// ---------------------------------------------------------------------------
// The same two multiplications with their rows in a LOOP (two rows per trip, so the
// accumulators keep their roles; the multiplier's limbs move down two places per trip, so every
// index stays a compile-time constant and everything stays in registers).  Same products, same
// carry chains, same result — about a third of the code.  The unrolled forms above are ~330
// (12 limbs) instructions per multiplication and a point addition inlines ten of them: the loop
// body of the 12-limb accumulation kernel is 144 KB of SASS, the instruction cache serves 79 %
// of its requests and `no_instruction` is its second largest stall (ncu, BLS12-381 2^22: FMA-
// heavy pipe 82 % against 91 % for the 8-limb kernel, whose 68 KB loop body still streams
// through the prefetcher at a 98 % hit rate).  Row 0 is an ordinary row over zeroed
// accumulators here: the same number of multiply instructions as the mul.wide start.
// ---------------------------------------------------------------------------
template <class F>
TB_DEV void fp_mul_rolled(Fp<F>& r, const Fp<F>& a, const Fp<F>& b) {
  constexpr int N = Fp<F>::N;
  static_assert(N % 2 == 0, "limb count must be even");
  uint32_t E[N], O[N], m[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    E[j] = 0u;
    O[j] = 0u;
    m[j] = b.l[j];
  }
#pragma unroll 1
  for (int trip = 0; trip < N / 2; ++trip) {
    mont_mul_row<F, N>(E, O, a.l, m[0]);
    mont_mul_row<F, N>(O, E, a.l, m[1]);
#pragma unroll
    for (int j = 0; j + 2 < N; ++j) m[j] = m[j + 2];
  }
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  fp_reduce_once<F>(r, t);
}

template <class F>
TB_DEV void fp_mul2_rolled(Fp<F>& r, const Fp<F>& a, const Fp<F>& b, const Fp<F>& a2,
                           const Fp<F>& b2) {
  constexpr int N = Fp<F>::N;
  uint32_t E[N], O[N], m[N], m2[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    E[j] = 0u;
    O[j] = 0u;
    m[j] = b.l[j];
    m2[j] = b2.l[j];
  }
#pragma unroll 1
  for (int trip = 0; trip < N / 2; ++trip) {
    mont_mul2_row<F, N>(E, O, a.l, m[0], a2.l, m2[0]);
    mont_mul2_row<F, N>(O, E, a.l, m[1], a2.l, m2[1]);
#pragma unroll
    for (int j = 0; j + 2 < N; ++j) {
      m[j] = m[j + 2];
      m2[j] = m2[j + 2];
    }
  }
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  fp_reduce_once<F>(r, t);
}

// Squaring: a^2 = sum_i a_i * (a_i 2^(32 i) + sum_(j > i) 2 a_j 2^(32 j)) 2^(32 i).  Row i
// multiplies a_i with the vector m_i = [0, .., 0, a_i, b_(i+1), .., b_(N-1)], b = 2a (fits N limbs:
// both moduli leave two spare bits; limb i+1 with its lowest bit cleared — that bit is the top
// bit of a_i, which belongs to the part of a that is NOT doubled in row i), so it has N - i
// products instead of N: N (N + 1) / 2 + N^2
// + N = 108 instead of 136 (BN254).  The rows keep the interleaved structure of fp_mul; where
// the multiplicand limb is zero the offset chain only moves its carry on (two adds on the
// idle ALU pipe instead of one IMAD.WIDE on the busy one) and the aligned chain starts later.
// Bounds: m_i < 2p keeps V < 3p + 2^33 p < 2^(32 (N + 1)); the total is a^2 + M p < R p + p^2,
// so the result is < 2p exactly as for fp_mul.
template <class F, int N, int kSkip>
TB_DEV void mont_sqr_row(uint32_t (&X)[N], uint32_t (&Y)[N], const uint32_t (&m)[N], uint32_t bi) {
  X[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    uint32_t c_lo = (j + 2 < N) ? Y[j + 2] : 0u;
    uint32_t c_hi = (j + 3 < N) ? Y[j + 3] : 0u;
    if (j + 1 < kSkip) {
      Y[j] = addc_cc(c_lo, 0u);
      Y[j + 1] = addc_cc(c_hi, 0u);
    } else {
      Y[j] = madc_lo_cc(m[j + 1], bi, c_lo);
      Y[j + 1] = madc_hi_cc(m[j + 1], bi, c_hi);
    }
  }
  constexpr int j0 = (kSkip + 1) & ~1;  // first even multiplicand index that is not zero
  if (j0 < N) {
    X[j0] = mad_lo_cc(m[j0], bi, X[j0]);
    X[j0 + 1] = madc_hi_cc(m[j0], bi, X[j0 + 1]);
#pragma unroll
    for (int j = j0 + 2; j < N; j += 2) {
      X[j] = madc_lo_cc(m[j], bi, X[j]);
      X[j + 1] = madc_hi_cc(m[j], bi, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0u);
  }
  mont_reduce_row<F, N>(X, Y);
}

template <class F, int N, int kRow>
struct SqrRows {
  // rows kRow, kRow + 1 (roles of the accumulators alternate), then the rest
  static TB_DEV void run(uint32_t (&E)[N], uint32_t (&O)[N], const uint32_t (&a)[N],
                         const uint32_t (&b)[N]) {
    uint32_t m[N];
#pragma unroll
    for (int j = 0; j < N; ++j)
      m[j] = j < kRow ? 0u : (j == kRow ? a[j] : (j == kRow + 1 ? (b[j] & ~1u) : b[j]));
    mont_sqr_row<F, N, kRow>(O, E, m, a[kRow]);
    if (kRow + 1 < N) {
      uint32_t m2[N];
#pragma unroll
      for (int j = 0; j < N; ++j)
        m2[j] = j < kRow + 1 ? 0u : (j == kRow + 1 ? a[j] : (j == kRow + 2 ? (b[j] & ~1u) : b[j]));
      mont_sqr_row<F, N, kRow + 1>(E, O, m2, a[kRow + 1]);
    }
    SqrRows<F, N, kRow + 2>::run(E, O, a, b);
  }
};
template <class F, int N>
struct SqrRows<F, N, N + 1> {
  static TB_DEV void run(uint32_t (&)[N], uint32_t (&)[N], const uint32_t (&)[N],
                         const uint32_t (&)[N]) {}
};

template <class F>
TB_DEV void fp_sqr(Fp<F>& r, const Fp<F>& a) {
  constexpr int N = Fp<F>::N;
  static_assert(N % 2 == 0, "limb count must be even");
  uint32_t b[N];  // 2a, no reduction: a < p < 2^(32N - 2)
  b[0] = a.l[0] << 1;
#pragma unroll
  for (int j = 1; j < N; ++j) b[j] = __funnelshift_l(a.l[j - 1], a.l[j], 1);
  uint32_t E[N], O[N];
  // row 0: a_0 * [a_0, b_1, .., b_(N-1)]
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    mul_wide(E[j], E[j + 1], j == 0 ? a.l[0] : b[j], a.l[0]);
    mul_wide(O[j], O[j + 1], j == 0 ? (b[1] & ~1u) : b[j + 1], a.l[0]);
  }
  mont_reduce_row<F, N>(E, O);
  SqrRows<F, N, 1>::run(E, O, a.l, b);
  // N rows in total, N even: the last row had X = O (O[0] == 0), Y = E
  uint32_t t[N];
  t[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 1; j < N - 1; ++j) t[j] = addc_cc(E[j], O[j + 1]);
  t[N - 1] = addc(E[N - 1], 0u);
  fp_reduce_once<F>(r, t);
}

// Montgomery -> canonical: a * 1 * R^-1  (value of big_int.h:1049-1076
// FromMontgomery).  Fully reduced for every modulus, including BLS12-381 Fr
// which has a single spare bit (the lazy [0,2p) form of
// prime_field_gpu.h:256 would not be valid there).
template <class F>
TB_DEV void fp_from_mont(Fp<F>& r, const Fp<F>& a) {
  Fp<F> one;
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) one.l[i] = (i == 0) ? 1u : 0u;
  fp_mul<F>(r, a, one);
}

template <class F>
TB_DEV void fp_to_mont(Fp<F>& r, const Fp<F>& a) {
  Fp<F> r2;
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) r2.l[i] = F::r2(i);
  fp_mul<F>(r, a, r2);
}

// a^(p-2): only used by test/bench utilities (point normalisation).
template <class F>
__device__ __noinline__ void fp_inv(Fp<F>& r, const Fp<F>& a) {
  constexpr int N = Fp<F>::N;
  uint32_t e[N];
  e[0] = sub_cc(F::mod(0), 2u);
#pragma unroll
  for (int i = 1; i < N; ++i) e[i] = subc_cc(F::mod(i), 0u);
  Fp<F> acc;
  fp_set_one<F>(acc);
  for (int i = F::kBits - 1; i >= 0; --i) {
    fp_sqr<F>(acc, acc);
    if ((e[i >> 5] >> (i & 31)) & 1) fp_mul<F>(acc, acc, a);
  }
  r = acc;
}

// 128-bit vector load/store of a field element (address must be 16-B aligned)
template <class F>
TB_DEV void fp_load(Fp<F>& r, const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < Fp<F>::N / 4; ++i) {
    uint4 v = __ldg(q + i);
    r.l[4 * i] = v.x;
    r.l[4 * i + 1] = v.y;
    r.l[4 * i + 2] = v.z;
    r.l[4 * i + 3] = v.w;
  }
}
template <class F>
TB_DEV void fp_load_rw(Fp<F>& r, const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < Fp<F>::N / 4; ++i) {
    uint4 v = q[i];
    r.l[4 * i] = v.x;
    r.l[4 * i + 1] = v.y;
    r.l[4 * i + 2] = v.z;
    r.l[4 * i + 3] = v.w;
  }
}
template <class F>
TB_DEV void fp_store(void* p, const Fp<F>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < Fp<F>::N / 4; ++i)
    q[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}

// ---------------------------------------------------------------------------
// Coordinate fields of the point code.  A "field kind" K names an element type
// K::El (kWords u32 words, stored contiguously) and the operations the XYZZ
// formulas need; the point and MSM kernels are written once against K and
// instantiated for the base field (G1) and its quadratic extension (G2).
// ---------------------------------------------------------------------------
// kRoll: 0 = unrolled multiplications, 1 = looped multiplications (fp_mul_rolled /
// fp_mul2_rolled) and the unrolled squaring, 2 = squarings through the looped multiplication
// too (28 % more products per squaring for the smallest code).
template <class F, int kRoll = 0>
struct FpField {
  using Params = F;
  using El = Fp<F>;
  template <int R>
  using WithRoll = FpField<F, R>;  // the same field, other code shape
  static constexpr int kWords = Fp<F>::N;
  static constexpr int kDegree = 1;
  static TB_DEV void set_zero(El& r) { fp_set_zero<F>(r); }
  static TB_DEV void set_one(El& r) { fp_set_one<F>(r); }
  static TB_DEV bool is_zero(const El& a) { return fp_is_zero<F>(a); }
  static TB_DEV bool eq(const El& a, const El& b) { return fp_eq<F>(a, b); }
  static TB_DEV void add(El& r, const El& a, const El& b) { fp_add<F>(r, a, b); }
  static TB_DEV void sub(El& r, const El& a, const El& b) { fp_sub<F>(r, a, b); }
  static TB_DEV void dbl(El& r, const El& a) { fp_dbl<F>(r, a); }
  static TB_DEV void neg(El& r, const El& a) { fp_neg<F>(r, a); }
  static TB_DEV void cneg(El& r, const El& a, bool n) { fp_cneg<F>(r, a, n); }
  static TB_DEV void mul(El& r, const El& a, const El& b) {
    if (kRoll) fp_mul_rolled<F>(r, a, b);
    else fp_mul<F>(r, a, b);
  }
  static TB_DEV void sqr(El& r, const El& a) {
    if (kRoll == 2) fp_mul_rolled<F>(r, a, a);
    else fp_sqr<F>(r, a);
  }
  // r = a * b + a2 * b2
  static TB_DEV void mul2(El& r, const El& a, const El& b, const El& a2, const El& b2) {
    if (kRoll) fp_mul2_rolled<F>(r, a, b, a2, b2);
    else fp_mul2<F>(r, a, b, a2, b2);
  }
  static TB_DEV void inv(El& r, const El& a) { fp_inv<F>(r, a); }
  static TB_DEV void select(El& r, bool take_a, const El& a, const El& b) {
#pragma unroll
    for (int i = 0; i < kWords; ++i) r.l[i] = take_a ? a.l[i] : b.l[i];
  }
  // element from the constant words G::aff32(base), G::aff32(base + 1), ...
  template <class G>
  static TB_DEV void set_words(El& r, int base) {
#pragma unroll
    for (int i = 0; i < kWords; ++i) r.l[i] = G::aff32(base + i);
  }
  static TB_DEV void load(El& r, const void* p) { fp_load<F>(r, p); }
  static TB_DEV void load_rw(El& r, const void* p) { fp_load_rw<F>(r, p); }
  static TB_DEV void store(void* p, const El& a) { fp_store<F>(p, a); }
  // r = a of lane `src` (index inside a group of `width` lanes; `mask` names the group's lanes)
  static TB_DEV void shfl(El& r, const El& a, uint32_t mask, int src, int width) {
#pragma unroll
    for (int i = 0; i < kWords; ++i) r.l[i] = __shfl_sync(mask, a.l[i], src, width);
  }
};

// Fq2 = Fq[u] / (u^2 + 1): both curves use the non-residue -1
// (bn/bn254/BUILD.bazel:62-71, bls12/bls12_381/BUILD.bazel fq2).  Formulas of
// tachyon/math/finite_fields/quadratic_extension_field.h: multiplication as two sums of
// two products (:326-338, SumOfProductsSerial — here one Montgomery reduction each via
// fp_mul2), complex squaring for q = -1 (:371-385), norm inverse (:407-427).
template <class F>
struct Fp2 {
  Fp<F> c0, c1;
};

template <class F, int kRoll = 0>
struct Fp2Field {
  using Params = F;
  using El = Fp2<F>;
  template <int R>
  using WithRoll = Fp2Field<F, R>;
  using Base = FpField<F, kRoll>;  // component arithmetic in the chosen code shape
  static constexpr int kWords = 2 * Fp<F>::N;
  static constexpr int kDegree = 2;
  static TB_DEV void set_zero(El& r) {
    fp_set_zero<F>(r.c0);
    fp_set_zero<F>(r.c1);
  }
  static TB_DEV void set_one(El& r) {
    fp_set_one<F>(r.c0);
    fp_set_zero<F>(r.c1);
  }
  static TB_DEV bool is_zero(const El& a) { return fp_is_zero<F>(a.c0) && fp_is_zero<F>(a.c1); }
  static TB_DEV bool eq(const El& a, const El& b) {
    return fp_eq<F>(a.c0, b.c0) && fp_eq<F>(a.c1, b.c1);
  }
  static TB_DEV void add(El& r, const El& a, const El& b) {
    fp_add<F>(r.c0, a.c0, b.c0);
    fp_add<F>(r.c1, a.c1, b.c1);
  }
  static TB_DEV void sub(El& r, const El& a, const El& b) {
    fp_sub<F>(r.c0, a.c0, b.c0);
    fp_sub<F>(r.c1, a.c1, b.c1);
  }
  static TB_DEV void dbl(El& r, const El& a) {
    fp_dbl<F>(r.c0, a.c0);
    fp_dbl<F>(r.c1, a.c1);
  }
  static TB_DEV void neg(El& r, const El& a) {
    fp_neg<F>(r.c0, a.c0);
    fp_neg<F>(r.c1, a.c1);
  }
  static TB_DEV void cneg(El& r, const El& a, bool n) {
    fp_cneg<F>(r.c0, a.c0, n);
    fp_cneg<F>(r.c1, a.c1, n);
  }
  static TB_DEV void mul(El& r, const El& a, const El& b) {
    Fp<F> nb1, t0;
    fp_neg<F>(nb1, b.c1);
    Base::mul2(t0, a.c0, b.c0, a.c1, nb1);    // c0 = a0 b0 - a1 b1
    Base::mul2(r.c1, a.c0, b.c1, a.c1, b.c0);  // c1 = a0 b1 + a1 b0
    r.c0 = t0;
  }
  static TB_DEV void sqr(El& r, const El& a) {
    Fp<F> s, d, m;
    fp_add<F>(s, a.c0, a.c1);
    fp_sub<F>(d, a.c0, a.c1);
    Base::mul(m, a.c0, a.c1);
    Base::mul(r.c0, s, d);  // c0^2 - c1^2
    fp_dbl<F>(r.c1, m);     // 2 c0 c1
  }
  // r = a * b + a2 * b2: four products per component, still one reduction each
  static __device__ __noinline__ void mul2(El& r, const El& a, const El& b, const El& a2,
                                           const El& b2) {
    El t, u;
    mul(t, a, b);
    mul(u, a2, b2);
    add(r, t, u);
  }
  static __device__ __noinline__ void inv(El& r, const El& a) {
    Fp<F> n, t;
    fp_sqr<F>(n, a.c0);
    fp_sqr<F>(t, a.c1);
    fp_add<F>(n, n, t);  // norm = c0^2 + c1^2
    fp_inv<F>(t, n);
    fp_mul<F>(r.c0, a.c0, t);
    fp_mul<F>(n, a.c1, t);
    fp_neg<F>(r.c1, n);
  }
  static TB_DEV void select(El& r, bool take_a, const El& a, const El& b) {
#pragma unroll
    for (int i = 0; i < Fp<F>::N; ++i) {
      r.c0.l[i] = take_a ? a.c0.l[i] : b.c0.l[i];
      r.c1.l[i] = take_a ? a.c1.l[i] : b.c1.l[i];
    }
  }
  template <class G>
  static TB_DEV void set_words(El& r, int base) {
#pragma unroll
    for (int i = 0; i < Fp<F>::N; ++i) {
      r.c0.l[i] = G::aff32(base + i);
      r.c1.l[i] = G::aff32(base + Fp<F>::N + i);
    }
  }
  static TB_DEV void load(El& r, const void* p) {
    fp_load<F>(r.c0, p);
    fp_load<F>(r.c1, static_cast<const char*>(p) + 4 * Fp<F>::N);
  }
  static TB_DEV void load_rw(El& r, const void* p) {
    fp_load_rw<F>(r.c0, p);
    fp_load_rw<F>(r.c1, static_cast<const char*>(p) + 4 * Fp<F>::N);
  }
  static TB_DEV void store(void* p, const El& a) {
    fp_store<F>(p, a.c0);
    fp_store<F>(static_cast<char*>(p) + 4 * Fp<F>::N, a.c1);
  }
  static TB_DEV void shfl(El& r, const El& a, uint32_t mask, int src, int width) {
#pragma unroll
    for (int i = 0; i < Fp<F>::N; ++i) {
      r.c0.l[i] = __shfl_sync(mask, a.c0.l[i], src, width);
      r.c1.l[i] = __shfl_sync(mask, a.c1.l[i], src, width);
    }
  }
};

// ---------------------------------------------------------------------------
// Fq2 with its two components on NEIGHBOURING LANES (lane 2k: c0, lane 2k + 1: c1).
//
// The one-thread Fq2 kernels are bound by registers, not by the multiply pipe: a G2 XYZZ
// accumulator plus the point in flight is 96 (BN254) / 144 (BLS12-381) words before any
// temporary, so only 8 warps fit an SM and the dependent carry chains of one point addition
// leave the FMA-heavy pipe idle a third of the time (ncu: 66 %).  Split over a lane pair every
// lane holds HALF of every element — half the registers, twice the warps — and the
// multiplication splits without redundancy: with a = (a0, a1), b = (b0, b1) and u^2 = -1
//   lane 0:  c0 = a0 b0 + a1 (-b1)        lane 1:  c1 = a1 b0 + a0 b1
// i.e. both lanes run ONE fp_mul2 (two products, one Montgomery reduction: exactly the two
// sums of products of quadratic_extension_field.h:326-338) on operands exchanged by shuffle.
// The complex squaring (:371-385) is one fp_mul per lane: (a0 + a1)(a0 - a1) | (2 a1) a0.
// Addition, subtraction, negation are component-wise and need no exchange.
//
// All shuffles use the full-warp mask: the callers keep their warps converged (predicated
// selects instead of branches), which spares the WARPSYNC / collective bracket ptxas puts
// around every partial-mask shuffle.  `role` is the lane's component (threadIdx.x & 1).
// Results are the canonical values of Fp2Field's operations, bit for bit.
// ---------------------------------------------------------------------------
template <class F, int kRoll = 0>
struct Fp2Lanes {
  using E = Fp<F>;
  using Base = FpField<F, kRoll>;
  static constexpr int N = Fp<F>::N;
  static constexpr uint32_t kFull = 0xffffffffu;

  // the partner lane's component
  static TB_DEV void other(E& r, const E& a) {
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = __shfl_xor_sync(kFull, a.l[i], 1);
  }
  // true when the predicate holds on both lanes of the pair
  static TB_DEV bool both(bool v) {
    const int partner = __shfl_xor_sync(kFull, (int)v, 1);  // unconditional: every lane takes part
    return v && partner != 0;
  }
  static TB_DEV bool is_zero(const E& a) { return both(fp_is_zero<F>(a)); }
  static TB_DEV void set_one(E& r, uint32_t role) {
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = role ? 0u : F::one(i);
  }
  static TB_DEV void select(E& r, bool take_a, const E& a, const E& b) {
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = take_a ? a.l[i] : b.l[i];
  }

  // A multiplier prepared once for several products: (u, v) = (b0, -b1) on lane 0 and
  // (b0, b1) on lane 1 — the factors of this lane's own and of its partner's component of a.
  struct Multiplier {
    E u, v;
  };
  static TB_DEV void prepare(Multiplier& m, const E& b, uint32_t role) {
    E nb, send, recv;
    fp_neg<F>(nb, b);
    select(send, role != 0, nb, b);  // lane 1 hands over -b1, lane 0 hands over b0
    other(recv, send);
    select(m.u, role != 0, recv, b);  // lane 0: b0        lane 1: b0
    select(m.v, role != 0, b, recv);  // lane 0: -b1       lane 1: b1
  }
  // r = this lane's component of a * b
  static TB_DEV void mul(E& r, const E& a, const Multiplier& m) {
    E ao;
    other(ao, a);
    // lane 0: a0 b0 + a1 (-b1)      lane 1: a1 b0 + a0 b1
    Base::mul2(r, a, m.u, ao, m.v);
  }
  static TB_DEV void mul(E& r, const E& a, const E& b, uint32_t role) {
    Multiplier m;
    prepare(m, b, role);
    mul(r, a, m);
  }
  static TB_DEV void sqr(E& r, const E& a, uint32_t role) {
    E ao, x, y, d;
    other(ao, a);
    select(x, role != 0, a, ao);
    fp_add<F>(x, a, x);  // lane 0: a0 + a1      lane 1: 2 a1
    fp_sub<F>(d, a, ao);
    select(y, role != 0, ao, d);  // lane 0: a0 - a1      lane 1: a0
    Base::mul(r, x, y);
  }
};

}  // namespace tb200
