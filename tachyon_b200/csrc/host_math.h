// Host-side field / point arithmetic for the MSM epilogue (window Horner,
// combination of per-GPU partials, XYZZ -> Jacobian), on u64 limbs.
// Follows tachyon/math/elliptic_curves/msm/algorithms/pippenger/pippenger_base.h:59-77
// (AccumulateWindowSums) and short_weierstrass/point_xyzz.h:228-237 (ToJacobian);
// independent of oracle/ (which is test infrastructure and never linked here).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "field_constants.h"

namespace tb200 {

template <class F>
struct HostFp {
  static constexpr int N = F::kLimbs64;
  uint64_t v[N];

  static HostFp Zero() {
    HostFp r;
    for (int i = 0; i < N; ++i) r.v[i] = 0;
    return r;
  }
  static HostFp One() {
    HostFp r;
    for (int i = 0; i < N; ++i) r.v[i] = F::kOne64[i];
    return r;
  }
  bool IsZero() const {
    uint64_t acc = 0;
    for (int i = 0; i < N; ++i) acc |= v[i];
    return acc == 0;
  }
  bool IsOne() const {
    uint64_t acc = 0;
    for (int i = 0; i < N; ++i) acc |= v[i] ^ F::kOne64[i];
    return acc == 0;
  }
  static bool Geq(const uint64_t* a, const uint64_t* b) {
    for (int i = N; i-- > 0;) {
      if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
  }
  static void SubMod(uint64_t* a) {
    unsigned __int128 borrow = 0;
    for (int i = 0; i < N; ++i) {
      unsigned __int128 t = (unsigned __int128)a[i] - F::kMod64[i] - (uint64_t)borrow;
      a[i] = (uint64_t)t;
      borrow = (t >> 64) & 1;
    }
  }
  HostFp Add(const HostFp& o) const {
    HostFp r;
    unsigned __int128 carry = 0;
    for (int i = 0; i < N; ++i) {
      unsigned __int128 t = (unsigned __int128)v[i] + o.v[i] + (uint64_t)carry;
      r.v[i] = (uint64_t)t;
      carry = t >> 64;
    }
    if (carry || Geq(r.v, F::kMod64)) SubMod(r.v);
    return r;
  }
  HostFp Dbl() const { return Add(*this); }
  HostFp Sub(const HostFp& o) const {
    HostFp r;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < N; ++i) {
      unsigned __int128 t = (unsigned __int128)v[i] - o.v[i] - (uint64_t)borrow;
      r.v[i] = (uint64_t)t;
      borrow = (t >> 64) & 1;
    }
    if (borrow) {
      unsigned __int128 carry = 0;
      for (int i = 0; i < N; ++i) {
        unsigned __int128 t = (unsigned __int128)r.v[i] + F::kMod64[i] + (uint64_t)carry;
        r.v[i] = (uint64_t)t;
        carry = t >> 64;
      }
    }
    return r;
  }
  // CIOS with the product and reduction rows interleaved; both moduli leave the top bit of
  // the top limb clear, so the two carries of a row fit one word (the "no-carry" variant of
  // prime_field_fallback.h:331-355).
  HostFp Mul(const HostFp& o) const {
    typedef unsigned __int128 u128;
    uint64_t t[N];
    for (int i = 0; i < N; ++i) t[i] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; ++i) {
      const uint64_t bi = o.v[i];
      u128 x = (u128)v[0] * bi + t[0];
      uint64_t lo = (uint64_t)x;
      uint64_t c1 = (uint64_t)(x >> 64);
      const uint64_t m = lo * F::kInv64;
      u128 y = (u128)m * F::kMod64[0] + lo;
      uint64_t c2 = (uint64_t)(y >> 64);
#pragma GCC unroll 8
      for (int j = 1; j < N; ++j) {
        x = (u128)v[j] * bi + t[j] + c1;
        c1 = (uint64_t)(x >> 64);
        y = (u128)m * F::kMod64[j] + (uint64_t)x + c2;
        c2 = (uint64_t)(y >> 64);
        t[j - 1] = (uint64_t)y;
      }
      t[N - 1] = c1 + c2;
    }
    HostFp r;
    for (int i = 0; i < N; ++i) r.v[i] = t[i];
    if (Geq(r.v, F::kMod64)) SubMod(r.v);
    return r;
  }
  HostFp Sqr() const { return Mul(*this); }
  // a^(p-2); zero maps to zero
  HostFp Inv() const {
    uint64_t e[N];
    uint64_t borrow = 2;
    for (int i = 0; i < N; ++i) {
      e[i] = F::kMod64[i] - borrow;
      borrow = F::kMod64[i] < borrow ? 1 : 0;
    }
    HostFp acc = One();
    for (int i = 64 * N - 1; i >= 0; --i) {
      acc = acc.Sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.Mul(*this);
    }
    return acc;
  }
};

// Fq2 = Fq[u] / (u^2 + 1) on the host (quadratic_extension_field.h:315-427 for q = -1).
template <class F>
struct HostFp2 {
  HostFp<F> c0, c1;

  static HostFp2 Zero() { return HostFp2{HostFp<F>::Zero(), HostFp<F>::Zero()}; }
  static HostFp2 One() { return HostFp2{HostFp<F>::One(), HostFp<F>::Zero()}; }
  bool IsZero() const { return c0.IsZero() && c1.IsZero(); }
  bool IsOne() const { return c0.IsOne() && c1.IsZero(); }
  HostFp2 Add(const HostFp2& o) const { return HostFp2{c0.Add(o.c0), c1.Add(o.c1)}; }
  HostFp2 Sub(const HostFp2& o) const { return HostFp2{c0.Sub(o.c0), c1.Sub(o.c1)}; }
  HostFp2 Dbl() const { return HostFp2{c0.Dbl(), c1.Dbl()}; }
  HostFp2 Mul(const HostFp2& o) const {
    return HostFp2{c0.Mul(o.c0).Sub(c1.Mul(o.c1)), c0.Mul(o.c1).Add(c1.Mul(o.c0))};
  }
  HostFp2 Sqr() const { return HostFp2{c0.Add(c1).Mul(c0.Sub(c1)), c0.Mul(c1).Dbl()}; }
  HostFp2 Inv() const {
    HostFp<F> t = c0.Sqr().Add(c1.Sqr()).Inv();
    return HostFp2{c0.Mul(t), HostFp<F>::Zero().Sub(c1.Mul(t))};
  }
};

// Points over an element type E (HostFp<F> for G1, HostFp2<F> for G2).
template <class E>
struct HostPointXYZZ {
  E x, y, zz, zzz;

  static HostPointXYZZ Zero() { return HostPointXYZZ{E::One(), E::One(), E::Zero(), E::Zero()}; }
  bool IsZero() const { return zz.IsZero(); }

  HostPointXYZZ Dbl() const {
    if (IsZero()) return *this;
    E u = y.Dbl(), v = u.Sqr(), w = u.Mul(v), s = x.Mul(v);
    E m = x.Sqr();
    m = m.Add(m.Dbl());
    HostPointXYZZ r;
    r.x = m.Sqr().Sub(s.Dbl());
    r.y = m.Mul(s.Sub(r.x)).Sub(w.Mul(y));
    r.zz = v.Mul(zz);
    r.zzz = w.Mul(zzz);
    return r;
  }
  HostPointXYZZ Add(const HostPointXYZZ& b) const {
    if (IsZero()) return b;
    if (b.IsZero()) return *this;
    E u1 = x.Mul(b.zz), s1 = y.Mul(b.zzz);
    E p = b.x.Mul(zz).Sub(u1), r = b.y.Mul(zzz).Sub(s1);
    if (p.IsZero() && r.IsZero()) return Dbl();
    E pp = p.Sqr(), ppp = p.Mul(pp), q = u1.Mul(pp);
    HostPointXYZZ c;
    c.x = r.Sqr().Sub(ppp).Sub(q.Dbl());
    c.y = r.Mul(q.Sub(c.x)).Sub(s1.Mul(ppp));
    c.zz = zz.Mul(b.zz).Mul(pp);
    c.zzz = zzz.Mul(b.zzz).Mul(ppp);
    return c;
  }
};

template <class E>
struct HostPointJacobian {
  E x, y, z;
};

template <class E>
struct HostPointAffine {
  E x, y;
};

template <class F>
using HostXYZZ = HostPointXYZZ<HostFp<F>>;
template <class F>
using HostJacobian = HostPointJacobian<HostFp<F>>;
template <class F>
using HostAffine = HostPointAffine<HostFp<F>>;

// point_xyzz.h:228-237
template <class E>
HostPointJacobian<E> ToJacobian(const HostPointXYZZ<E>& p) {
  if (p.IsZero()) return HostPointJacobian<E>{E::One(), E::One(), E::Zero()};
  if (p.zz.IsOne()) return HostPointJacobian<E>{p.x, p.y, E::One()};
  E z = p.zz.Mul(p.zzz);
  return HostPointJacobian<E>{p.x.Mul(p.zzz).Mul(z), p.y.Mul(p.zz).Mul(z.Sqr()), z};
}

// XYZZ -> affine for n points with ONE field inversion (Montgomery's trick), the
// BatchNormalize of short_weierstrass/point_xyzz.h:109-163: x / zz, y / zzz; the identity
// maps to (0, 0).
template <class E>
void BatchNormalize(const HostPointXYZZ<E>* in, size_t n, HostPointAffine<E>* out) {
  if (n == 0) return;
  E* prefix = new E[n];
  E acc = E::One();
  for (size_t i = 0; i < n; ++i) {
    prefix[i] = acc;
    if (!in[i].IsZero()) acc = acc.Mul(in[i].zzz);
  }
  E inv = acc.Inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].IsZero()) {
      out[i].x = E::Zero();
      out[i].y = E::Zero();
      continue;
    }
    E zi3 = inv.Mul(prefix[i]);   // 1 / zzz_i
    inv = inv.Mul(in[i].zzz);
    E zi2 = zi3.Mul(in[i].zz).Sqr();  // (zz / zzz)^2 = 1 / zz
    out[i].x = in[i].x.Mul(zi2);
    out[i].y = in[i].y.Mul(zi3);
  }
  delete[] prefix;
}

// [k] p by double-and-add, k canonical little-endian u64 limbs (used for the r / s blinding
// terms of a Groth16 proof, zk/r1cs/groth16/prove.h:113-148).
template <class E>
HostPointXYZZ<E> ScalarMul(const HostPointXYZZ<E>& p, const uint64_t* k, int limbs) {
  HostPointXYZZ<E> acc = HostPointXYZZ<E>::Zero();
  for (int i = limbs * 64 - 1; i >= 0; --i) {
    acc = acc.Dbl();
    if ((k[i >> 6] >> (i & 63)) & 1) acc = acc.Add(p);
  }
  return acc;
}

template <class E>
HostPointXYZZ<E> FromAffine(const HostPointAffine<E>& a) {
  if (a.x.IsZero() && a.y.IsZero()) return HostPointXYZZ<E>::Zero();
  return HostPointXYZZ<E>{a.x, a.y, E::One(), E::One()};
}

// pippenger_base.h:59-77: Horner over window sums, c doublings per window.
template <class E>
HostPointXYZZ<E> CombineWindows(const HostPointXYZZ<E>* sums, uint32_t windows, uint32_t c) {
  HostPointXYZZ<E> total = HostPointXYZZ<E>::Zero();
  for (uint32_t w = windows; w-- > 1;) {
    total = total.Add(sums[w]);
    for (uint32_t i = 0; i < c; ++i) total = total.Dbl();
  }
  return sums[0].Add(total);
}

}  // namespace tb200
