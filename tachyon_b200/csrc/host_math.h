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

template <class F>
struct HostXYZZ {
  HostFp<F> x, y, zz, zzz;

  static HostXYZZ Zero() {
    return HostXYZZ{HostFp<F>::One(), HostFp<F>::One(), HostFp<F>::Zero(), HostFp<F>::Zero()};
  }
  bool IsZero() const { return zz.IsZero(); }

  HostXYZZ Dbl() const {
    if (IsZero()) return *this;
    HostFp<F> u = y.Dbl(), v = u.Sqr(), w = u.Mul(v), s = x.Mul(v);
    HostFp<F> m = x.Sqr();
    m = m.Add(m.Dbl());
    HostXYZZ r;
    r.x = m.Sqr().Sub(s.Dbl());
    r.y = m.Mul(s.Sub(r.x)).Sub(w.Mul(y));
    r.zz = v.Mul(zz);
    r.zzz = w.Mul(zzz);
    return r;
  }
  HostXYZZ Add(const HostXYZZ& b) const {
    if (IsZero()) return b;
    if (b.IsZero()) return *this;
    HostFp<F> u1 = x.Mul(b.zz), s1 = y.Mul(b.zzz);
    HostFp<F> p = b.x.Mul(zz).Sub(u1), r = b.y.Mul(zzz).Sub(s1);
    if (p.IsZero() && r.IsZero()) return Dbl();
    HostFp<F> pp = p.Sqr(), ppp = p.Mul(pp), q = u1.Mul(pp);
    HostXYZZ c;
    c.x = r.Sqr().Sub(ppp).Sub(q.Dbl());
    c.y = r.Mul(q.Sub(c.x)).Sub(s1.Mul(ppp));
    c.zz = zz.Mul(b.zz).Mul(pp);
    c.zzz = zzz.Mul(b.zzz).Mul(ppp);
    return c;
  }
};

template <class F>
struct HostJacobian {
  HostFp<F> x, y, z;
};

// point_xyzz.h:228-237
template <class F>
HostJacobian<F> ToJacobian(const HostXYZZ<F>& p) {
  if (p.IsZero()) return HostJacobian<F>{HostFp<F>::One(), HostFp<F>::One(), HostFp<F>::Zero()};
  if (p.zz.IsOne()) return HostJacobian<F>{p.x, p.y, HostFp<F>::One()};
  HostFp<F> z = p.zz.Mul(p.zzz);
  return HostJacobian<F>{p.x.Mul(p.zzz).Mul(z), p.y.Mul(p.zz).Mul(z.Sqr()), z};
}

template <class F>
struct HostAffine {
  HostFp<F> x, y;
};

// XYZZ -> affine for n points with ONE field inversion (Montgomery's trick), the
// BatchNormalize of short_weierstrass/point_xyzz.h:109-163: x / zz, y / zzz; the identity
// maps to (0, 0).
template <class F>
void BatchNormalize(const HostXYZZ<F>* in, size_t n, HostAffine<F>* out) {
  if (n == 0) return;
  HostFp<F>* prefix = new HostFp<F>[n];
  HostFp<F> acc = HostFp<F>::One();
  for (size_t i = 0; i < n; ++i) {
    prefix[i] = acc;
    if (!in[i].IsZero()) acc = acc.Mul(in[i].zzz);
  }
  HostFp<F> inv = acc.Inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].IsZero()) {
      out[i].x = HostFp<F>::Zero();
      out[i].y = HostFp<F>::Zero();
      continue;
    }
    HostFp<F> zi3 = inv.Mul(prefix[i]);   // 1 / zzz_i
    inv = inv.Mul(in[i].zzz);
    HostFp<F> zi2 = zi3.Mul(in[i].zz).Sqr();  // (zz / zzz)^2 = 1 / zz
    out[i].x = in[i].x.Mul(zi2);
    out[i].y = in[i].y.Mul(zi3);
  }
  delete[] prefix;
}

// pippenger_base.h:59-77: Horner over window sums, c doublings per window.
template <class F>
HostXYZZ<F> CombineWindows(const HostXYZZ<F>* sums, uint32_t windows, uint32_t c) {
  HostXYZZ<F> total = HostXYZZ<F>::Zero();
  for (uint32_t w = windows; w-- > 1;) {
    total = total.Add(sums[w]);
    for (uint32_t i = 0; i < c; ++i) total = total.Dbl();
  }
  return sums[0].Add(total);
}

}  // namespace tb200
