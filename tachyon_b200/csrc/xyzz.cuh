// Short-Weierstrass (a = 0) points in XYZZ coordinates on the device.
//
// Formulas are the EFD ones the reference's CPU path uses
// (tachyon/math/elliptic_curves/short_weierstrass/point_xyzz_impl.h):
//   madd-2008-s  (:114-176)   XYZZ += affine        8M + 2S
//   add-2008-s   (:44-97)     XYZZ += XYZZ         12M + 2S
//   dbl-2008-s-1 (:199-236)   XYZZ  = 2 * XYZZ      6M + 3S (a = 0)
// with the same exceptional-case behaviour: identity = (zz == 0)
// (point_xyzz.h:193), affine identity = (0, 0) (affine_point.h:125),
// P == R == 0 -> doubling, P == 0 && R != 0 -> zz = 0 falls out of the formula.
//
// Everything is templated on a field kind K (fp.cuh: FpField<F> for G1, Fp2Field<F> for
// G2), so the same formulas serve both groups, as the reference's templates do.
#pragma once
#include "fp.cuh"

namespace tb200 {

template <class K>
struct Affine {
  typename K::El x, y;
};

template <class K>
struct XYZZ {
  typename K::El x, y, zz, zzz;
};

template <class K>
TB_DEV void xyzz_set_zero(XYZZ<K>& p) {
  K::set_one(p.x);
  K::set_one(p.y);
  K::set_zero(p.zz);
  K::set_zero(p.zzz);
}
template <class K>
TB_DEV bool xyzz_is_zero(const XYZZ<K>& p) {
  return K::is_zero(p.zz);
}
template <class K>
TB_DEV bool affine_is_zero(const Affine<K>& p) {
  return K::is_zero(p.x) && K::is_zero(p.y);
}

template <class K>
TB_DEV void affine_load(Affine<K>& p, const void* src) {
  K::load(p.x, src);
  K::load(p.y, reinterpret_cast<const char*>(src) + sizeof(uint32_t) * K::kWords);
}
template <class K>
TB_DEV void affine_store(void* dst, const Affine<K>& p) {
  K::store(dst, p.x);
  K::store(reinterpret_cast<char*>(dst) + sizeof(uint32_t) * K::kWords, p.y);
}
template <class K>
TB_DEV void xyzz_load(XYZZ<K>& p, const void* src) {
  const char* s = reinterpret_cast<const char*>(src);
  constexpr int B = sizeof(uint32_t) * K::kWords;
  K::load_rw(p.x, s);
  K::load_rw(p.y, s + B);
  K::load_rw(p.zz, s + 2 * B);
  K::load_rw(p.zzz, s + 3 * B);
}
template <class K>
TB_DEV void xyzz_store(void* dst, const XYZZ<K>& p) {
  char* d = reinterpret_cast<char*>(dst);
  constexpr int B = sizeof(uint32_t) * K::kWords;
  K::store(d, p.x);
  K::store(d + B, p.y);
  K::store(d + 2 * B, p.zz);
  K::store(d + 3 * B, p.zzz);
}

// dbl-2008-s-1 with a = 0; p must not be the identity.
template <class K>
__device__ __noinline__ void xyzz_dbl_nz(XYZZ<K>& p) {
  typename K::El u, v, w, s, m, t;
  K::dbl(u, p.y);       // U = 2 Y1
  K::sqr(v, u);         // V = U^2
  K::mul(w, u, v);      // W = U V
  K::mul(s, p.x, v);    // S = X1 V
  K::sqr(m, p.x);       // M = 3 X1^2
  K::dbl(t, m);
  K::add(m, m, t);
  K::mul(p.zz, v, p.zz);    // ZZ3 = V ZZ1
  K::mul(p.zzz, w, p.zzz);  // ZZZ3 = W ZZZ1
  K::sqr(p.x, m);           // X3 = M^2 - 2 S
  K::dbl(t, s);
  K::sub(p.x, p.x, t);
  K::sub(t, p.x, s);        // Y3 = M (S - X3) - W Y1 = -(W Y1 + M (X3 - S)), one reduction
  K::mul2(t, w, p.y, m, t);
  K::neg(p.y, t);
}
template <class K>
TB_DEV void xyzz_dbl(XYZZ<K>& p) {
  if (!xyzz_is_zero<K>(p)) xyzz_dbl_nz<K>(p);
}

// acc += (neg ? -q : q),  q affine.  madd-2008-s.
template <class K>
TB_DEV void xyzz_madd(XYZZ<K>& acc, const Affine<K>& q, bool neg) {
  if (affine_is_zero<K>(q)) return;
  typename K::El y2;
  K::cneg(y2, q.y, neg);
  if (xyzz_is_zero<K>(acc)) {
    acc.x = q.x;
    acc.y = y2;
    K::set_one(acc.zz);
    K::set_one(acc.zzz);
    return;
  }
  typename K::El p, r, pp, ppp, qq, t;
  K::mul(p, q.x, acc.zz);  // P = X2 ZZ1 - X1
  K::sub(p, p, acc.x);
  K::mul(r, y2, acc.zzz);  // R = Y2 ZZZ1 - Y1
  K::sub(r, r, acc.y);
  if (K::is_zero(p) && K::is_zero(r)) {
    xyzz_dbl_nz<K>(acc);
    return;
  }
  K::sqr(pp, p);                  // PP = P^2
  K::mul(ppp, p, pp);             // PPP = P PP
  K::mul(qq, acc.x, pp);          // Q = X1 PP
  K::mul(acc.zz, acc.zz, pp);     // ZZ3 = ZZ1 PP
  K::mul(acc.zzz, acc.zzz, ppp);  // ZZZ3 = ZZZ1 PPP
  K::sqr(acc.x, r);               // X3 = R^2 - PPP - 2Q
  K::sub(acc.x, acc.x, ppp);
  K::dbl(t, qq);
  K::sub(acc.x, acc.x, t);
  K::sub(t, acc.x, qq);           // Y3 = R (Q - X3) - Y1 PPP = -(Y1 PPP + R (X3 - Q)),
  K::mul2(t, acc.y, ppp, r, t);   //      one Montgomery reduction for both products
  K::neg(acc.y, t);
}

// acc += b for acc, b both != identity (add-2008-s).  Straight-line body: the
// identity cases and the P == R == 0 doubling are dispatched by the inlined
// wrapper below, so this out-of-line function has no divergent early exit.
// Returns true when the operands are equal (caller must double instead; acc is
// left untouched in that case).
template <class K>
TB_DEV bool xyzz_add_nz_body(XYZZ<K>& acc, const XYZZ<K>& b) {
  typename K::El u1, s1, p, r, pp, ppp, qq, t;
  K::mul(u1, acc.x, b.zz);   // U1 = X1 ZZ2
  K::mul(s1, acc.y, b.zzz);  // S1 = Y1 ZZZ2
  K::mul(p, b.x, acc.zz);    // P = X2 ZZ1 - U1
  K::sub(p, p, u1);
  K::mul(r, b.y, acc.zzz);   // R = Y2 ZZZ1 - S1
  K::sub(r, r, s1);
  bool same = K::is_zero(p) && K::is_zero(r);
  K::sqr(pp, p);
  K::mul(ppp, p, pp);
  K::mul(qq, u1, pp);              // Q = U1 PP
  XYZZ<K> o;
  K::mul(o.zz, acc.zz, b.zz);      // ZZ3 = ZZ1 ZZ2 PP
  K::mul(o.zz, o.zz, pp);
  K::mul(o.zzz, acc.zzz, b.zzz);   // ZZZ3 = ZZZ1 ZZZ2 PPP
  K::mul(o.zzz, o.zzz, ppp);
  K::sqr(o.x, r);                  // X3 = R^2 - PPP - 2Q
  K::sub(o.x, o.x, ppp);
  K::dbl(t, qq);
  K::sub(o.x, o.x, t);
  K::sub(t, o.x, qq);              // Y3 = R (Q - X3) - S1 PPP = -(S1 PPP + R (X3 - Q))
  K::mul2(t, s1, ppp, r, t);
  K::neg(o.y, t);
  K::select(acc.x, same, acc.x, o.x);
  K::select(acc.y, same, acc.y, o.y);
  K::select(acc.zz, same, acc.zz, o.zz);
  K::select(acc.zzz, same, acc.zzz, o.zzz);
  return same;
}
template <class K>
__device__ __noinline__ bool xyzz_add_nz(XYZZ<K>& acc, const XYZZ<K>& b) {
  return xyzz_add_nz_body<K>(acc, b);
}

template <class K>
TB_DEV void xyzz_copy(XYZZ<K>& dst, const XYZZ<K>& src) {
  dst = src;
}

// acc += b with the reference's case analysis (point_xyzz_impl.h:14-41).
template <class K>
TB_DEV void xyzz_add(XYZZ<K>& acc, const XYZZ<K>& b) {
  bool bz = xyzz_is_zero<K>(b), az = xyzz_is_zero<K>(acc);
  if (az && !bz) xyzz_copy<K>(acc, b);
  if (!az && !bz) {
    if (xyzz_add_nz<K>(acc, b)) xyzz_dbl_nz<K>(acc);
  }
}

// xyzz_add with the addition itself inlined at the call site (for kernels that have exactly one)
template <class K>
TB_DEV void xyzz_add_inlined(XYZZ<K>& acc, const XYZZ<K>& b) {
  bool bz = xyzz_is_zero<K>(b), az = xyzz_is_zero<K>(acc);
  if (az && !bz) xyzz_copy<K>(acc, b);
  if (!az && !bz) {
    if (xyzz_add_nz_body<K>(acc, b)) xyzz_dbl_nz<K>(acc);
  }
}

// ---------------------------------------------------------------------------
// Four-lane cooperative point operations for the LATENCY-bound parts of an MSM (merge tree of
// the bucket reduction, window combination): one lane per independent field multiplication of
// a formula level, results exchanged with warp shuffles.  A lane running a point doubling alone
// executes its 9 multiplications back to back (~4 us for 254-bit limbs: every IMAD.WIDE waits
// for the carry of the previous one); four lanes do the same doubling in 3 multiplication
// levels, an addition in 4 instead of 14 — the same number of multiply instructions per point
// operation, spread over lanes that would otherwise idle.
//
// All four lanes of a group hold the same operands and end with the same result.  `lane` is
// the lane's index inside its group (0..3), `mask` the group's four lanes; branches are
// uniform inside a group, so groups of one warp may diverge from each other.
// ---------------------------------------------------------------------------
template <class K>
struct Coop4 {
  using El = typename K::El;

  static TB_DEV void pick(El& r, uint32_t lane, const El& a0, const El& a1, const El& a2,
                          const El& a3) {
    El lo, hi;
    K::select(lo, (lane & 1) == 0, a0, a1);
    K::select(hi, (lane & 1) == 0, a2, a3);
    K::select(r, lane < 2, lo, hi);
  }

  // p = 2 p, p != identity  (dbl-2008-s-1, a = 0: the levels of xyzz_dbl_nz)
  static TB_DEV void dbl_nz(XYZZ<K>& p, uint32_t lane, uint32_t mask) {
    El u, a, b, r, v, xx, m, t, w, s, mm;
    K::dbl(u, p.y);
    // level 1: V = U^2 | XX = X^2
    K::select(a, (lane & 1) == 0, u, p.x);
    K::sqr(r, a);
    K::shfl(v, r, mask, 0, 4);
    K::shfl(xx, r, mask, 1, 4);
    K::dbl(t, xx);
    K::add(m, xx, t);  // M = 3 X^2
    // level 2: W = U V | S = X V | ZZ3 = V ZZ | MM = M M
    pick(a, lane, u, p.x, v, m);
    pick(b, lane, v, v, p.zz, m);
    K::mul(r, a, b);
    K::shfl(w, r, mask, 0, 4);
    K::shfl(s, r, mask, 1, 4);
    K::shfl(p.zz, r, mask, 2, 4);
    K::shfl(mm, r, mask, 3, 4);
    K::dbl(t, s);
    K::sub(mm, mm, t);  // X3 = M^2 - 2 S
    K::sub(t, s, mm);   // S - X3
    // level 3: ZZZ3 = W ZZZ | WY = W Y | MT = M (S - X3)
    pick(a, lane, w, w, m, m);
    pick(b, lane, p.zzz, p.y, t, t);
    K::mul(r, a, b);
    K::shfl(p.zzz, r, mask, 0, 4);
    K::shfl(a, r, mask, 1, 4);
    K::shfl(b, r, mask, 2, 4);
    p.x = mm;
    K::sub(p.y, b, a);  // Y3 = M (S - X3) - W Y1
  }

  // acc += b with the reference's case analysis (point_xyzz_impl.h:14-97, add-2008-s)
  static TB_DEV void add(XYZZ<K>& acc, const XYZZ<K>& b, uint32_t lane, uint32_t mask) {
    const bool az = xyzz_is_zero<K>(acc), bz = xyzz_is_zero<K>(b);
    if (az || bz) {  // uniform inside the group
      if (az) acc = b;
      return;
    }
    El x, y, r, u1, s1, u2, s2, p, rr, pp, zz12, zzz12, ppp, q, t;
    // level 1: U1 = X1 ZZ2 | S1 = Y1 ZZZ2 | U2 = X2 ZZ1 | S2 = Y2 ZZZ1
    pick(x, lane, acc.x, acc.y, b.x, b.y);
    pick(y, lane, b.zz, b.zzz, acc.zz, acc.zzz);
    K::mul(r, x, y);
    K::shfl(u1, r, mask, 0, 4);
    K::shfl(s1, r, mask, 1, 4);
    K::shfl(u2, r, mask, 2, 4);
    K::shfl(s2, r, mask, 3, 4);
    K::sub(p, u2, u1);
    K::sub(rr, s2, s1);
    if (K::is_zero(p) && K::is_zero(rr)) {  // equal points: double
      dbl_nz(acc, lane, mask);
      return;
    }
    // level 2: PP = P^2 | RR = R^2 | ZZ12 = ZZ1 ZZ2 | ZZZ12 = ZZZ1 ZZZ2
    pick(x, lane, p, rr, acc.zz, acc.zzz);
    pick(y, lane, p, rr, b.zz, b.zzz);
    K::mul(r, x, y);
    K::shfl(pp, r, mask, 0, 4);
    K::shfl(t, r, mask, 1, 4);  // R^2
    K::shfl(zz12, r, mask, 2, 4);
    K::shfl(zzz12, r, mask, 3, 4);
    // level 3: PPP = P PP | Q = U1 PP | ZZ3 = ZZ12 PP
    pick(x, lane, p, u1, zz12, zz12);
    K::mul(r, x, pp);
    K::shfl(ppp, r, mask, 0, 4);
    K::shfl(q, r, mask, 1, 4);
    K::shfl(acc.zz, r, mask, 2, 4);
    K::sub(t, t, ppp);  // X3 = R^2 - PPP - 2 Q
    K::dbl(x, q);
    K::sub(t, t, x);
    K::sub(q, q, t);    // Q - X3
    // level 4: ZZZ3 = ZZZ12 PPP | S1 PPP | R (Q - X3)
    pick(x, lane, zzz12, s1, rr, rr);
    pick(y, lane, ppp, ppp, q, q);
    K::mul(r, x, y);
    K::shfl(acc.zzz, r, mask, 0, 4);
    K::shfl(x, r, mask, 1, 4);
    K::shfl(y, r, mask, 2, 4);
    acc.x = t;
    K::sub(acc.y, y, x);  // Y3 = R (Q - X3) - S1 PPP
  }
};

// ---------------------------------------------------------------------------
// G2 point operations with every Fq2 coordinate split over a LANE PAIR (fp.cuh, Fp2Lanes): a
// lane holds its component of x, y, zz, zzz.  Same formulas and the same case analysis as
// xyzz_madd / xyzz_dbl_nz above, but written for warps that stay CONVERGED: the general
// formula always runs and the exceptional cases (identity operands, inactive lane pairs, the
// P == R == 0 doubling) are selects or warp-uniform branches, so every shuffle inside is a
// plain full-mask SHFL.
// ---------------------------------------------------------------------------
template <class F>
struct PairPoint {
  Fp<F> x, y, zz, zzz;  // this lane's components
};
template <class F>
struct PairAffine {
  Fp<F> x, y;
};

template <class F, int kRoll = 0>
struct PairOps {
  using L = Fp2Lanes<F, kRoll>;
  using E = Fp<F>;

  // p = 2 p for the lane pairs with `take` set (p != identity there); all lanes compute.
  static __device__ __noinline__ void dbl_where(PairPoint<F>& p, bool take, uint32_t role) {
    E u, v, w, s, m, t, nzz, nzzz, nx;
    fp_dbl<F>(u, p.y);         // U = 2 Y1
    L::sqr(v, u, role);        // V = U^2
    typename L::Multiplier mv;
    L::prepare(mv, v, role);
    L::mul(w, u, mv);          // W = U V
    L::mul(s, p.x, mv);        // S = X1 V
    L::mul(nzz, p.zz, mv);     // ZZ3 = V ZZ1
    L::sqr(m, p.x, role);      // M = 3 X1^2
    fp_dbl<F>(t, m);
    fp_add<F>(m, m, t);
    typename L::Multiplier mw;
    L::prepare(mw, w, role);
    L::mul(nzzz, p.zzz, mw);   // ZZZ3 = W ZZZ1
    L::sqr(nx, m, role);       // X3 = M^2 - 2 S
    fp_dbl<F>(t, s);
    fp_sub<F>(nx, nx, t);
    fp_sub<F>(t, s, nx);       // Y3 = M (S - X3) - W Y1
    L::mul(t, t, m, role);
    L::mul(u, p.y, mw);
    fp_sub<F>(t, t, u);
    L::select(p.x, take, nx, p.x);
    L::select(p.y, take, t, p.y);
    L::select(p.zz, take, nzz, p.zz);
    L::select(p.zzz, take, nzzz, p.zzz);
  }

  // acc += b for the lane pairs with `active` set (add-2008-s, case analysis of
  // point_xyzz_impl.h:14-97); all lanes compute, warp-uniform branches only.
  static TB_DEV void add(PairPoint<F>& acc, const PairPoint<F>& b, bool active, uint32_t role) {
    const bool az = L::is_zero(acc.zz), bz = L::is_zero(b.zz);
    const bool copy = active && az && !bz;
    const bool plain = active && !az && !bz;
    if (__any_sync(L::kFull, copy)) {
      L::select(acc.x, copy, b.x, acc.x);
      L::select(acc.y, copy, b.y, acc.y);
      L::select(acc.zz, copy, b.zz, acc.zz);
      L::select(acc.zzz, copy, b.zzz, acc.zzz);
    }
    if (__any_sync(L::kFull, plain)) {
      E u1, s1, p, r, pp, ppp, qq, t, u, nx, nzz, nzzz;
      {
        typename L::Multiplier m;
        L::prepare(m, b.zz, role);
        L::mul(u1, acc.x, m);            // U1 = X1 ZZ2
        L::mul(nzz, acc.zz, m);          // ZZ1 ZZ2
        L::prepare(m, b.zzz, role);
        L::mul(s1, acc.y, m);            // S1 = Y1 ZZZ2
        L::mul(nzzz, acc.zzz, m);        // ZZZ1 ZZZ2
      }
      L::mul(p, b.x, acc.zz, role);      // P = X2 ZZ1 - U1
      fp_sub<F>(p, p, u1);
      L::mul(r, b.y, acc.zzz, role);     // R = Y2 ZZZ1 - S1
      fp_sub<F>(r, r, s1);
      const bool pr_zero = L::both(fp_is_zero<F>(p) && fp_is_zero<F>(r));  // every lane takes part
      const bool same = plain && pr_zero;
      const bool take = plain && !same;
      L::sqr(pp, p, role);               // PP = P^2
      {
        typename L::Multiplier m;
        L::prepare(m, pp, role);
        L::mul(ppp, p, m);               // PPP = P PP
        L::mul(qq, u1, m);               // Q = U1 PP
        L::mul(nzz, nzz, m);             // ZZ3 = ZZ1 ZZ2 PP
        L::prepare(m, ppp, role);
        L::mul(nzzz, nzzz, m);           // ZZZ3 = ZZZ1 ZZZ2 PPP
        L::mul(u, s1, m);                // S1 PPP
      }
      L::select(acc.zz, take, nzz, acc.zz);
      L::select(acc.zzz, take, nzzz, acc.zzz);
      L::sqr(nx, r, role);               // X3 = R^2 - PPP - 2 Q
      fp_sub<F>(nx, nx, ppp);
      fp_dbl<F>(t, qq);
      fp_sub<F>(nx, nx, t);
      fp_sub<F>(t, qq, nx);              // Y3 = R (Q - X3) - S1 PPP
      L::mul(t, t, r, role);
      fp_sub<F>(t, t, u);
      L::select(acc.x, take, nx, acc.x);
      L::select(acc.y, take, t, acc.y);
      // equal operands: double instead (those pairs' acc is still untouched)
      if (__any_sync(L::kFull, same)) dbl_where(acc, same, role);
    }
  }

  // acc += (neg ? -q : q) for the lane pairs with `active` set; q affine.  madd-2008-s with
  // the case analysis of point_xyzz_impl.h:114-176.
  static TB_DEV void madd(PairPoint<F>& acc, const PairAffine<F>& q, bool neg, bool active,
                          uint32_t role) {
    const bool qz = L::both(fp_is_zero<F>(q.x) && fp_is_zero<F>(q.y));
    const bool az = L::is_zero(acc.zz);
    const bool adds = active && !qz;  // something is added at all
    const bool plain = adds && !az;   // ... to a non-identity accumulator: the formula
    const bool copy = adds && az;     // ... to the identity: acc = q
    E y2;
    fp_cneg<F>(y2, q.y, neg);
    // Both blocks are warp-uniform branches.  The copy comes first so that q is not live across
    // the formula; the formula's selects skip the pairs that copied (`plain` is false there).
    if (__any_sync(L::kFull, copy)) {  // first point of fresh tasks, mostly
      E one;
      L::set_one(one, role);
      L::select(acc.x, copy, q.x, acc.x);
      L::select(acc.y, copy, y2, acc.y);
      L::select(acc.zz, copy, one, acc.zz);
      L::select(acc.zzz, copy, one, acc.zzz);
    }
    if (__any_sync(L::kFull, plain)) {
      E p, r, pp, ppp, qq, t, u, nx, nzz, nzzz;
      L::mul(p, q.x, acc.zz, role);       // P = X2 ZZ1 - X1
      fp_sub<F>(p, p, acc.x);
      L::mul(r, y2, acc.zzz, role);       // R = Y2 ZZZ1 - Y1
      fp_sub<F>(r, r, acc.y);
      const bool pr_zero = L::both(fp_is_zero<F>(p) && fp_is_zero<F>(r));  // every lane takes part
      const bool same = plain && pr_zero;
      const bool take = plain && !same;
      L::sqr(pp, p, role);                // PP = P^2
      {
        typename L::Multiplier mpp;
        L::prepare(mpp, pp, role);
        L::mul(ppp, p, mpp);              // PPP = P PP
        L::mul(qq, acc.x, mpp);           // Q = X1 PP
        L::mul(nzz, acc.zz, mpp);         // ZZ3 = ZZ1 PP
      }
      L::select(acc.zz, take, nzz, acc.zz);
      {
        typename L::Multiplier mppp;
        L::prepare(mppp, ppp, role);
        L::mul(nzzz, acc.zzz, mppp);      // ZZZ3 = ZZZ1 PPP
        L::mul(u, acc.y, mppp);           // Y1 PPP
      }
      L::select(acc.zzz, take, nzzz, acc.zzz);
      L::sqr(nx, r, role);                // X3 = R^2 - PPP - 2 Q
      fp_sub<F>(nx, nx, ppp);
      fp_dbl<F>(t, qq);
      fp_sub<F>(nx, nx, t);
      fp_sub<F>(t, qq, nx);               // Y3 = R (Q - X3) - Y1 PPP
      L::mul(t, t, r, role);
      fp_sub<F>(t, t, u);
      L::select(acc.x, take, nx, acc.x);
      L::select(acc.y, take, t, acc.y);
      // P == R == 0: the operands are equal, double instead (those pairs' acc is still untouched)
      if (__any_sync(L::kFull, same)) dbl_where(acc, same, role);
    }
  }
};

}  // namespace tb200
