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
#pragma once
#include "fp.cuh"

namespace tb200 {

template <class F>
struct Affine {
  Fp<F> x, y;
};

template <class F>
struct XYZZ {
  Fp<F> x, y, zz, zzz;
};

template <class F>
TB_DEV void xyzz_set_zero(XYZZ<F>& p) {
  fp_set_one<F>(p.x);
  fp_set_one<F>(p.y);
  fp_set_zero<F>(p.zz);
  fp_set_zero<F>(p.zzz);
}
template <class F>
TB_DEV bool xyzz_is_zero(const XYZZ<F>& p) {
  return fp_is_zero<F>(p.zz);
}
template <class F>
TB_DEV bool affine_is_zero(const Affine<F>& p) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) acc |= p.x.l[i] | p.y.l[i];
  return acc == 0;
}

template <class F>
TB_DEV void affine_load(Affine<F>& p, const void* src) {
  fp_load<F>(p.x, src);
  fp_load<F>(p.y, reinterpret_cast<const char*>(src) + sizeof(uint32_t) * Fp<F>::N);
}
template <class F>
TB_DEV void xyzz_load(XYZZ<F>& p, const void* src) {
  const char* s = reinterpret_cast<const char*>(src);
  constexpr int B = sizeof(uint32_t) * Fp<F>::N;
  fp_load_rw<F>(p.x, s);
  fp_load_rw<F>(p.y, s + B);
  fp_load_rw<F>(p.zz, s + 2 * B);
  fp_load_rw<F>(p.zzz, s + 3 * B);
}
template <class F>
TB_DEV void xyzz_store(void* dst, const XYZZ<F>& p) {
  char* d = reinterpret_cast<char*>(dst);
  constexpr int B = sizeof(uint32_t) * Fp<F>::N;
  fp_store<F>(d, p.x);
  fp_store<F>(d + B, p.y);
  fp_store<F>(d + 2 * B, p.zz);
  fp_store<F>(d + 3 * B, p.zzz);
}

// dbl-2008-s-1 with a = 0; p must not be the identity.
template <class F>
__device__ __noinline__ void xyzz_dbl_nz(XYZZ<F>& p) {
  Fp<F> u, v, w, s, m, t;
  fp_dbl<F>(u, p.y);       // U = 2 Y1
  fp_sqr<F>(v, u);         // V = U^2
  fp_mul<F>(w, u, v);      // W = U V
  fp_mul<F>(s, p.x, v);    // S = X1 V
  fp_sqr<F>(m, p.x);       // M = 3 X1^2
  fp_dbl<F>(t, m);
  fp_add<F>(m, m, t);
  fp_mul<F>(p.zz, v, p.zz);    // ZZ3 = V ZZ1
  fp_mul<F>(p.zzz, w, p.zzz);  // ZZZ3 = W ZZZ1
  fp_sqr<F>(p.x, m);           // X3 = M^2 - 2 S
  fp_dbl<F>(t, s);
  fp_sub<F>(p.x, p.x, t);
  fp_sub<F>(t, p.x, s);        // Y3 = M (S - X3) - W Y1 = -(W Y1 + M (X3 - S)), one reduction
  fp_mul2<F>(t, w, p.y, m, t);
  fp_neg<F>(p.y, t);
}
template <class F>
TB_DEV void xyzz_dbl(XYZZ<F>& p) {
  if (!xyzz_is_zero<F>(p)) xyzz_dbl_nz<F>(p);
}

// acc += (neg ? -q : q),  q affine.  madd-2008-s.
template <class F>
TB_DEV void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q, bool neg) {
  if (affine_is_zero<F>(q)) return;
  Fp<F> y2;
  fp_cneg<F>(y2, q.y, neg);
  if (xyzz_is_zero<F>(acc)) {
    acc.x = q.x;
    acc.y = y2;
    fp_set_one<F>(acc.zz);
    fp_set_one<F>(acc.zzz);
    return;
  }
  Fp<F> p, r, pp, ppp, qq, t;
  fp_mul<F>(p, q.x, acc.zz);  // P = X2 ZZ1 - X1
  fp_sub<F>(p, p, acc.x);
  fp_mul<F>(r, y2, acc.zzz);  // R = Y2 ZZZ1 - Y1
  fp_sub<F>(r, r, acc.y);
  if (fp_is_zero<F>(p) && fp_is_zero<F>(r)) {
    xyzz_dbl_nz<F>(acc);
    return;
  }
  fp_sqr<F>(pp, p);                  // PP = P^2
  fp_mul<F>(ppp, p, pp);             // PPP = P PP
  fp_mul<F>(qq, acc.x, pp);          // Q = X1 PP
  fp_mul<F>(acc.zz, acc.zz, pp);     // ZZ3 = ZZ1 PP
  fp_mul<F>(acc.zzz, acc.zzz, ppp);  // ZZZ3 = ZZZ1 PPP
  fp_sqr<F>(acc.x, r);               // X3 = R^2 - PPP - 2Q
  fp_sub<F>(acc.x, acc.x, ppp);
  fp_dbl<F>(t, qq);
  fp_sub<F>(acc.x, acc.x, t);
  fp_sub<F>(t, acc.x, qq);           // Y3 = R (Q - X3) - Y1 PPP = -(Y1 PPP + R (X3 - Q)),
  fp_mul2<F>(t, acc.y, ppp, r, t);   //      one Montgomery reduction for both products
  fp_neg<F>(acc.y, t);
}

// acc += b for acc, b both != identity (add-2008-s).  Straight-line body: the
// identity cases and the P == R == 0 doubling are dispatched by the inlined
// wrapper below, so this out-of-line function has no divergent early exit.
// Returns true when the operands are equal (caller must double instead; acc is
// left untouched in that case).
template <class F>
__device__ __noinline__ bool xyzz_add_nz(XYZZ<F>& acc, const XYZZ<F>& b) {
  Fp<F> u1, s1, p, r, pp, ppp, qq, t;
  fp_mul<F>(u1, acc.x, b.zz);   // U1 = X1 ZZ2
  fp_mul<F>(s1, acc.y, b.zzz);  // S1 = Y1 ZZZ2
  fp_mul<F>(p, b.x, acc.zz);    // P = X2 ZZ1 - U1
  fp_sub<F>(p, p, u1);
  fp_mul<F>(r, b.y, acc.zzz);   // R = Y2 ZZZ1 - S1
  fp_sub<F>(r, r, s1);
  bool same = fp_is_zero<F>(p) && fp_is_zero<F>(r);
  fp_sqr<F>(pp, p);
  fp_mul<F>(ppp, p, pp);
  fp_mul<F>(qq, u1, pp);              // Q = U1 PP
  XYZZ<F> o;
  fp_mul<F>(o.zz, acc.zz, b.zz);      // ZZ3 = ZZ1 ZZ2 PP
  fp_mul<F>(o.zz, o.zz, pp);
  fp_mul<F>(o.zzz, acc.zzz, b.zzz);   // ZZZ3 = ZZZ1 ZZZ2 PPP
  fp_mul<F>(o.zzz, o.zzz, ppp);
  fp_sqr<F>(o.x, r);                  // X3 = R^2 - PPP - 2Q
  fp_sub<F>(o.x, o.x, ppp);
  fp_dbl<F>(t, qq);
  fp_sub<F>(o.x, o.x, t);
  fp_sub<F>(t, o.x, qq);              // Y3 = R (Q - X3) - S1 PPP = -(S1 PPP + R (X3 - Q))
  fp_mul2<F>(t, s1, ppp, r, t);
  fp_neg<F>(o.y, t);
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) {
    acc.x.l[i] = same ? acc.x.l[i] : o.x.l[i];
    acc.y.l[i] = same ? acc.y.l[i] : o.y.l[i];
    acc.zz.l[i] = same ? acc.zz.l[i] : o.zz.l[i];
    acc.zzz.l[i] = same ? acc.zzz.l[i] : o.zzz.l[i];
  }
  return same;
}

template <class F>
TB_DEV void xyzz_copy(XYZZ<F>& dst, const XYZZ<F>& src) {
#pragma unroll
  for (int i = 0; i < Fp<F>::N; ++i) {
    dst.x.l[i] = src.x.l[i];
    dst.y.l[i] = src.y.l[i];
    dst.zz.l[i] = src.zz.l[i];
    dst.zzz.l[i] = src.zzz.l[i];
  }
}

// acc += b with the reference's case analysis (point_xyzz_impl.h:14-41).
template <class F>
TB_DEV void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& b) {
  bool bz = xyzz_is_zero<F>(b), az = xyzz_is_zero<F>(acc);
  if (az && !bz) xyzz_copy<F>(acc, b);
  if (!az && !bz) {
    if (xyzz_add_nz<F>(acc, b)) xyzz_dbl_nz<F>(acc);
  }
}

}  // namespace tb200
