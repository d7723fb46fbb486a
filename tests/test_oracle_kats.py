"""Pins the CPU oracle against everything the reference's own tests fix for the
MSM path: GF(7) point KATs, curve constants, the zkey decimal
KATs, and the reference's relation tests restated with fixed seeds.  CPU only.
"""
import json
import os

import numpy as np
import pytest

from oracle import cpu_oracle, pymodel

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- GF(7) KATs: short_weierstrass/point_xyzz_unittest.cc:77-166 ------------
def test_gf7_additive_group_operators():
    p, p2, p3, p4 = [5, 5, 1, 1], [3, 2, 1, 1], [3, 5, 1, 1], [6, 5, 1, 1]
    aff = lambda q: cpu_oracle.gf7("xyzz_to_affine", 2, q)
    # p + p2 == p3 ; p + p == p4 ; p.Double() == p4       (:88-91, :108)
    assert aff(cpu_oracle.gf7("xyzz_add", 4, p, p2)) == [3, 5]
    assert aff(cpu_oracle.gf7("xyzz_add", 4, p, p)) == [6, 5]
    assert aff(cpu_oracle.gf7("xyzz_double", 4, p)) == [6, 5]
    # mixed: p + ap2 == p3 ; p + ap == p4                  (:102-103)
    assert aff(cpu_oracle.gf7("xyzz_madd", 4, p, [3, 2])) == [3, 5]
    assert aff(cpu_oracle.gf7("xyzz_madd", 4, p, [5, 5])) == [6, 5]
    # p - p3 == -p2 ; p - p4 == -p                         (:89,:91)
    n3 = cpu_oracle.gf7("xyzz_neg", 4, p3)
    n4 = cpu_oracle.gf7("xyzz_neg", 4, p4)
    assert aff(cpu_oracle.gf7("xyzz_add", 4, p, n3)) == [3, 5]  # -p2 = (3, -2) = (3, 5)
    assert aff(cpu_oracle.gf7("xyzz_add", 4, p, n4)) == [5, 2]  # -p  = (5, 2)
    # -p == (5, 2, 1, 1)                                   (:115)
    assert cpu_oracle.gf7("xyzz_neg", 4, p) == [5, 2, 1, 1]
    # p + (-p) == identity
    assert aff(cpu_oracle.gf7("xyzz_add", 4, p, cpu_oracle.gf7("xyzz_neg", 4, p))) == [0, 0]
    assert p3 == [3, 5, 1, 1] and p4 == [6, 5, 1, 1]


def test_gf7_scalar_mul_orbit():
    # :128-142 — {k * G : k in 0..6} is exactly the 7-point group
    pts = sorted(tuple(cpu_oracle.gf7_scalar_mul([5, 5], k)) for k in range(7))
    assert pts == sorted([(0, 0), (3, 2), (5, 2), (6, 2), (3, 5), (5, 5), (6, 5)])


def test_gf7_conversions():
    # ToAffine :144-151 ; ToJacobian :162-169
    assert cpu_oracle.gf7("xyzz_to_affine", 2, [1, 2, 0, 0]) == [0, 0]
    assert cpu_oracle.gf7("xyzz_to_affine", 2, [1, 2, 1, 1]) == [1, 2]
    assert cpu_oracle.gf7("xyzz_to_affine", 2, [1, 2, 2, 6]) == [4, 5]
    assert cpu_oracle.gf7("xyzz_to_jacobian", 3, [1, 2, 0, 0])[2] == 0
    assert cpu_oracle.gf7("xyzz_to_jacobian", 3, [1, 2, 1, 1]) == [1, 2, 1]
    assert cpu_oracle.gf7("xyzz_to_jacobian", 3, [1, 2, 2, 6]) == [2, 2, 5]
    # the Jacobian image normalises to the same affine point
    assert cpu_oracle.gf7("jacobian_to_affine", 2, [2, 2, 5]) == [4, 5]


# ---- constants: BUILD.bazel moduli/generators + SURVEY §8c derived values ----
EXPECTED = {
    "bn254": dict(
        fq_r=0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d,
        fq_inv=0x87d20782e4866389,
        fr_r=0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb,
        fr_inv=0xc2e1f593efffffff),
    "bls12_381": dict(
        fq_r=0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd,
        fq_inv=0x89f3fffcfffcfffd,
        fr_r=0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe,
        fr_inv=0xfffffffeffffffff),
}


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_constants(oracles, name):
    o, c = oracles[name], pymodel.CURVES[name]
    k = o.constants()
    L = pymodel.from_limbs
    assert L(k["fq_mod"]) == c.p and L(k["fr_mod"]) == c.r
    assert L(k["fq_r"]) == c.fq_R % c.p == EXPECTED[name]["fq_r"]
    assert L(k["fr_r"]) == c.fr_R % c.r == EXPECTED[name]["fr_r"]
    assert L(k["fq_r2"]) == c.fq_R ** 2 % c.p and L(k["fr_r2"]) == c.fr_R ** 2 % c.r
    assert int(k["fq_inv"][0]) == (-pow(c.p, -1, 1 << 64)) % (1 << 64) == EXPECTED[name]["fq_inv"]
    assert int(k["fr_inv"][0]) == (-pow(c.r, -1, 1 << 64)) % (1 << 64) == EXPECTED[name]["fr_inv"]
    gx = pymodel.fq_from_mont(c, L(k["gen"][:c.fq_limbs]))
    gy = pymodel.fq_from_mont(c, L(k["gen"][c.fq_limbs:]))
    assert (gx, gy) == (c.gx, c.gy) and pymodel.is_on_curve(c, (gx, gy))
    assert pymodel.mul(c, c.r, (gx, gy)) is pymodel.INF


# ---- zkey decimal KATs pin the Montgomery byte layout -----------------------
def test_zkey_montgomery_layout(oracles):
    fx = json.load(open(os.path.join(HERE, "golden", "zkey_multiplier_3_g1.json")))
    o, c = oracles["bn254"], pymodel.BN254
    assert int(fx["q"]) == c.p and int(fx["r"]) == c.r
    for name, hx in fx["montgomery_bytes_hex"].items():
        if name.endswith("_g2"):
            continue  # the G2 header points are checked in tests/test_oracle_g2.py
        limbs = np.frombuffer(bytes.fromhex(hx), dtype="<u8").reshape(2, 4)
        canon = o.fq_from_mont(limbs)
        got = [pymodel.from_limbs(canon[0]), pymodel.from_limbs(canon[1])]
        assert got == [int(v) for v in fx["expected_decimal_xy"][name]], name
        assert pymodel.is_on_curve(c, tuple(got))


# ---- field + point arithmetic against the independent Python model ----------
def _rand_fq(c, rng, n):
    return [int.from_bytes(rng.bytes(8 * c.fq_limbs), "little") % c.p for _ in range(n)]


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_field_ops_vs_python(oracles, name):
    o, c = oracles[name], pymodel.CURVES[name]
    rng = np.random.default_rng(1234)
    xs = _rand_fq(c, rng, 64) + [0, 1, c.p - 1, c.p - 2]
    ys = _rand_fq(c, rng, 64) + [c.p - 1, 0, c.p - 1, 1]
    A = np.array([pymodel.to_limbs(pymodel.fq_to_mont(c, x), c.fq_limbs) for x in xs], dtype=np.uint64)
    B = np.array([pymodel.to_limbs(pymodel.fq_to_mont(c, y), c.fq_limbs) for y in ys], dtype=np.uint64)
    assert (o.fq_to_mont(np.array([pymodel.to_limbs(x, c.fq_limbs) for x in xs], dtype=np.uint64)) == A).all()
    dec = lambda arr: [pymodel.fq_from_mont(c, pymodel.from_limbs(r)) for r in arr]
    assert dec(o.fq_op("add", A, B)) == [(x + y) % c.p for x, y in zip(xs, ys)]
    assert dec(o.fq_op("sub", A, B)) == [(x - y) % c.p for x, y in zip(xs, ys)]
    assert dec(o.fq_op("mul", A, B)) == [(x * y) % c.p for x, y in zip(xs, ys)]
    assert dec(o.fq_op("square", A)) == [(x * x) % c.p for x in xs]
    assert dec(o.fq_op("neg", A)) == [(-x) % c.p for x in xs]
    assert dec(o.fq_op("double", A)) == [(2 * x) % c.p for x in xs]
    nz = A[[i for i, x in enumerate(xs) if x]]
    assert dec(o.fq_op("inverse", nz)) == [pow(x, -1, c.p) for x in xs if x]


def _decode_affine(c, a):
    x = pymodel.fq_from_mont(c, pymodel.from_limbs(a[0]))
    y = pymodel.fq_from_mont(c, pymodel.from_limbs(a[1]))
    return pymodel.INF if (x, y) == (0, 0) else (x, y)


def _decode_bases(c, bases):
    return [_decode_affine(c, b.reshape(2, c.fq_limbs)) for b in bases]


def _decode_scalars(c, scalars):
    return [pymodel.fr_from_mont(c, pymodel.from_limbs(s)) for s in scalars]


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_generated_points_and_scalars(oracles, name):
    o, c = oracles[name], pymodel.CURVES[name]
    pts = _decode_bases(c, o.generate_points(7, 40))
    assert all(pymodel.is_on_curve(c, p) and p is not pymodel.INF for p in pts)
    for a, b in zip(pts, pts[1:]):
        assert pymodel.add(c, a, a) == b          # doubling chain (test/random.h:22-25)
    # windows of the same stream agree
    assert (o.generate_points(7, 10, first=4090) == o.generate_points(7, 4100)[4090:]).all()
    for dist in ("uniform", "non_uniform", "witness"):
        s = o.generate_scalars(9, 64, dist)
        assert all(v < c.r for v in (pymodel.from_limbs(x) for x in s))
        assert (o.generate_scalars(9, 16, dist, first=48) == s[48:]).all()
    assert len({tuple(x) for x in o.generate_scalars(9, 64, "non_uniform")}) == 1


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_fill_digits(oracles, name):
    o, c = oracles[name], pymodel.CURVES[name]
    s = o.generate_scalars(11, 16)
    edge = np.array([pymodel.to_limbs(pymodel.fr_to_mont(c, v), 4) for v in (0, 1, c.r - 1, (1 << 253) - 1)],
                    dtype=np.uint64)
    bits = c.r.bit_length()
    for sc in list(s) + list(edge):
        k = pymodel.fr_from_mont(c, pymodel.from_limbs(sc))
        for cb in (3, 5, 8, 13, 16):
            W = (bits + cb - 1) // cb
            d = [int(v) for v in o.fill_digits(sc, cb, W)]
            assert d == pymodel.fill_digits(k, cb, W)
            assert sum(v << (cb * i) for i, v in enumerate(d)) == k
            assert all(-(1 << (cb - 1)) <= v < (1 << (cb - 1)) for v in d[:-1])


# ---- the reference's relation tests, restated with fixed seeds --------------
# pippenger_unittest.cc:40-70 (n = 40: MSM == naive), variable_base_msm_unittest.cc,
# c/.../msm_unittest.cc:39-61 (sizes 32, 2, 5)
@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
@pytest.mark.parametrize("n", [0, 1, 2, 5, 32, 40])
def test_msm_equals_naive_and_python(oracles, name, n):
    o, c = oracles[name], pymodel.CURVES[name]
    bases, scalars = o.generate_points(21 + n, n), o.generate_scalars(22 + n, n)
    want = pymodel.msm(c, _decode_bases(c, bases), _decode_scalars(c, scalars))
    naive = _decode_affine(c, o.xyzz_to_affine(o.msm_naive(bases, scalars)))
    assert naive == want
    for strat in ("none", "parallel_window", "parallel_term"):
        got = _decode_affine(c, o.msm_affine(bases, scalars, strategy=strat, threads=3))
        assert got == want, strat
    # Jacobian output (what the C API returns) normalises to the same point
    jac = o.xyzz_to_jacobian(o.msm(bases, scalars))
    assert _decode_affine(c, o.jacobian_to_affine(jac)) == want


# pippenger_adapter_unittest.cc:31-48 — n = 1024, every strategy agrees
@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_adapter_strategies_agree_1024(oracles, name):
    o = oracles[name]
    bases, scalars = o.generate_points(31, 1024), o.generate_scalars(32, 1024)
    ref = o.msm_affine(bases, scalars, strategy="none")
    for strat, th in (("parallel_window", 4), ("parallel_term", 1), ("parallel_term", 7), ("parallel_term", 64)):
        assert (o.msm_affine(bases, scalars, strategy=strat, threads=th) == ref).all()
    folded = o.fold_chain_scalars(scalars)
    assert (o.msm_affine(bases[::4096], folded) == ref).all()


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_edge_cases(oracles, name):
    o, c = oracles[name], pymodel.CURVES[name]
    n = 48
    bases, scalars = o.generate_points(41, n), o.generate_scalars(42, n, "witness")
    bases[3] = 0                                   # identity base (0,0): affine_point.h:125
    bases[9] = bases[8]                            # duplicate point -> doubling branch
    scalars[9] = scalars[8]
    neg_y = o.fq_op("neg", bases[10].reshape(2, -1)[1:2])
    bases[11] = np.concatenate([bases[10][:c.fq_limbs], neg_y[0]])   # P, -P
    scalars[11] = scalars[10]
    scalars[12] = np.array(pymodel.to_limbs(pymodel.fr_to_mont(c, c.r - 1), 4), dtype=np.uint64)
    want = pymodel.msm(c, _decode_bases(c, bases), _decode_scalars(c, scalars))
    for strat in ("none", "parallel_term"):
        assert _decode_affine(c, o.msm_affine(bases, scalars, strategy=strat, threads=5)) == want
    # all-zero scalars -> identity
    z = np.zeros_like(scalars)
    assert _decode_affine(c, o.msm_affine(bases, z)) is pymodel.INF


def test_window_rule(oracles):
    # msm_ctx.h:22-48
    o = oracles["bn254"]
    assert o.window_bits(31) == 3 and o.window_bits(32) == 5
    assert o.window_bits(1 << 17) == 13 and o.window_count(13) == 20
    assert o.window_bits(1 << 24) == 18 and o.window_count(18) == 15
    assert oracles["bls12_381"].window_count(13) == 20  # ceil(255/13)
