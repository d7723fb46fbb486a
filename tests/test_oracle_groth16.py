"""Pins the oracle's restatement of tachyon/zk/r1cs/groth16/prove.h:33-165 (proof from
assignments) against the independent Python big-int model.  CPU only."""
import numpy as np
import pytest

from oracle import cpu_oracle, pymodel
from tests.groth16_util import make_case


def _g1(c, aff):
    aff = np.asarray(aff).reshape(2, -1)
    if not aff.any():
        return pymodel.INF
    Ri = pow(c.fq_R, -1, c.p)
    return tuple(pymodel.from_limbs(v) * Ri % c.p for v in aff)


def _g2(c, aff):
    aff = np.asarray(aff).reshape(4, -1)
    if not aff.any():
        return pymodel.INF
    Ri = pow(1 << (64 * c.fq_limbs), -1, c.p)
    v = [pymodel.from_limbs(x) * Ri % c.p for x in aff]
    return ((v[0], v[1]), (v[2], v[3]))


def _fr(c, limbs):
    return pymodel.from_limbs(limbs) * pow(c.fr_R, -1, c.r) % c.r


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
@pytest.mark.parametrize("blind", [True, False])
def test_groth16_proof_matches_python_model(oracles, curve, blind):
    c1, c2 = pymodel.CURVES[curve], pymodel.CURVES_G2[curve + "_g2"]
    pk, r, s, h, witness, full = make_case(oracles, curve, n_full=14, n_pub=3, h_size=9, seed=50, h_query_size=8,
                                           blind=blind)
    a, b, c = cpu_oracle.groth16_prove(curve, pk, r, s, h, witness, full)
    R, S = _fr(c1, r), _fr(c1, s)
    xs = [_fr(c1, v) for v in full]
    ws = [_fr(c1, v) for v in witness]
    hs = [_fr(c1, v) for v in h][:8]                       # h one longer than the query: last dropped
    add1, mul1 = (lambda p, q: pymodel.add(c1, p, q)), (lambda k, p: pymodel.mul(c1, k % c1.r, p))
    add2, mul2 = (lambda p, q: pymodel.g2_add(c2, p, q)), (lambda k, p: pymodel.g2_mul(c2, k % c2.r, p))
    aq = [_g1(c1, p) for p in pk["a_g1_query"]]
    b1q = [_g1(c1, p) for p in pk["b_g1_query"]]
    b2q = [_g2(c2, p) for p in pk["b_g2_query"]]
    delta1, delta2 = _g1(c1, pk["delta_g1"]), _g2(c2, pk["delta_g2"])
    A = add1(add1(add1(mul1(R, delta1), aq[0]), pymodel.msm(c1, aq[1:], xs)), _g1(c1, pk["alpha_g1"]))
    B2 = add2(add2(add2(mul2(S, delta2), b2q[0]), pymodel.g2_msm(c2, b2q[1:], xs)), _g2(c2, pk["beta_g2"]))
    C = mul1(S, A)
    if R:
        B1 = add1(add1(add1(mul1(S, delta1), b1q[0]), pymodel.msm(c1, b1q[1:], xs)), _g1(c1, pk["beta_g1"]))
        C = add1(C, mul1(R, B1))
        C = add1(C, pymodel.neg(c1, mul1(S * R, delta1)))
    C = add1(C, pymodel.msm(c1, [_g1(c1, p) for p in pk["l_g1_query"]], ws))
    C = add1(C, pymodel.msm(c1, [_g1(c1, p) for p in pk["h_g1_query"]], hs))
    assert _g1(c1, a) == A
    assert _g2(c2, b) == B2
    assert _g1(c1, c) == C
