"""File-level Groth16, CPU part: the Python model of the zkey / wtns formats is
pinned to the values the reference's own unit tests expect (vendors/circom/circomlib/zkey/
zkey_unittest.cc:56-174, wtns/wtns_unittest.cc:22-45), and the product's host-side parser +
NTT-based witness map (tachyon_<c>_groth16_witness_map_from_files_b200, no GPU needed) must
give the same h scalars as the model's long-division form."""
import os

import numpy as np

from oracle import circom_model, pymodel
from tachyon_b200 import msm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ZKEY, WTNS = os.path.join(GOLDEN, "multiplier_3.zkey"), os.path.join(GOLDEN, "multiplier_3.wtns")
R_BN254 = circom_model.FR["bn254"][0]


def test_model_matches_reference_unit_test_values():
    z = circom_model.parse_zkey(ZKEY)
    assert (z["num_vars"], z["num_public"], z["domain_size"]) == (6, 1, 4)        # zkey_unittest.cc:127-129
    assert z["r"] == R_BN254 and z["q"] == pymodel.CURVES["bn254"].p
    assert z["coefficients"] == [(0, 0, 2, R_BN254 - 1), (1, 0, 3, 1), (0, 1, 5, R_BN254 - 1), (1, 1, 4, 1),
                                 (0, 2, 0, 1), (0, 3, 1, 1)]                         # zkey_unittest.cc:159-166
    c = pymodel.CURVES["bn254"]
    Ri = pow(c.fq_R, -1, c.p)
    ic0 = [int.from_bytes(z["ic"][i * 32:(i + 1) * 32], "little") * Ri % c.p for i in range(2)]
    assert ic0 == [1400989341879513116647759947859271187117391672677487101192308885590924596480,
                   18827163924960691750679623127657074266908067481725903803154895122477837234033]  # :144-147
    r, w = circom_model.parse_wtns(WTNS)
    assert r == R_BN254 and w == [1, 60, 3, 4, 5, 12]                                # wtns_unittest.cc:37-39


def test_witness_map_matches_model():
    z = circom_model.parse_zkey(ZKEY)
    _, w = circom_model.parse_wtns(WTNS)
    want = circom_model.witness_map("bn254", z, w)
    h, dom, pub = msm.groth16_witness_map_from_files("bn254", ZKEY, WTNS)
    assert (dom, pub) == (4, 1)
    c = pymodel.CURVES["bn254"]
    got = [pymodel.from_limbs(v) * pow(c.fr_R, -1, c.r) % c.r for v in h]
    assert got == want


def test_witness_map_on_a_larger_synthetic_circuit(tmp_path):
    """A chain of 37 multiplications (domain 64) written in the same formats: exercises the
    radix-2 NTT beyond the 4-point fixture."""
    import struct
    r = R_BN254
    n_mul, n_pub = 37, 2
    vals = [1, 0, 0] + [pow(7, i + 1, 1000003) for i in range(n_mul + 1)]   # 1, out placeholders, x_0..x_n
    # constraints: x_i * x_(i+1) = y_i; y_i are further variables
    ys = [vals[3 + i] * vals[4 + i] % r for i in range(n_mul)]
    witness = vals + ys
    witness[1], witness[2] = ys[0], ys[-1]
    num_vars = len(witness)
    domain = 64
    R2 = pow(1 << 256, 2, r)
    coefs = []
    for i in range(n_mul):
        coefs.append((0, i, 3 + i, 1))
        coefs.append((1, i, 4 + i, 1))
    # public-input rows as snarkjs adds them: A = signal, B = 0
    for j in range(n_pub + 1):
        coefs.append((0, n_mul + j, j, 1))

    def sec(t, payload):
        return struct.pack("<IQ", t, len(payload)) + payload
    g1, g2 = bytes(64), bytes(128)
    q = pymodel.CURVES["bn254"].p
    hdr = struct.pack("<I", 32) + q.to_bytes(32, "little") + struct.pack("<I", 32) + r.to_bytes(32, "little") + \
        struct.pack("<III", num_vars, n_pub, domain) + g1 + g1 + g2 + g2 + g1 + g2
    cbytes = struct.pack("<I", len(coefs)) + b"".join(
        struct.pack("<III", m, c, s) + (v * R2 % r).to_bytes(32, "little") for m, c, s, v in coefs)
    zkey = b"zkey" + struct.pack("<II", 1, 9) + sec(1, struct.pack("<I", 1)) + sec(2, hdr) + \
        sec(3, g1 * (n_pub + 1)) + sec(4, cbytes) + sec(5, g1 * num_vars) + sec(6, g1 * num_vars) + \
        sec(7, g2 * num_vars) + sec(8, g1 * (num_vars - n_pub - 1)) + sec(9, g1 * domain)
    wtns = b"wtns" + struct.pack("<II", 2, 2) + sec(1, struct.pack("<I", 32) + r.to_bytes(32, "little") +
                                                   struct.pack("<I", num_vars)) + \
        sec(2, b"".join(v.to_bytes(32, "little") for v in witness))
    zp, wp = tmp_path / "chain.zkey", tmp_path / "chain.wtns"
    zp.write_bytes(zkey)
    wp.write_bytes(wtns)
    z = circom_model.parse_zkey(str(zp))
    _, w = circom_model.parse_wtns(str(wp))
    want = circom_model.witness_map("bn254", z, w)
    h, dom, pub = msm.groth16_witness_map_from_files("bn254", str(zp), str(wp))
    c = pymodel.CURVES["bn254"]
    got = [pymodel.from_limbs(v) * pow(c.fr_R, -1, c.r) % c.r for v in h]
    assert dom == 64 and pub == n_pub and got == want and any(got)


def test_bad_files_are_rejected(tmp_path):
    import pytest
    bad = tmp_path / "bad.zkey"
    bad.write_bytes(b"nope" + bytes(40))
    with pytest.raises(RuntimeError):
        msm.groth16_witness_map_from_files("bn254", str(bad), WTNS)
    with pytest.raises(RuntimeError):
        msm.groth16_witness_map_from_files("bls12_381", ZKEY, WTNS)     # a BN254 key on the other curve
    with pytest.raises(RuntimeError):
        msm.groth16_witness_map_from_files("bn254", str(tmp_path / "missing.zkey"), WTNS)
