"""TEST INFRASTRUCTURE (not used by the product): independent pure-Python model of the circom /
snarkjs file formats and of the QAP witness map, for the file-level Groth16 tests.

Restates, with Python integers and schoolbook polynomial arithmetic,
  vendors/circom/circomlib/zkey/zkey.h:88-317        zkey v1 sections
  vendors/circom/circomlib/wtns/wtns.h:66-154        wtns v2 sections
  vendors/circom/circomlib/base/sections.h:26-60     the section table
  vendors/circom/circomlib/circuit/quadratic_arithmetic_program.h:25-118  WitnessMapFromMatrices
The witness map is NOT computed the way the product does it (radix-2 NTTs): here a, b, c are
interpolated, (a b - c) is divided by the vanishing polynomial x^n - 1 by long division (the
remainder must be zero) and the quotient is evaluated on the coset of the 2n-th root of unity,
where x^n - 1 = -2 — so h_i = (a b - c)(g w^i) = -2 q(g w^i).
"""
import struct

FR = {
    "bn254": (21888242871839275222246405745257275088548364400416034343698204186575808495617, 5),
    "bls12_381": (52435875175126190479447740508185965837690552500527637822603658699938581184513, 7),
}


def sections(buf, magic, version):
    assert buf[:4] == magic, "bad magic"
    ver, nsec = struct.unpack_from("<II", buf, 4)
    assert ver == version, "bad version"
    off, out = 12, {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", buf, off)
        off += 12
        out.setdefault(typ, buf[off:off + size])
        off += size
    return out


def _modulus(sec, off):
    n8, = struct.unpack_from("<I", sec, off)
    off += 4
    return int.from_bytes(sec[off:off + n8], "little"), n8, off + n8


def parse_zkey(path):
    """dict with q, r, num_vars, num_public, domain_size, header and query points as raw bytes
    (Montgomery, as stored) and the coefficient list [(matrix, constraint, signal, value)] with
    canonical integer values."""
    buf = open(path, "rb").read()
    sec = sections(buf, b"zkey", 1)
    assert struct.unpack_from("<I", sec[1], 0)[0] == 1, "not groth16"
    h = sec[2]
    q, n8q, off = _modulus(h, 0)
    r, n8r, off = _modulus(h, off)
    num_vars, num_public, domain_size = struct.unpack_from("<III", h, off)
    off += 12
    g1, g2 = 2 * n8q, 4 * n8q
    z = {"q": q, "r": r, "n8q": n8q, "n8r": n8r, "num_vars": num_vars, "num_public": num_public,
         "domain_size": domain_size}
    for name, size in (("alpha_g1", g1), ("beta_g1", g1), ("beta_g2", g2), ("gamma_g2", g2), ("delta_g1", g1),
                       ("delta_g2", g2)):
        z[name] = h[off:off + size]
        off += size
    z["ic"] = sec[3][:(num_public + 1) * g1]
    ncoef, = struct.unpack_from("<I", sec[4], 0)
    R = 1 << (8 * n8r)
    Rinv2 = pow(R, -2, r)
    coefs, off = [], 4
    for _ in range(ncoef):
        m, c, s = struct.unpack_from("<III", sec[4], off)
        v = int.from_bytes(sec[4][off + 12:off + 12 + n8r], "little")
        coefs.append((m, c, s, v * Rinv2 % r))      # stored as value * R^2 (zkey.h:216-219 strips one R)
        off += 12 + n8r
    z["coefficients"] = coefs
    z["a_g1"] = sec[5][:num_vars * g1]
    z["b_g1"] = sec[6][:num_vars * g1]
    z["b_g2"] = sec[7][:num_vars * g2]
    z["c_g1"] = sec[8][:(num_vars - num_public - 1) * g1]
    z["h_g1"] = sec[9][:domain_size * g1]
    return z


def parse_wtns(path):
    """(modulus, [canonical integers])"""
    buf = open(path, "rb").read()
    sec = sections(buf, b"wtns", 2)
    r, n8, off = _modulus(sec[1], 0)
    n, = struct.unpack_from("<I", sec[1], off)
    return r, [int.from_bytes(sec[2][i * n8:(i + 1) * n8], "little") for i in range(n)]


def _interpolate(evals, w, r):
    """coefficients of the polynomial with p(w^i) = evals[i] (naive inverse DFT)."""
    n = len(evals)
    ninv, winv = pow(n, -1, r), pow(w, -1, r)
    return [ninv * sum(evals[i] * pow(winv, i * k, r) for i in range(n)) % r for k in range(n)]


def witness_map(curve, zkey, witness):
    """h scalars (canonical integers), see the module docstring."""
    r, gen = FR[curve]
    assert zkey["r"] == r
    n = zkey["domain_size"]
    a, b = [0] * n, [0] * n
    for m, c, s, v in zkey["coefficients"]:
        ab = a if m == 0 else b
        ab[c] = (ab[c] + v * witness[s]) % r
    cvals = [x * y % r for x, y in zip(a, b)]
    w = pow(gen, (r - 1) // n, r)
    g = pow(gen, (r - 1) // (2 * n), r)
    pa, pb, pc = (_interpolate(e, w, r) for e in (a, b, cvals))
    prod = [0] * (2 * n - 1)
    for i, x in enumerate(pa):
        for j, y in enumerate(pb):
            prod[i + j] = (prod[i + j] + x * y) % r
    for i, x in enumerate(pc):
        prod[i] = (prod[i] - x) % r
    # long division by x^n - 1
    quot = [0] * max(n - 1, 1)
    rem = prod[:]
    for k in range(len(rem) - 1, n - 1, -1):
        quot[k - n] = rem[k]
        rem[k - n] = (rem[k - n] + rem[k]) % r
        rem[k] = 0
    assert not any(rem), "a * b - c is not divisible by the vanishing polynomial"
    out = []
    for i in range(n):
        x = g * pow(w, i, r) % r
        q = 0
        for coeff in reversed(quot):
            q = (q * x + coeff) % r
        out.append(-2 * q % r)
    return out
