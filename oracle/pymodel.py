"""ORACLE — TEST INFRASTRUCTURE ONLY (see tachyon_cpu_msm.cc header).

Independent pure-Python big-int model of the group law, used by tests/ to
cross-check the C++ oracle at small sizes.  It shares no code with the C++
restatement: affine chord-and-tangent arithmetic over Python ints, canonical
(non-Montgomery) values.  Curve constants are the ones pinned by the
reference's BUILD files:
  bn254:      tachyon/math/elliptic_curves/bn/bn254/BUILD.bazel:27-60,119-148
  bls12_381:  tachyon/math/elliptic_curves/bls12/bls12_381/BUILD.bazel:36-71,117-151
  gf7:        tachyon/math/elliptic_curves/short_weierstrass/test/sw_curve_config.h:31-45
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Curve:
    name: str
    p: int          # base field modulus
    r: int          # scalar field modulus
    b: int          # y^2 = x^3 + b
    gx: int
    gy: int
    fq_limbs: int   # u64 limbs of Fq
    fr_limbs: int

    @property
    def fq_R(self):
        return 1 << (64 * self.fq_limbs)

    @property
    def fr_R(self):
        return 1 << (64 * self.fr_limbs)


BN254 = Curve(
    "bn254",
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    3, 1, 2, 4, 4)

BLS12_381 = Curve(
    "bls12_381",
    4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787,
    52435875175126190479447740508185965837690552500527637822603658699938581184513,
    4,
    3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
    1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569,
    6, 4)

GF7 = Curve("gf7", 7, 7, 5, 5, 5, 1, 1)

CURVES = {"bn254": BN254, "bls12_381": BLS12_381, "gf7": GF7}

INF = None  # point at infinity


def is_on_curve(c, pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - c.b) % c.p == 0


def neg(c, pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % c.p)


def add(c, p1, p2):
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % c.p == 0:
            return INF
        lam = (3 * x1 * x1) * pow(2 * y1, -1, c.p) % c.p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, c.p) % c.p
    x3 = (lam * lam - x1 - x2) % c.p
    y3 = (lam * (x1 - x3) - y1) % c.p
    return (x3, y3)


def mul(c, k, pt):
    acc = INF
    while k > 0:
        if k & 1:
            acc = add(c, acc, pt)
        pt = add(c, pt, pt)
        k >>= 1
    return acc


def msm(c, points, scalars):
    """sum_i scalars[i] * points[i]; points are canonical (x, y) or INF."""
    acc = INF
    for pt, k in zip(points, scalars):
        acc = add(c, acc, mul(c, k % c.r, pt))
    return acc


# --- byte-level helpers (little-endian u64 limbs, Montgomery form) ----------
def to_limbs(x, n):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(limbs):
    return sum(int(v) << (64 * i) for i, v in enumerate(limbs))


def fq_to_mont(c, x):
    return x * c.fq_R % c.p


def fq_from_mont(c, x):
    return x * pow(c.fq_R, -1, c.p) % c.p


def fr_to_mont(c, x):
    return x * c.fr_R % c.r


def fr_from_mont(c, x):
    return x * pow(c.fr_R, -1, c.r) % c.r


def fill_digits(k, c_bits, num_digits):
    """Independent statement of signed-window recoding: digits d_i in
    [-2^(c-1), 2^(c-1)) with sum d_i 2^(c i) == k; the top digit absorbs the
    last carry (semantics of tachyon pippenger.h:27-51)."""
    digits = []
    carry = 0
    for i in range(num_digits):
        coeff = ((k >> (c_bits * i)) & ((1 << c_bits) - 1)) + carry
        carry = 1 if coeff >= (1 << (c_bits - 1)) else 0
        digits.append(coeff - (carry << c_bits))
    digits[-1] += carry << c_bits
    return digits


# --- G2: the same curves over Fq2 = Fq[u] / (u^2 + 1) --------------------------------------
# Constants from bn/bn254/BUILD.bazel:150-200 (b = 3 / (9 + u)) and
# bls12/bls12_381/BUILD.bazel:153-200 (b = 4 + 4u).  Elements are (c0, c1) tuples of ints.
@dataclass(frozen=True)
class CurveG2:
    name: str
    p: int
    r: int
    b: tuple
    gx: tuple
    gy: tuple
    fq_limbs: int   # u64 limbs of ONE Fq component
    fr_limbs: int


BN254_G2 = CurveG2(
    "bn254_g2", BN254.p, BN254.r,
    (19485874751759354771024239261021720505790618469301721065564631296452457478373,
     266929791119991161246907387137283842545076965332900288569378510910307636690),
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
    4, 4)

BLS12_381_G2 = CurveG2(
    "bls12_381_g2", BLS12_381.p, BLS12_381.r, (4, 4),
    (352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
     3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
    (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
     927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582),
    6, 4)

CURVES_G2 = {"bn254_g2": BN254_G2, "bls12_381_g2": BLS12_381_G2}


def f2_add(p, a, b):
    return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)


def f2_sub(p, a, b):
    return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)


def f2_mul(p, a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f2_inv(p, a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return (a[0] * n % p, (-a[1] * n) % p)


def g2_is_on_curve(c, pt):
    if pt is INF:
        return True
    x, y = pt
    return f2_sub(c.p, f2_mul(c.p, y, y), f2_add(c.p, f2_mul(c.p, f2_mul(c.p, x, x), x), c.b)) == (0, 0)


def g2_add(c, p1, p2):
    p = c.p
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if f2_add(p, y1, y2) == (0, 0):
            return INF
        x1sq = f2_mul(p, x1, x1)
        lam = f2_mul(p, f2_add(p, f2_add(p, x1sq, x1sq), x1sq), f2_inv(p, f2_add(p, y1, y1)))
    else:
        lam = f2_mul(p, f2_sub(p, y2, y1), f2_inv(p, f2_sub(p, x2, x1)))
    x3 = f2_sub(p, f2_sub(p, f2_mul(p, lam, lam), x1), x2)
    y3 = f2_sub(p, f2_mul(p, lam, f2_sub(p, x1, x3)), y1)
    return (x3, y3)


def g2_mul(c, k, pt):
    acc = INF
    while k > 0:
        if k & 1:
            acc = g2_add(c, acc, pt)
        pt = g2_add(c, pt, pt)
        k >>= 1
    return acc


def g2_msm(c, points, scalars):
    acc = INF
    for pt, k in zip(points, scalars):
        acc = g2_add(c, acc, g2_mul(c, k % c.r, pt))
    return acc
