"""ORACLE — TEST INFRASTRUCTURE ONLY (see tachyon_cpu_msm.cc header).

ctypes front-end of liboracle_tachyon_msm.so for tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Arrays are numpy uint64,
shape (..., limbs), little-endian limbs, Montgomery form (the byte layout of
the tachyon_<curve>_fr / _g1_affine C structs).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_tachyon_msm.so")

_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the oracle with the committed Makefile."""
    src = os.path.join(_HERE, "tachyon_cpu_msm.cc")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_u64p)


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


class CurveOracle:
    """One instance per curve ("bn254" or "bls12_381")."""

    def __init__(self, name):
        self.name = name
        self.L = lib()
        self.fq_limbs = getattr(self.L, f"oracle_{name}_fq_limbs")()
        self.fr_limbs = getattr(self.L, f"oracle_{name}_fr_limbs")()
        getattr(self.L, f"oracle_{name}_window_bits").restype = ctypes.c_uint
        getattr(self.L, f"oracle_{name}_window_bits").argtypes = [ctypes.c_size_t]
        getattr(self.L, f"oracle_{name}_window_count").restype = ctypes.c_uint
        getattr(self.L, f"oracle_{name}_window_count").argtypes = [ctypes.c_uint]

    def _f(self, fn):
        return getattr(self.L, f"oracle_{self.name}_{fn}")

    # -- constants -----------------------------------------------------------
    def constants(self):
        q, r = self.fq_limbs, self.fr_limbs
        out = {k: np.zeros(n, dtype=np.uint64) for k, n in [
            ("fq_mod", q), ("fq_r", q), ("fq_r2", q), ("fq_inv", 1),
            ("fr_mod", r), ("fr_r", r), ("fr_r2", r), ("fr_inv", 1), ("gen", 2 * q)]}
        self._f("constants")(*[_p(out[k]) for k in (
            "fq_mod", "fq_r", "fq_r2", "fq_inv", "fr_mod", "fr_r", "fr_r2", "fr_inv", "gen")])
        return out

    # -- field ---------------------------------------------------------------
    OPS = {"add": 0, "sub": 1, "mul": 2, "square": 3, "neg": 4, "double": 5, "inverse": 6}

    def fq_op(self, op, a, b=None):
        a = _arr(a).reshape(-1, self.fq_limbs)
        b = a if b is None else _arr(b).reshape(-1, self.fq_limbs)
        out = np.empty_like(a)
        self._f("fq_op")(ctypes.c_int(self.OPS[op]), _p(a), _p(b), _p(out),
                         ctypes.c_size_t(a.shape[0]))
        return out

    def _conv(self, fn, a, limbs):
        a = _arr(a).reshape(-1, limbs)
        out = np.empty_like(a)
        self._f(fn)(_p(a), _p(out), ctypes.c_size_t(a.shape[0]))
        return out

    def fq_to_mont(self, a):
        return self._conv("fq_to_mont", a, self.fq_limbs)

    def fq_from_mont(self, a):
        return self._conv("fq_from_mont", a, self.fq_limbs)

    def fr_to_mont(self, a):
        return self._conv("fr_to_mont", a, self.fr_limbs)

    def fr_from_mont(self, a):
        return self._conv("fr_from_mont", a, self.fr_limbs)

    def fill_digits(self, scalar_mont, window_bits, num_digits):
        s = _arr(scalar_mont).reshape(self.fr_limbs)
        out = np.zeros(num_digits, dtype=np.int64)
        self._f("fill_digits")(_p(s), ctypes.c_size_t(window_bits), ctypes.c_size_t(num_digits),
                               out.ctypes.data_as(_i64p))
        return out

    # -- points --------------------------------------------------------------
    def _pt(self, fn, coords_out, *ins):
        ins = [_arr(x).reshape(-1) for x in ins]
        out = np.zeros(coords_out * self.fq_limbs, dtype=np.uint64)
        self._f(fn)(*[_p(x) for x in ins], _p(out))
        return out.reshape(coords_out, self.fq_limbs)

    def xyzz_add(self, a, b):
        return self._pt("xyzz_add", 4, a, b)

    def xyzz_madd(self, a, b_affine):
        return self._pt("xyzz_madd", 4, a, b_affine)

    def xyzz_double(self, a):
        return self._pt("xyzz_double", 4, a)

    def xyzz_to_affine(self, a):
        return self._pt("xyzz_to_affine", 2, a)

    def xyzz_to_jacobian(self, a):
        return self._pt("xyzz_to_jacobian", 3, a)

    def jacobian_to_affine(self, a):
        return self._pt("jacobian_to_affine", 2, a)

    def scalar_mul(self, p_affine, k_canonical):
        return self._pt("scalar_mul", 4, p_affine, _arr(k_canonical).reshape(self.fr_limbs))

    def xyzz_zero(self):
        one = self.constants()["fq_r"]
        z = np.zeros((4, self.fq_limbs), dtype=np.uint64)
        z[0] = one
        z[1] = one
        return z

    # -- MSM -----------------------------------------------------------------
    def window_bits(self, n):
        return self._f("window_bits")(n)

    def window_count(self, c):
        return self._f("window_count")(c)

    STRATEGY = {"none": 0, "parallel_window": 1, "parallel_term": 2}

    def msm(self, bases, scalars, strategy="parallel_term", threads=None):
        """Tachyon VariableBaseMSM::Run restated; returns XYZZ (4, fq_limbs)."""
        bases = _arr(bases).reshape(-1, 2 * self.fq_limbs)
        scalars = _arr(scalars).reshape(-1, self.fr_limbs)
        assert bases.shape[0] == scalars.shape[0]
        if threads is None:
            threads = self.L.oracle_max_threads()
        out = np.zeros(4 * self.fq_limbs, dtype=np.uint64)
        self._f("msm")(_p(bases), _p(scalars), ctypes.c_size_t(bases.shape[0]),
                       ctypes.c_int(self.STRATEGY[strategy]), ctypes.c_int(threads), _p(out))
        return out.reshape(4, self.fq_limbs)

    def msm_affine(self, bases, scalars, **kw):
        """Normalised affine result (2, fq_limbs) — the parity comparator."""
        return self.xyzz_to_affine(self.msm(bases, scalars, **kw))

    def msm_naive(self, bases, scalars):
        bases = _arr(bases).reshape(-1, 2 * self.fq_limbs)
        scalars = _arr(scalars).reshape(-1, self.fr_limbs)
        out = np.zeros(4 * self.fq_limbs, dtype=np.uint64)
        self._f("msm_naive")(_p(bases), _p(scalars), ctypes.c_size_t(bases.shape[0]), _p(out))
        return out.reshape(4, self.fq_limbs)

    # -- synthetic inputs --------------------------------------
    DIST = {"uniform": 0, "non_uniform": 1, "witness": 2}

    def generate_points(self, seed, n, first=0):
        out = np.zeros((n, 2 * self.fq_limbs), dtype=np.uint64)
        if n:
            self._f("generate_points")(ctypes.c_uint64(seed), ctypes.c_size_t(first),
                                       ctypes.c_size_t(n), _p(out))
        return out

    def generate_scalars(self, seed, n, dist="uniform", first=0):
        out = np.zeros((n, self.fr_limbs), dtype=np.uint64)
        if n:
            self._f("generate_scalars")(ctypes.c_uint64(seed), ctypes.c_int(self.DIST[dist]),
                                        ctypes.c_size_t(first), ctypes.c_size_t(n), _p(out))
        return out

    def fold_chain_scalars(self, scalars):
        scalars = _arr(scalars).reshape(-1, self.fr_limbs)
        n = scalars.shape[0]
        chains = (n + 4095) // 4096
        out = np.zeros((chains, self.fr_limbs), dtype=np.uint64)
        self._f("fold_chain_scalars")(_p(scalars), ctypes.c_size_t(n), _p(out))
        return out


def max_threads():
    return lib().oracle_max_threads()


# GF(7) toy curve (reference KATs) -------------------------------------------
def gf7(fn, n_out, *ins):
    L = lib()
    ins = [np.ascontiguousarray(x, dtype=np.uint64) for x in ins]
    out = np.zeros(n_out, dtype=np.uint64)
    getattr(L, f"oracle_gf7_{fn}")(*[_p(x) for x in ins], _p(out))
    return [int(v) for v in out]


def gf7_scalar_mul(p_affine, k):
    L = lib()
    p = np.ascontiguousarray(p_affine, dtype=np.uint64)
    out = np.zeros(2, dtype=np.uint64)
    L.oracle_gf7_scalar_mul(_p(p), ctypes.c_uint64(k), _p(out))
    return [int(v) for v in out]


def groth16_prove(curve, pk, r, s, h, witness, full):
    """tachyon/zk/r1cs/groth16/prove.h:54-165 CreateProofWithAssignment restated on the CPU MSMs.
    pk: dict of numpy arrays (alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2, a_g1_query,
    b_g1_query, b_g2_query, h_g1_query, l_g1_query).  Returns (a, b, c) affine, Montgomery limbs."""
    L = lib()
    fq = getattr(L, f"oracle_{curve}_fq_limbs")()
    pts = np.concatenate([_arr(pk[k]).reshape(-1) for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2")])
    q = {k: _arr(pk[k]) for k in ("a_g1_query", "b_g1_query", "b_g2_query", "h_g1_query", "l_g1_query")}
    out = np.zeros(8 * fq, dtype=np.uint64)
    sz = ctypes.c_size_t
    r, s, h, witness, full = (_arr(x) for x in (r, s, h, witness, full))
    getattr(L, f"oracle_{curve}_groth16_prove")(
        _p(pts), _p(q["a_g1_query"]), sz(len(q["a_g1_query"])), _p(q["b_g1_query"]), sz(len(q["b_g1_query"])),
        _p(q["b_g2_query"]), sz(len(q["b_g2_query"])), _p(q["h_g1_query"]), sz(len(q["h_g1_query"])),
        _p(q["l_g1_query"]), sz(len(q["l_g1_query"])), _p(r), _p(s), _p(h), sz(len(h)),
        _p(witness), sz(len(witness)), _p(full), sz(len(full)), _p(out))
    return out[:2 * fq], out[2 * fq:6 * fq], out[6 * fq:]
