"""Host-side mirror of the reference's MSM-GPU interface over the C ABI.

`MSMGpu` corresponds to the object behind `tachyon_<curve>_g1_msm_gpu_ptr`
(tachyon/c/math/elliptic_curves/msm/msm_gpu.h:22-122): create(degree) ->
affine_msm / point2_msm(bases, scalars) -> Jacobian point -> destroy.  Inputs are
numpy uint64 arrays with the byte layout of the C structs (little-endian limbs,
Montgomery form), or raw device pointers.
"""
import ctypes

import numpy as np

from . import _lib

DIST = {"uniform": 0, "non_uniform": 1, "witness": 2}
FIELD_OPS = {"add": 0, "sub": 1, "mul": 2, "square": 3, "neg": 4, "double": 5, "inverse": 6,
             "from_mont": 7, "to_mont": 8}
POINT_OPS = {"add": 0, "madd": 1, "msub": 2, "double": 3}


def _ptr(x):
    """numpy array -> host pointer; int -> raw (device) pointer."""
    if isinstance(x, (int, np.integer)):
        return ctypes.c_void_p(int(x))
    assert x.dtype == np.uint64 and x.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(x.ctypes.data)


class MSMGpu:
    def __init__(self, curve="bn254", degree=20, device=None, banner=False):
        """curve: "bn254", "bls12_381" (G1) or "bn254_g2", "bls12_381_g2" (G2)."""
        self.name = curve
        curve, group = _lib.split_name(curve)
        self.curve, self.group = curve, group
        self.fq_limbs = _lib.element_limbs(self.name)   # u64 limbs of one point coordinate
        self.L = _lib.load()
        self._f = lambda name: getattr(self.L, f"tachyon_{curve}_" + name.replace("g1_", group + "_", 1))
        self._f("g1_init")()
        if banner:
            self.ptr = self._f("g1_create_msm_gpu")(degree)   # reference entry point
        else:
            self.ptr = self._f("g1_create_msm_gpu_b200")(degree, 0 if device is None else device)
        if not self.ptr:
            raise RuntimeError("create_msm_gpu failed: " + _lib.last_error())

    def close(self):
        if getattr(self, "ptr", None):
            self._f("g1_destroy_msm_gpu")(self.ptr)
            self.ptr = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name, value):
        _lib.check(self._f("g1_msm_gpu_set_option_b200")(self.ptr, name.encode(), int(value)), "set_option")

    def set_stream(self, cuda_stream):
        _lib.check(self._f("g1_msm_gpu_set_stream_b200")(self.ptr, ctypes.c_void_p(cuda_stream)), "set_stream")

    def join_ranks(self, unique_id, rank, world):
        """Point-range sharding over `world` processes (one GPU each): afterwards every MSM call
        returns the sum over all ranks (one ncclAllGather of the partials on the context's
        stream).  unique_id: the 128 bytes of nccl_unique_id(), the same on every rank."""
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id))
        _lib.check(self._f("g1_msm_gpu_join_ranks_b200")(self.ptr, ctypes.cast(buf, ctypes.c_void_p), rank, world),
                   "join_ranks")

    def _jacobian(self, fn, bases, scalars, size):
        """Calls the reference-shaped entry point; returns (3, fq_limbs) and frees the result."""
        p = self._f(fn)(self.ptr, _ptr(bases), _ptr(scalars), size)
        if not p:
            raise RuntimeError(fn + " returned NULL")
        n = 3 * self.fq_limbs
        out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(n,)).copy()
        _free_cxx(p)
        return out.reshape(3, self.fq_limbs)

    def affine_msm(self, bases, scalars, size=None):
        size = len(scalars) if size is None else size
        return self._jacobian("g1_affine_msm_gpu", bases, scalars, size)

    def point2_msm(self, bases, scalars, size=None):
        size = len(scalars) if size is None else size
        return self._jacobian("g1_point2_msm_gpu", bases, scalars, size)

    def msm_xyzz(self, bases, scalars, size=None):
        """Un-normalised XYZZ sum (4, fq_limbs); error codes instead of abort."""
        size = len(scalars) if size is None else size
        out = np.zeros((4, self.fq_limbs), dtype=np.uint64)
        _lib.check(self._f("g1_msm_gpu_xyzz_b200")(self.ptr, _ptr(bases), _ptr(scalars), size, _ptr(out)),
                   "msm_gpu_xyzz")
        return out

    def register_bases(self, bases, size=None):
        """Upload (or copy, for a device pointer) the bases once; later commit_batch calls use them."""
        size = len(bases) if size is None else size
        self._registered_keepalive = None
        _lib.check(self._f("g1_msm_gpu_register_bases_b200")(self.ptr, _ptr(bases), size), "register_bases")

    def commit_batch(self, scalars_list, sizes=None):
        """One MSM per entry of scalars_list over the registered bases -> (count, 4, fq_limbs) XYZZ."""
        count = len(scalars_list)
        sizes = [len(s) for s in scalars_list] if sizes is None else list(sizes)
        ptrs = (ctypes.c_void_p * count)(*[_ptr(s) for s in scalars_list])
        szs = (ctypes.c_size_t * count)(*sizes)
        out = np.zeros((count, 4, self.fq_limbs), dtype=np.uint64)
        _lib.check(self._f("g1_msm_gpu_commit_batch_b200")(self.ptr, ptrs, szs, count, _ptr(out)), "commit_batch")
        return out

    def msm_batch(self, bases_list, scalars_list, sizes=None):
        """MSM i over bases_list[i] (None = registered bases) and scalars_list[i] -> (count, 4, fq_limbs)."""
        count = len(scalars_list)
        sizes = [len(s) for s in scalars_list] if sizes is None else list(sizes)
        bptr = (ctypes.c_void_p * count)(*[ctypes.c_void_p(0) if b is None else _ptr(b) for b in bases_list])
        sptr = (ctypes.c_void_p * count)(*[_ptr(s) for s in scalars_list])
        szs = (ctypes.c_size_t * count)(*sizes)
        out = np.zeros((count, 4, self.fq_limbs), dtype=np.uint64)
        _lib.check(self._f("g1_msm_gpu_batch_b200")(self.ptr, bptr, sptr, szs, count, _ptr(out)), "msm_batch")
        return out

    def last_timing(self):
        t = _lib.MsmTiming()
        _lib.check(self._f("g1_msm_gpu_last_timing_b200")(self.ptr, ctypes.byref(t)), "last_timing")
        return t.as_dict()


def _sym(name, fn):
    """C symbol of a per-group function for a curve name such as "bn254" or "bn254_g2"."""
    curve, group = _lib.split_name(name)
    return getattr(_lib.load(), f"tachyon_{curve}_{group}_{fn}")


_libstdcxx = None


def _free_cxx(p):
    """`delete` for the heap Jacobian the C API returns (POD, so operator delete(void*))."""
    global _libstdcxx
    if _libstdcxx is None:
        _libstdcxx = ctypes.CDLL("libstdc++.so.6")
        _libstdcxx._ZdlPv.argtypes = [ctypes.c_void_p]
        _libstdcxx._ZdlPv.restype = None
    _libstdcxx._ZdlPv(ctypes.c_void_p(p))


def generate_bases_device(curve, seed, n, device_ptr, first=0):
    _lib.check(_sym(curve, "generate_bases_b200")(seed, first, n, ctypes.c_void_p(device_ptr)), "generate_bases")


def generate_scalars_device(curve, seed, n, device_ptr, dist="uniform", first=0):
    _lib.check(_sym(curve, "generate_scalars_b200")(
        seed, DIST[dist], first, n, ctypes.c_void_p(device_ptr)), "generate_scalars")


def field_op(curve, field, op, a, b=None):
    """Element-wise device field op on host arrays (parity hook)."""
    L = _lib.load()
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    curve = _lib.split_name(curve)[0]
    _lib.check(getattr(L, f"tachyon_{curve}_{field}_op_b200")(FIELD_OPS[op], _ptr(a), _ptr(b), _ptr(out), a.shape[0]),
               f"{field}_op")
    return out


def point_op(curve, op, a_xyzz, b=None):
    L = _lib.load()
    a = np.ascontiguousarray(a_xyzz, dtype=np.uint64)
    out = np.empty_like(a)
    bp = _ptr(np.ascontiguousarray(b, dtype=np.uint64)) if b is not None else ctypes.c_void_p(0)
    _lib.check(_sym(curve, "point_op_b200")(POINT_OPS[op], _ptr(a), bp, _ptr(out), a.shape[0]), "point_op")
    return out


def imad_peak(device=0, variant=0, repeats=5):
    v = _lib.load().tachyon_b200_imad_peak(device, variant, repeats)
    if v < 0:
        raise RuntimeError("imad_peak failed: " + _lib.last_error())
    return v


def nccl_unique_id():
    """128-byte ncclUniqueId (call on rank 0, broadcast to the others)."""
    buf = (ctypes.c_char * 128)()
    _lib.check(_lib.load().tachyon_b200_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "nccl_unique_id")
    return bytes(buf)


def kernel_launch_count():
    return int(_lib.load().tachyon_b200_kernel_launch_count())


def device_count():
    return int(_lib.load().tachyon_b200_device_count())


def xyzz_add(curve, a, b):
    """Host-side a + b on XYZZ points (4, fq_limbs): combination of per-rank partial sums."""
    L = _lib.load()
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    _sym(curve, "xyzz_add_b200")(_ptr(a), _ptr(b), _ptr(out))
    return out


def xyzz_to_jacobian(curve, a):
    L = _lib.load()
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.zeros((3, a.shape[-1]), dtype=np.uint64)
    _sym(curve, "xyzz_to_jacobian_b200")(_ptr(a), _ptr(out))
    return out


def batch_normalize(curve, xyzz):
    """(n, 4, fq_limbs) XYZZ -> (n, 2 * fq_limbs) affine with one inversion (host)."""
    L = _lib.load()
    a = np.ascontiguousarray(xyzz, dtype=np.uint64)
    n = a.shape[0]
    out = np.zeros((n, 2 * a.shape[-1]), dtype=np.uint64)
    _sym(curve, "xyzz_batch_normalize_b200")(_ptr(a), n, _ptr(out))
    return out


def window_bits(n, scalar_bits):
    return int(_lib.load().tachyon_b200_window_bits(n, scalar_bits))


def window_count(scalar_bits, c):
    return int(_lib.load().tachyon_b200_window_count(scalar_bits, c))


def groth16_prove(g1_ctx, g2_ctx, pk, r, s, h, witness, full):
    """Groth16 proof assembly (tachyon_<c>_groth16_prove_b200; zk/r1cs/groth16/prove.h:33-165).

    pk: dict with affine points alpha_g1, beta_g1, delta_g1 (2 * fq limbs), beta_g2, delta_g2
    (4 * fq limbs) and query arrays a_g1_query, b_g1_query, b_g2_query, h_g1_query, l_g1_query
    (numpy arrays, or (device_pointer, count) tuples).  r, s: Montgomery Fr (4 limbs).
    Returns (a, b, c) affine points as uint64 arrays."""
    curve = g1_ctx.curve
    fq = _lib.CURVES[curve]
    parts = []
    for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
        parts.append(np.ascontiguousarray(pk[k], dtype=np.uint64).reshape(-1))
    keep = []
    tail = []
    for k in ("a_g1_query", "b_g1_query", "b_g2_query", "h_g1_query", "l_g1_query"):
        q = pk[k]
        if isinstance(q, tuple):
            ptr, cnt = int(q[0]), int(q[1])
        else:
            q = np.ascontiguousarray(q, dtype=np.uint64)
            keep.append(q)
            ptr, cnt = q.ctypes.data, q.shape[0]
        tail += [ptr, cnt]
    blob = np.concatenate(parts + [np.array(tail, dtype=np.uint64)])
    assert blob.size == 3 * 2 * fq + 2 * 4 * fq + 10
    out = np.zeros(2 * fq + 4 * fq + 2 * fq, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    rc = getattr(_lib.load(), f"tachyon_{curve}_groth16_prove_b200")(
        ctypes.c_void_p(g1_ctx.ptr), ctypes.c_void_p(g2_ctx.ptr), _ptr(blob), _ptr(r), _ptr(s),
        _ptr(h), len(h), _ptr(witness), len(witness), _ptr(full), len(full), _ptr(out))
    _lib.check(rc, "groth16_prove")
    return out[:2 * fq], out[2 * fq:6 * fq], out[6 * fq:]


def groth16_witness_map_from_files(curve, zkey_path, wtns_path):
    """Host-only: (h scalars (domain_size, 4) uint64 Montgomery, domain_size, num_public) of a
    .zkey + .wtns pair (tachyon_<c>_groth16_witness_map_from_files_b200)."""
    fn = getattr(_lib.load(), f"tachyon_{curve}_groth16_witness_map_from_files_b200")
    dom, pub = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(fn(zkey_path.encode(), wtns_path.encode(), None, 0, ctypes.byref(dom), ctypes.byref(pub)),
               "witness_map_from_files")
    h = np.zeros((dom.value, 4), dtype=np.uint64)
    _lib.check(fn(zkey_path.encode(), wtns_path.encode(), _ptr(h), dom.value, ctypes.byref(dom), ctypes.byref(pub)),
               "witness_map_from_files")
    return h, dom.value, pub.value


def groth16_prove_from_files(g1_ctx, g2_ctx, zkey_path, wtns_path, r=None, s=None, proof_json=None,
                             public_json=None):
    """Proof (a, b, c) from a snarkjs .zkey and a .wtns (tachyon_<c>_groth16_prove_from_files_b200;
    vendors/circom/prover_main.cc:81-186).  r, s: Montgomery Fr (4 limbs) or None (= --no_zk)."""
    curve = g1_ctx.curve
    fq = _lib.CURVES[curve]
    out = np.zeros(2 * fq + 4 * fq + 2 * fq, dtype=np.uint64)
    rp = _ptr(np.ascontiguousarray(r, dtype=np.uint64)) if r is not None else None
    sp = _ptr(np.ascontiguousarray(s, dtype=np.uint64)) if s is not None else None
    rc = getattr(_lib.load(), f"tachyon_{curve}_groth16_prove_from_files_b200")(
        ctypes.c_void_p(g1_ctx.ptr), ctypes.c_void_p(g2_ctx.ptr), zkey_path.encode(), wtns_path.encode(), rp, sp,
        _ptr(out), proof_json.encode() if proof_json else None, public_json.encode() if public_json else None)
    _lib.check(rc, "groth16_prove_from_files")
    return out[:2 * fq], out[2 * fq:6 * fq], out[6 * fq:]


class HostBuffer:
    """Page-locked host memory from tachyon_b200_alloc_host, viewed as a (rows, limbs) uint64
    array.  write_combined pages are for inputs the CPU only writes: several GPUs copying at
    once read them faster than ordinary pinned memory."""

    def __init__(self, rows, limbs, write_combined=False):
        self.nbytes = rows * limbs * 8
        self.ptr = _lib.load().tachyon_b200_alloc_host(self.nbytes, 1 if write_combined else 0)
        if not self.ptr:
            raise RuntimeError("alloc_host failed: " + _lib.last_error())
        raw = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=np.uint64, count=rows * limbs).reshape(rows, limbs)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            _lib.load().tachyon_b200_free_host(ctypes.c_void_p(self.ptr))
            self.ptr = None

    __del__ = free
