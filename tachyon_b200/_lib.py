"""ctypes binding of libtachyon_msm_b200.so (the C ABI of include/tachyon_msm_b200.h).

There is no fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TACHYON_B200_LIB: load another build of the library (A/B comparisons on one GPU box)
LIB_PATH = os.environ.get("TACHYON_B200_LIB") or os.path.join(HERE, "lib", "libtachyon_msm_b200.so")

CURVES = {"bn254": 4, "bls12_381": 6}  # curve -> Fq u64 limbs (Fr is 4 for both)
GROUPS = ("g1", "g2")                  # g2: coordinates in Fq2, twice the limbs


def split_name(name):
    """"bn254" / "bn254_g2" -> (curve, group)."""
    if name.endswith("_g2"):
        return name[:-3], "g2"
    if name.endswith("_g1"):
        return name[:-3], "g1"
    return name, "g1"


def element_limbs(name):
    curve, group = split_name(name)
    return CURVES[curve] * (2 if group == "g2" else 1)


class MsmTiming(ctypes.Structure):
    _fields_ = [
        ("h2d_ms", ctypes.c_float), ("sort_ms", ctypes.c_float), ("accumulate_ms", ctypes.c_float),
        ("reduce_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("host_ms", ctypes.c_float),
        ("window_bits", ctypes.c_uint32), ("windows", ctypes.c_uint32), ("tasks", ctypes.c_uint32),
        ("entries", ctypes.c_uint32), ("kernel_launches", ctypes.c_uint32), ("devices", ctypes.c_uint32),
        ("ranges", ctypes.c_uint32), ("enqueue_ms", ctypes.c_float), ("wait_ms", ctypes.c_float),
        ("pair_rounds", ctypes.c_uint32), ("acc_kernel_ms", ctypes.c_float),
        ("acc_kernel_entries", ctypes.c_uint32), ("low_windows", ctypes.c_uint32),
        ("combine_ms", ctypes.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/tachyon_msm_b200.h declares: per curve and group, per curve, global
GROUP_SYMBOLS = [
    "tachyon_{c}_{g}_init", "tachyon_{c}_{g}_create_msm_gpu", "tachyon_{c}_{g}_destroy_msm_gpu",
    "tachyon_{c}_{g}_point2_msm_gpu", "tachyon_{c}_{g}_affine_msm_gpu",
    "tachyon_{c}_{g}_create_msm_gpu_b200", "tachyon_{c}_{g}_msm_gpu_set_stream_b200",
    "tachyon_{c}_{g}_msm_gpu_set_option_b200", "tachyon_{c}_{g}_msm_gpu_xyzz_b200",
    "tachyon_{c}_{g}_msm_gpu_last_timing_b200", "tachyon_{c}_{g}_generate_bases_b200",
    "tachyon_{c}_{g}_generate_scalars_b200", "tachyon_{c}_{g}_point_op_b200",
    "tachyon_{c}_{g}_xyzz_add_b200", "tachyon_{c}_{g}_xyzz_to_jacobian_b200",
    "tachyon_{c}_{g}_msm_gpu_register_bases_b200", "tachyon_{c}_{g}_msm_gpu_commit_batch_b200",
    "tachyon_{c}_{g}_xyzz_batch_normalize_b200", "tachyon_{c}_{g}_msm_gpu_batch_b200",
    "tachyon_{c}_{g}_msm_gpu_join_ranks_b200",
]
FIELD_SYMBOLS = ["tachyon_{c}_fq_op_b200", "tachyon_{c}_fr_op_b200", "tachyon_{c}_fq2_op_b200"]
CURVE_SYMBOLS = ["tachyon_{c}_groth16_prove_b200", "tachyon_{c}_groth16_prove_from_files_b200",
                 "tachyon_{c}_groth16_witness_map_from_files_b200"]
GLOBAL_SYMBOLS = ["tachyon_b200_device_count", "tachyon_b200_last_error", "tachyon_b200_imad_peak",
                  "tachyon_b200_kernel_launch_count", "tachyon_b200_window_bits",
                  "tachyon_b200_window_count", "tachyon_b200_nccl_unique_id", "tachyon_b200_alloc_host",
                  "tachyon_b200_free_host"]


def all_symbols():
    return ([s.format(c=c, g=g) for c in CURVES for g in GROUPS for s in GROUP_SYMBOLS] +
            [s.format(c=c) for c in CURVES for s in FIELD_SYMBOLS + CURVE_SYMBOLS] + GLOBAL_SYMBOLS)


_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built (python -m tachyon_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m tachyon_b200.build` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    for c in CURVES:
        for n in FIELD_SYMBOLS:
            getattr(lib, n.format(c=c)).argtypes = [i32, vp, vp, vp, sz]
        getattr(lib, f"tachyon_{c}_groth16_prove_b200").argtypes = [vp, vp, vp, vp, vp, vp, sz, vp, sz, vp, sz, vp]
        getattr(lib, f"tachyon_{c}_groth16_prove_from_files_b200").argtypes = [
            vp, vp, ctypes.c_char_p, ctypes.c_char_p, vp, vp, vp, ctypes.c_char_p, ctypes.c_char_p]
        getattr(lib, f"tachyon_{c}_groth16_witness_map_from_files_b200").argtypes = [
            ctypes.c_char_p, ctypes.c_char_p, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    for c, g in [(c, g) for c in CURVES for g in GROUPS]:
        f = lambda name: getattr(lib, name.format(c=c).replace("_g1_", "_%s_" % g))
        f("tachyon_{c}_g1_init").restype = None
        f("tachyon_{c}_g1_create_msm_gpu").restype = vp
        f("tachyon_{c}_g1_create_msm_gpu").argtypes = [ctypes.c_uint8]
        f("tachyon_{c}_g1_create_msm_gpu_b200").restype = vp
        f("tachyon_{c}_g1_create_msm_gpu_b200").argtypes = [ctypes.c_uint8, i32]
        f("tachyon_{c}_g1_destroy_msm_gpu").restype = None
        f("tachyon_{c}_g1_destroy_msm_gpu").argtypes = [vp]
        for n in ("tachyon_{c}_g1_point2_msm_gpu", "tachyon_{c}_g1_affine_msm_gpu"):
            f(n).restype = vp
            f(n).argtypes = [vp, vp, vp, sz]
        f("tachyon_{c}_g1_msm_gpu_set_stream_b200").argtypes = [vp, vp]
        f("tachyon_{c}_g1_msm_gpu_set_option_b200").argtypes = [vp, ctypes.c_char_p, ctypes.c_long]
        f("tachyon_{c}_g1_msm_gpu_xyzz_b200").argtypes = [vp, vp, vp, sz, vp]
        f("tachyon_{c}_g1_msm_gpu_last_timing_b200").argtypes = [vp, ctypes.POINTER(MsmTiming)]
        f("tachyon_{c}_g1_generate_bases_b200").argtypes = [u64, sz, sz, vp]
        f("tachyon_{c}_g1_generate_scalars_b200").argtypes = [u64, i32, sz, sz, vp]
        f("tachyon_{c}_g1_point_op_b200").argtypes = [i32, vp, vp, vp, sz]
        f("tachyon_{c}_g1_xyzz_add_b200").restype = None
        f("tachyon_{c}_g1_xyzz_add_b200").argtypes = [vp, vp, vp]
        f("tachyon_{c}_g1_msm_gpu_register_bases_b200").argtypes = [vp, vp, sz]
        f("tachyon_{c}_g1_msm_gpu_commit_batch_b200").argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(sz), sz, vp]
        f("tachyon_{c}_g1_msm_gpu_batch_b200").argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                           ctypes.POINTER(sz), sz, vp]
        f("tachyon_{c}_g1_xyzz_batch_normalize_b200").restype = None
        f("tachyon_{c}_g1_xyzz_batch_normalize_b200").argtypes = [vp, sz, vp]
        f("tachyon_{c}_g1_xyzz_to_jacobian_b200").restype = None
        f("tachyon_{c}_g1_xyzz_to_jacobian_b200").argtypes = [vp, vp]
        f("tachyon_{c}_g1_msm_gpu_join_ranks_b200").argtypes = [vp, vp, i32, i32]
    lib.tachyon_b200_window_bits.restype = ctypes.c_uint32
    lib.tachyon_b200_window_bits.argtypes = [sz, ctypes.c_uint32]
    lib.tachyon_b200_window_count.restype = ctypes.c_uint32
    lib.tachyon_b200_window_count.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    lib.tachyon_b200_last_error.restype = ctypes.c_char_p
    lib.tachyon_b200_imad_peak.restype = ctypes.c_double
    lib.tachyon_b200_imad_peak.argtypes = [i32, i32, i32]
    lib.tachyon_b200_kernel_launch_count.restype = u64
    lib.tachyon_b200_nccl_unique_id.argtypes = [vp]
    lib.tachyon_b200_alloc_host.restype = vp
    lib.tachyon_b200_alloc_host.argtypes = [sz, i32]
    lib.tachyon_b200_free_host.restype = None
    lib.tachyon_b200_free_host.argtypes = [vp]
    _lib = lib
    return lib


def last_error():
    return load().tachyon_b200_last_error().decode()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
