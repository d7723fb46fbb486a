"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bit-exact bar: the normalised affine result
(and every intermediate field/point value tested) must equal the oracle's limbs.
Mirrors the reference's own GPU-vs-CPU tests:
  finite_fields/prime_field_correctness_gpu_test.cc, short_weierstrass/*_correctness_gpu_test.cc,
  msm/variable_base_msm_gpu_unittest.cc:52-78, c/math/elliptic_curves/msm/msm_gpu_unittest.cc:33-67.
"""
import numpy as np
import pytest

from oracle import pymodel
from tachyon_b200 import msm

pytestmark = pytest.mark.gpu

CURVES = ["bn254", "bls12_381"]
G2 = ["bn254_g2", "bls12_381_g2"]          # the same pipeline over Fq2 coordinates
ALL = CURVES + G2


def _consts(name):
    """(r, u64 limbs of one point coordinate) for a G1 or G2 name."""
    if name in pymodel.CURVES:
        c = pymodel.CURVES[name]
        return c.r, c.fq_limbs
    c = pymodel.CURVES_G2[name]
    return c.r, 2 * c.fq_limbs


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    torch.cuda.init()
    return torch


def _rand_elems(c_mod, limbs, n, seed):
    rng = np.random.default_rng(seed)
    vals = [int.from_bytes(rng.bytes(8 * limbs), "little") % c_mod for _ in range(n)]
    vals[:6] = [0, 1, c_mod - 1, c_mod - 2, 2, (1 << (c_mod.bit_length() - 1))]
    return np.array([pymodel.to_limbs(v, limbs) for v in vals], dtype=np.uint64)


@pytest.mark.parametrize("name", CURVES)
def test_fq_ops(oracles, torch_cuda, name):
    o, c = oracles[name], pymodel.CURVES[name]
    a = _rand_elems(c.p, c.fq_limbs, 3000, 1)
    b = _rand_elems(c.p, c.fq_limbs, 3000, 2)[::-1].copy()
    for op in ("add", "sub", "mul"):
        assert (msm.field_op(name, "fq", op, a, b) == o.fq_op(op, a, b)).all(), op
    for op in ("square", "neg", "double"):
        assert (msm.field_op(name, "fq", op, a) == o.fq_op(op, a)).all(), op
    nz = a[6:300]
    assert (msm.field_op(name, "fq", "inverse", nz) == o.fq_op("inverse", nz)).all()
    assert (msm.field_op(name, "fq", "from_mont", a) == o.fq_from_mont(a)).all()
    assert (msm.field_op(name, "fq", "to_mont", a) == o.fq_to_mont(a)).all()


@pytest.mark.parametrize("name", CURVES)
def test_fr_montgomery_conversion(oracles, torch_cuda, name):
    o, c = oracles[name], pymodel.CURVES[name]
    a = _rand_elems(c.r, 4, 3000, 3)
    assert (msm.field_op(name, "fr", "from_mont", a) == o.fr_from_mont(a)).all()
    assert (msm.field_op(name, "fr", "to_mont", a) == o.fr_to_mont(a)).all()


@pytest.mark.parametrize("name", G2)
def test_fq2_ops(oracles, torch_cuda, name):
    """Fq2 = Fq[u]/(u^2 + 1) on the device against the oracle (quadratic_extension_field.h)."""
    o = oracles[name]
    c = pymodel.CURVES_G2[name]
    a = np.concatenate([_rand_elems(c.p, c.fq_limbs, 2000, 5), _rand_elems(c.p, c.fq_limbs, 2000, 6)[::-1]], axis=1)
    b = np.concatenate([_rand_elems(c.p, c.fq_limbs, 2000, 7)[::-1], _rand_elems(c.p, c.fq_limbs, 2000, 8)], axis=1)
    for op in ("add", "sub", "mul"):
        assert (msm.field_op(name, "fq2", op, a, b) == o.fq_op(op, a, b)).all(), op
    for op in ("square", "neg", "double"):
        assert (msm.field_op(name, "fq2", op, a) == o.fq_op(op, a)).all(), op
    nz = a[6:200]
    assert (msm.field_op(name, "fq2", "inverse", nz) == o.fq_op("inverse", nz)).all()


@pytest.mark.parametrize("name", ALL)
def test_point_ops(oracles, torch_cuda, name):
    o = oracles[name]
    n = 96
    aff = o.generate_points(11, n)
    ks = o.fr_from_mont(o.generate_scalars(12, n))
    A = np.stack([o.scalar_mul(aff[i], ks[i]) for i in range(n)])            # XYZZ, zz != 1
    Bx = np.stack([o.scalar_mul(aff[(i * 7 + 3) % n], ks[(i + 1) % n]) for i in range(n)])
    Baff = aff[::-1].copy()
    zero = o.xyzz_zero()
    # exceptional cases: identity operands, P + P, P + (-P), affine identity
    A[0] = zero
    Bx[1] = zero
    Bx[2] = A[2]
    neg = A[3].copy()
    neg[1] = o.fq_op("neg", A[3][1:2])[0]
    Bx[3] = neg
    Baff[4] = 0
    A[5] = np.concatenate([Baff[5].reshape(2, -1), o.constants()["fq_r"][None], o.constants()["fq_r"][None]])
    flatA = A.reshape(n, -1)
    got = msm.point_op(name, "add", flatA, Bx.reshape(n, -1)).reshape(n, 4, -1)
    for i in range(n):
        assert (o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_add(A[i], Bx[i]))).all(), i
    got = msm.point_op(name, "madd", flatA, Baff).reshape(n, 4, -1)
    for i in range(n):
        assert (o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_madd(A[i], Baff[i]))).all(), i
    got = msm.point_op(name, "msub", flatA, Baff).reshape(n, 4, -1)
    for i in range(n):
        nb = Baff[i].reshape(2, -1).copy()
        nb[1] = o.fq_op("neg", nb[1:2])[0]
        assert (o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_madd(A[i], nb))).all(), i
    got = msm.point_op(name, "double", flatA).reshape(n, 4, -1)
    for i in range(n):
        assert (o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_double(A[i]))).all(), i


def _device_inputs(torch, name, o, seed, n, dist):
    """Synthetic test set generated ON the device by the library; returned as
    (bases_tensor, scalars_tensor) of int64 words."""
    fq = o.fq_limbs
    bases = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    msm.generate_bases_device(name, seed, n, bases.data_ptr())
    msm.generate_scalars_device(name, seed + 1, n, scalars.data_ptr(), dist)
    torch.cuda.synchronize()
    return bases, scalars


def _to_np(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("name", ALL)
def test_generators_match_oracle(oracles, torch_cuda, name):
    o = oracles[name]
    n = 4096 + 300
    for dist in ("uniform", "non_uniform", "witness"):
        bases, scalars = _device_inputs(torch_cuda, name, o, 77, n, dist)
        assert (_to_np(scalars) == o.generate_scalars(78, n, dist)).all(), dist
    assert (_to_np(bases) == o.generate_points(77, n)).all()


def _check_msm(o, name, ctx, bases, scalars, dev_bases=None, dev_scalars=None, entry="affine"):
    n = len(scalars)
    want = o.msm_affine(bases, scalars)
    fn = ctx.affine_msm if entry == "affine" else ctx.point2_msm
    jac = fn(bases if dev_bases is None else dev_bases, scalars if dev_scalars is None else dev_scalars, n)
    got = o.jacobian_to_affine(jac)
    assert (got == want).all(), f"{name} n={n}: GPU result differs from CPU oracle"


# sizes of the reference's tests: 32, 2, 5 (c/.../msm_gpu_unittest.cc:33-67), 40
# (pippenger_unittest.cc), 2^10 (variable_base_msm_gpu_unittest.cc:57), plus ragged ones
@pytest.mark.parametrize("name", CURVES)
@pytest.mark.parametrize("n", [0, 1, 2, 5, 32, 33, 40, 1000, 1 << 10, 5000, 1 << 14])
def test_msm_host_inputs(oracles, torch_cuda, name, n):
    o = oracles[name]
    bases, scalars = o.generate_points(100 + n, n), o.generate_scalars(200 + n, n)
    with msm.MSMGpu(name) as ctx:
        _check_msm(o, name, ctx, bases, scalars, entry="affine")
        _check_msm(o, name, ctx, bases, scalars, entry="point2")


@pytest.mark.parametrize("name", G2)
@pytest.mark.parametrize("n", [0, 1, 2, 5, 33, 1000, 5000])
def test_g2_msm_host_inputs(oracles, torch_cuda, name, n):
    # VariableBaseMSMGpu<G2AffinePoint> (groth16/prove.h:129-131) vs the CPU path
    o = oracles[name]
    bases, scalars = o.generate_points(100 + n, n), o.generate_scalars(200 + n, n)
    with msm.MSMGpu(name) as ctx:
        _check_msm(o, name, ctx, bases, scalars, entry="affine")
        _check_msm(o, name, ctx, bases, scalars, entry="point2")


@pytest.mark.parametrize("name", G2)
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_g2_accumulate_variants(oracles, torch_cuda, name, variant):
    # 0: one thread per accumulation task; 1 / 2: one LANE PAIR per task (a lane per Fq2 component)
    # at both register budgets.  Covers what only the accumulation sees: bucket values carried
    # across point ranges, buckets split into several tasks, identity bases, equal points
    # (the doubling branch) and warps whose pairs run different task lengths.
    o = oracles[name]
    r_mod, el = _consts(name)
    n = 3000
    bases, scalars = o.generate_points(51, n), o.generate_scalars(52, n, "non_uniform")
    bases[5] = 0
    bases[9] = bases[8]
    scalars[9] = scalars[8]
    bases[200:260] = bases[200]
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("acc_variant", variant)
        _check_msm(o, name, ctx, bases, scalars)
        ctx.set_option("ranges", 3)            # later ranges start from the stored bucket values
        _check_msm(o, name, ctx, bases, scalars)
        ctx.set_option("segment", 16)          # split buckets: partial sums to task_out, folded
        ctx.set_option("window_bits", 6)
        _check_msm(o, name, ctx, bases, scalars)
        same = bases.copy()
        same[:] = bases[7]
        _check_msm(o, name, ctx, same, o.generate_scalars(53, n))


@pytest.mark.parametrize("name", CURVES)
def test_msm_accumulate_lockstep(oracles, torch_cuda, name):
    # acc_variant 3 (G1): the warps of a CTA walk their tasks one addition per barrier
    o = oracles[name]
    n = 5000
    bases, scalars = o.generate_points(71, n), o.generate_scalars(72, n, "non_uniform")
    bases[5] = 0
    bases[9] = bases[8]
    scalars[9] = scalars[8]
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("acc_variant", 3)
        _check_msm(o, name, ctx, bases, scalars)
        ctx.set_option("ranges", 3)
        ctx.set_option("segment", 16)
        _check_msm(o, name, ctx, bases, scalars)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("roll", [0, 1, 2])
def test_msm_reduce_code_shapes(oracles, torch_cuda, name, roll):
    # the field multiplications of the running-sum kernel unrolled (0), looped (1), looped with
    # the squarings through the multiplier (2): same values
    o = oracles[name]
    n = 4000
    bases, scalars = o.generate_points(61, n), o.generate_scalars(62, n)
    bases[9] = bases[8]
    scalars[9] = scalars[8]
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("reduce_roll", roll)
        _check_msm(o, name, ctx, bases, scalars)
        ctx.set_option("window_bits", 7)
        ctx.set_option("ranges", 2)
        _check_msm(o, name, ctx, bases, o.generate_scalars(63, n, "non_uniform"))


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("dist", ["non_uniform", "witness"])
def test_msm_skewed_scalars(oracles, torch_cuda, name, dist):
    o = oracles[name]
    n = 20000
    bases, scalars = o.generate_points(301, n), o.generate_scalars(302, n, dist)
    with msm.MSMGpu(name) as ctx:
        _check_msm(o, name, ctx, bases, scalars)
        ctx.set_option("segment", 16)      # force heavy bucket splitting + folding
        _check_msm(o, name, ctx, bases, scalars)
        t = ctx.last_timing()
        assert t["tasks"] > 0 and t["entries"] > 0


@pytest.mark.parametrize("name", ALL)
def test_msm_edge_cases(oracles, torch_cuda, name):
    o = oracles[name]
    r_mod, el = _consts(name)
    n = 600
    bases, scalars = o.generate_points(41, n), o.generate_scalars(42, n, "witness")
    bases[3] = 0                                    # identity base (0, 0)
    bases[9] = bases[8]                             # duplicate point, equal scalar -> doubling branch
    scalars[9] = scalars[8]
    neg_y = o.fq_op("neg", bases[10].reshape(2, -1)[1:2])
    bases[11] = np.concatenate([bases[10][:el], neg_y[0]])   # P and -P with equal scalars
    scalars[11] = scalars[10]
    for i, v in enumerate((r_mod - 1, 1, 0, (1 << 253) + 12345, r_mod - 2)):
        scalars[20 + i] = o.fr_to_mont(np.array(pymodel.to_limbs(v % r_mod, 4), dtype=np.uint64))[0]
    bases[100:140] = bases[100]                     # many copies of one point, random scalars
    with msm.MSMGpu(name) as ctx:
        _check_msm(o, name, ctx, bases, scalars)
        z = np.zeros_like(scalars)                  # all-zero scalars -> identity (z == 0)
        jac = ctx.affine_msm(bases, z)
        assert (jac[2] == 0).all()
        same = bases.copy()
        same[:] = bases[7]
        _check_msm(o, name, ctx, same, o.generate_scalars(43, n))   # all bases equal


@pytest.mark.parametrize("name", ALL)
def test_msm_window_sweep(oracles, torch_cuda, name):
    o = oracles[name]
    n = 3000
    bases, scalars = o.generate_points(51, n), o.generate_scalars(52, n)
    # scalars that reach the top of the top window and carry through every window
    r_mod, _ = _consts(name)
    edge = (r_mod - 1, r_mod - 2, (1 << (r_mod.bit_length() - 1)) - 1, 1 << (r_mod.bit_length() - 1),
            r_mod >> 1, (r_mod >> 1) + 1, int("55" * 31, 16), int("aa" * 31, 16))
    for i, v in enumerate(edge):
        scalars[i] = o.fr_to_mont(np.array(pymodel.to_limbs(v % r_mod, 4), dtype=np.uint64))[0]
    want = o.msm_affine(bases, scalars)
    with msm.MSMGpu(name) as ctx:
        for balance in (1, 0):       # balanced windows (widths c / c - 1) and equal widths
            ctx.set_option("balance", balance)
            for cbits in (4, 5, 6, 7, 8, 11, 13, 16):
                ctx.set_option("window_bits", cbits)
                got = o.jacobian_to_affine(ctx.affine_msm(bases, scalars))
                assert (got == want).all(), (balance, cbits)
                assert ctx.last_timing()["window_bits"] == cbits


@pytest.mark.parametrize("name", ALL)
def test_msm_device_resident_inputs(oracles, torch_cuda, name):
    o = oracles[name]
    n = 1 << 13
    bases, scalars = _device_inputs(torch_cuda, name, o, 61, n, "uniform")
    hb, hs = _to_np(bases), _to_np(scalars)
    with msm.MSMGpu(name) as ctx:
        _check_msm(o, name, ctx, hb, hs, dev_bases=bases.data_ptr(), dev_scalars=scalars.data_ptr())
        # device bases + host scalars: the KZG calling convention (kzg.h:267-284)
        _check_msm(o, name, ctx, hb, hs, dev_bases=bases.data_ptr())
        assert ctx.last_timing()["kernel_launches"] >= 8


# One MSM consumed as several point ranges that share the bucket values (what the engine
# does with host inputs so that H2D overlaps the bucket work, and when memory is short;
# the analogue of the chunk loop of icicle_msm_bn254_g1.cc:56-73): every split must give
# the same group element, with ragged last ranges, more ranges than staging slots, skewed
# scalars whose buckets are split into several tasks in every range, and device inputs.
@pytest.mark.parametrize("name", CURVES)
def test_msm_point_ranges(oracles, torch_cuda, name):
    o = oracles[name]
    n = 6001
    bases, scalars = o.generate_points(81, n), o.generate_scalars(82, n)
    skew = o.generate_scalars(83, n, "witness")
    want, want_skew = o.msm_affine(bases, scalars), o.msm_affine(bases, skew)
    import torch
    db = torch.from_numpy(bases.view(np.int64)).cuda()
    ds = torch.from_numpy(scalars.view(np.int64)).cuda()
    with msm.MSMGpu(name) as ctx:
        for ranges in (1, 2, 3, 4, 7, 64):
            ctx.set_option("ranges", ranges)
            assert (o.jacobian_to_affine(ctx.affine_msm(bases, scalars)) == want).all(), ranges
            assert ctx.last_timing()["ranges"] == ranges
            assert (o.jacobian_to_affine(ctx.affine_msm(db.data_ptr(), ds.data_ptr(), n)) == want).all(), ranges
            assert (o.jacobian_to_affine(ctx.affine_msm(db.data_ptr(), scalars, n)) == want).all(), ranges
        ctx.set_option("ranges", 5)
        ctx.set_option("segment", 16)
        ctx.set_option("window_bits", 6)
        assert (o.jacobian_to_affine(ctx.affine_msm(bases, skew)) == want_skew).all()
        ctx.set_option("release_workspace", 1)   # buffers are re-allocated on demand
        ctx.set_option("prewarm", 1 << 14)
        ctx.set_option("ranges", 0)       # automatic again
        ctx.set_option("segment", 0)
        ctx.set_option("window_bits", 0)
        assert (o.jacobian_to_affine(ctx.affine_msm(bases, skew)) == want_skew).all()


# The experimental batched-affine pre-reduction (pair rounds: affine + affine additions sharing
# one inversion per thread batch) must give the same group element for
# every round count, including the doubling / cancelling / identity cases inside a pair.
@pytest.mark.parametrize("name", CURVES)
def test_msm_pair_rounds(oracles, torch_cuda, name):
    o, c = oracles[name], pymodel.CURVES[name]
    n = 5000
    bases, scalars = o.generate_points(111, n), o.generate_scalars(112, n)
    bases[3] = 0
    bases[9] = bases[8]
    scalars[9] = scalars[8]                                   # P + P inside one bucket
    neg_y = o.fq_op("neg", bases[10].reshape(2, -1)[1:2])
    bases[11] = np.concatenate([bases[10][:c.fq_limbs], neg_y[0]])
    scalars[11] = scalars[10]                                 # P + (-P)
    bases[100:164] = bases[100]
    scalars[100:164] = scalars[100]                           # 64 copies: doubling chain through every round
    skew = o.generate_scalars(113, n, "witness")
    want, want_skew = o.msm_affine(bases, scalars), o.msm_affine(bases, skew)
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("window_bits", 7)                      # ~78 entries per bucket
        for rounds in (1, 2, 3, 4, -2):
            ctx.set_option("pair_rounds", rounds)
            assert (o.jacobian_to_affine(ctx.affine_msm(bases, scalars)) == want).all(), rounds
            if rounds > 0:
                assert ctx.last_timing()["pair_rounds"] == rounds
            assert (o.jacobian_to_affine(ctx.affine_msm(bases, skew)) == want_skew).all(), rounds
        ctx.set_option("ranges", 3)
        ctx.set_option("pair_rounds", 2)
        assert (o.jacobian_to_affine(ctx.affine_msm(bases, scalars)) == want).all()


# Both sort implementations (one-level atomic counting sort, two-level shared-memory sort of
# msm_sort.cuh) must feed the accumulation the same buckets: uniform, skewed and degenerate
# scalars, ragged sizes, several ranges.
@pytest.mark.parametrize("name", CURVES)
def test_msm_sort_modes(oracles, torch_cuda, name):
    o = oracles[name]
    n = (1 << 16) + 777
    bases = o.generate_points(121, n)
    with msm.MSMGpu(name) as ctx:
        for dist in ("uniform", "witness", "non_uniform"):
            scalars = o.generate_scalars(122, n, dist)
            want = o.msm_affine(bases, scalars)
            # -1: automatic — the scalar sample flags the skewed sets for the two-level sort
            for mode, cbits, ranges in ((0, 0, 0), (1, 0, 0), (1, 12, 0), (1, 17, 3), (1, 21, 1), (-1, 0, 0),
                                        (-1, 15, 2)):
                ctx.set_option("sort_mode", mode)
                ctx.set_option("window_bits", cbits)
                ctx.set_option("ranges", ranges)
                got = o.jacobian_to_affine(ctx.affine_msm(bases, scalars))
                assert (got == want).all(), (dist, mode, cbits, ranges)


# Bucket reduction: two threads per block of buckets (default) and the one-thread form, with
# balanced and equal-width windows, for window sizes that give few / many blocks per window,
# and the host-input pipeline cut into a different number of ranges.
@pytest.mark.parametrize("name", ALL)
def test_msm_reduce_modes(oracles, torch_cuda, name):
    o = oracles[name]
    n = 6000 if name in CURVES else 2500
    bases, scalars = o.generate_points(131, n), o.generate_scalars(132, n)
    want = o.msm_affine(bases, scalars)
    with msm.MSMGpu(name) as ctx:
        for mode in (1, 0) if name in CURVES else (1, 2, 0):   # G2: 1 = lane pairs, 2 = two threads
            ctx.set_option("reduce_mode", mode)
            for balance, cbits in ((1, 0), (0, 0), (1, 5), (1, 9), (1, 14), (0, 14), (1, 18)):
                ctx.set_option("balance", balance)
                ctx.set_option("window_bits", cbits)
                got = o.jacobian_to_affine(ctx.affine_msm(bases, scalars))
                assert (got == want).all(), (mode, balance, cbits)
        ctx.set_option("reduce_mode", 1)
        ctx.set_option("balance", 1)
        for inline in (1, 0, -1):              # G1: one inlined call site of the addition
            ctx.set_option("reduce_inline", inline)
            for cbits in (0, 9, 14):
                ctx.set_option("window_bits", cbits)
                got = o.jacobian_to_affine(ctx.affine_msm(bases, scalars))
                assert (got == want).all(), (inline, cbits)
        ctx.set_option("window_bits", 0)
        for host_ranges in (1, 3, 16):
            ctx.set_option("host_ranges", host_ranges)
            got = o.jacobian_to_affine(ctx.affine_msm(bases, scalars))
            assert (got == want).all(), host_ranges


# Window combination (pippenger_base.h:59-77).  Default: the device reduces every window to its sum
# S_w (running-sum level, fused merge tree, per-window combine kernel) and the host runs the final,
# strictly sequential ladder over the W sums.  Option "device_ladder": the ladder runs as a kernel
# too, and the windows of the last point range are accumulated as two groups (high first) so that
# the high group's merge tree and doubling chain run on a second stream behind the low group's
# accumulation.
# Every split, window size, range count and scalar distribution must give the same point; skewed
# scalars exercise the per-group filter of the bucket-fold kernels.
@pytest.mark.parametrize("name", ALL)
def test_msm_window_groups(oracles, torch_cuda, name):
    o = oracles[name]
    n = 5000 if name in CURVES else 2000
    bases = o.generate_points(141, n)
    cases = {"uniform": o.generate_scalars(142, n), "witness": o.generate_scalars(143, n, "witness"),
             "non_uniform": o.generate_scalars(144, n, "non_uniform")}
    want = {k: o.msm_affine(bases, v) for k, v in cases.items()}
    with msm.MSMGpu(name) as ctx:
        # default: the device leaves one sum per window, the host runs the final ladder
        for cbits in (0, 5, 9, 13):
            ctx.set_option("window_bits", cbits)
            for dist, sc in cases.items():
                got = o.jacobian_to_affine(ctx.affine_msm(bases, sc))
                assert (got == want[dist]).all(), ("host ladder", cbits, dist)
                assert ctx.last_timing()["low_windows"] == 0
        ctx.set_option("device_ladder", 1)             # the ladder as a kernel too: ONE point leaves
        for cbits in (0, 5, 9, 13):
            ctx.set_option("window_bits", cbits)
            for low in (-1, 0, 1, 2, 5, 200):          # 200: clamped to W - 1
                ctx.set_option("low_windows", low)
                for ranges in (1, 3):
                    ctx.set_option("ranges", ranges)
                    for dist, sc in cases.items():
                        if dist != "uniform" and (ranges == 3 or cbits == 13):
                            continue
                        got = o.jacobian_to_affine(ctx.affine_msm(bases, sc))
                        assert (got == want[dist]).all(), (cbits, low, ranges, dist)
                        t = ctx.last_timing()
                        if low in (0, 1, 2, 5):
                            assert t["low_windows"] == low
        ctx.set_option("low_windows", 2)
        ctx.set_option("window_bits", 6)
        ctx.set_option("ranges", 1)
        ctx.set_option("segment", 16)                  # every bucket split: fold jobs in both groups
        got = o.jacobian_to_affine(ctx.affine_msm(bases, cases["witness"]))
        assert (got == want["witness"]).all()


# The H2D staging ring is grow-only; its slot stride must follow THIS call's range size, not the
# size the buffer happens to have (a stride a few bytes short of a range let the copy stream
# overwrite the tail of a range the kernels had not consumed yet).  Calls of different sizes and
# range counts on one context, sizes chosen around the 1/3 marks of the reserved buffer.
@pytest.mark.parametrize("name", CURVES)
def test_msm_staging_ring_reuse(oracles, torch_cuda, name):
    o = oracles[name]
    nmax = 33000
    bases, scalars = o.generate_points(151, nmax), o.generate_scalars(152, nmax)
    with msm.MSMGpu(name) as ctx:
        for n, ranges in ((9000, 1), (32763, 3), (32764, 3), (10921, 1), (10922, 1), (32765, 3),
                          (32767, 3), (21845, 2), (33000, 3), (10923, 4)):
            ctx.set_option("ranges", ranges)
            got = o.jacobian_to_affine(ctx.affine_msm(bases[:n], scalars[:n]))
            assert (got == o.msm_affine(bases[:n], scalars[:n])).all(), (n, ranges)


# Process-level sharding plumbing that needs no second GPU: the NCCL library resolves, a unique
# id is produced, joining a world of one is a no-op and an MSM still gives the right point.  (The
# N > 1 exchange itself runs in bench.py under torchrun, with its own parity check.)
def test_join_ranks_world_of_one(oracles, torch_cuda):
    o = oracles["bn254"]
    uid = msm.nccl_unique_id()
    assert len(uid) == 128 and any(uid)
    bases, scalars = o.generate_points(201, 777), o.generate_scalars(202, 777)
    with msm.MSMGpu("bn254") as ctx:
        ctx.join_ranks(uid, 0, 1)
        _check_msm(o, "bn254", ctx, bases, scalars)
        with pytest.raises(RuntimeError):
            ctx.join_ranks(uid, 3, 2)               # rank outside the world


# bases registered once (the SRS of kzg.h:91-113), then a batch of
# commitments with fresh scalars (kzg.h:217-313), results batch-normalised (point_xyzz.h:109-163).
@pytest.mark.parametrize("name", CURVES)
def test_registered_bases_commit_batch(oracles, torch_cuda, name):
    o = oracles[name]
    n = 3000
    bases = o.generate_points(91, n)
    sizes = [n, 1, 0, 2500, n, 17, n]
    scal = [o.generate_scalars(92 + i, s, "witness" if i % 3 == 2 else "uniform") for i, s in enumerate(sizes)]
    want = [o.msm_affine(bases[:s], scal[i]) if s else np.zeros_like(bases[0]).reshape(2, -1)
            for i, s in enumerate(sizes)]
    with msm.MSMGpu(name) as ctx:
        ctx.register_bases(bases)
        got = msm.batch_normalize(name, ctx.commit_batch(scal, sizes))
        for i in range(len(sizes)):
            assert (got[i] == np.asarray(want[i]).reshape(-1)).all(), i
        # device-resident scalars, and a re-registration from a device pointer with fewer bases
        import torch
        dsc = [torch.from_numpy(s.view(np.int64)).cuda() for s in scal]
        got = msm.batch_normalize(name, ctx.commit_batch([d.data_ptr() if s else 0 for d, s in zip(dsc, sizes)], sizes))
        for i in range(len(sizes)):
            assert (got[i] == np.asarray(want[i]).reshape(-1)).all(), i
        db = torch.from_numpy(bases[:2500].view(np.int64)).cuda()
        ctx.register_bases(db.data_ptr(), 2500)
        got = msm.batch_normalize(name, ctx.commit_batch([scal[3]], [2500]))
        assert (got[0] == np.asarray(want[3]).reshape(-1)).all()
        with pytest.raises(RuntimeError):
            ctx.commit_batch([scal[0]], [n])      # larger than what is registered now
        ndev = msm.device_count()
        if ndev > 1:
            ctx.set_option("devices", min(ndev, 4))
            ctx.register_bases(bases)
            got = msm.batch_normalize(name, ctx.commit_batch(scal, sizes))
            for i in range(len(sizes)):
                assert (got[i] == np.asarray(want[i]).reshape(-1)).all(), i


# Registered bases with the table of window multiples (option "precompute"; the role of
# precompute_factor in algorithms/icicle/icicle_msm.h:21): all windows share one bucket set, the
# digit of window w selects 2^(bit offset of w) * P from the table.  Same commitments, bit for
# bit, for several window sizes (incl. balanced narrow top windows), ragged sizes, skewed
# scalars, point ranges, identity bases, G2, and the general batch call mixing table and
# explicit bases.
@pytest.mark.parametrize("name", ALL)
def test_registered_bases_precomputed_table(oracles, torch_cuda, name):
    o = oracles[name]
    n = 3000 if name in CURVES else 900
    bases = o.generate_points(191, n)
    bases[5] = 0                                      # an identity base stays the identity in every slice
    sizes = [n, 1, 0, n - 400, 17, n]
    scal = [o.generate_scalars(192 + i, s, ("witness", "uniform", "non_uniform")[i % 3]) for i, s in enumerate(sizes)]
    r_mod, _ = _consts(name)
    scal[0][0] = o.fr_to_mont(np.array(pymodel.to_limbs(r_mod - 1, 4), dtype=np.uint64))[0]
    want = [np.asarray(o.msm_affine(bases[:s], scal[i])).reshape(-1) if s else np.zeros(bases.shape[1], dtype=np.uint64)
            for i, s in enumerate(sizes)]
    other = o.generate_points(193, 500)
    other_sc = o.generate_scalars(194, 500)
    other_want = np.asarray(o.msm_affine(other, other_sc)).reshape(-1)
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("precompute", 1)
        for cbits in (0, 5, 8, 13):
            ctx.set_option("window_bits", cbits)
            ctx.register_bases(bases)
            for ranges in (0, 3):
                ctx.set_option("ranges", ranges)
                got = msm.batch_normalize(name, ctx.commit_batch(scal, sizes))
                for i in range(len(sizes)):
                    assert (got[i] == want[i]).all(), (cbits, ranges, i)
            ctx.set_option("ranges", 0)
        ctx.set_option("window_bits", 0)
        # the general batch: registered table for one MSM, explicit bases for the other
        got = msm.batch_normalize(name, ctx.msm_batch([None, other], [scal[5], other_sc], [n, 500]))
        assert (got[0] == want[5]).all() and (got[1] == other_want).all()
        # a plain MSM call on the same context is unaffected by the table
        assert (np.asarray(o.jacobian_to_affine(ctx.affine_msm(other, other_sc))).reshape(-1) == other_want).all()
        ctx.set_option("precompute", 0)               # back to a plain registration
        ctx.register_bases(bases)
        got = msm.batch_normalize(name, ctx.commit_batch(scal, sizes))
        for i in range(len(sizes)):
            assert (got[i] == want[i]).all(), i
    if name == "bls12_381":
        # more than 1 MiB of pageable host bases: the registration copy must have landed before
        # the table kernel reads it (a plain cudaMemcpy may return while its last staged chunk is
        # in flight and the engine's non-blocking stream does not wait for the legacy stream)
        n = 11708
        bases, sc = o.generate_points(195, n), o.generate_scalars(196, n)
        want = np.asarray(o.msm_affine(bases, sc)).reshape(-1)
        for rep in range(4):
            with msm.MSMGpu(name) as ctx:
                ctx.set_option("precompute", 1)
                ctx.set_option("window_bits", 9)
                ctx.register_bases(bases)
                assert (msm.batch_normalize(name, ctx.commit_batch([sc], [n]))[0] == want).all(), rep


# BASELINE.json configs[4] in miniature: the four G1 MSMs of a Groth16 proof (A, B1 over the
# full assignment, L over its witness part, H over uniform coefficients; prove.h:100-131),
# each with its own bases, through the general batch call.
@pytest.mark.parametrize("name", CURVES)
def test_groth16_msm_set_batch(oracles, torch_cuda, name):
    import torch
    o = oracles[name]
    m, n_pub = 1 << 12, 10
    full = o.generate_scalars(131, m, "witness")
    hco = o.generate_scalars(132, m, "uniform")
    sizes = [m, m, m - n_pub, m]
    bases = [o.generate_points(140 + j, n) for j, n in enumerate(sizes)]
    scal = [full, full, full[n_pub:].copy(), hco]
    want = [o.msm_affine(b, s) for b, s in zip(bases, scal)]
    with msm.MSMGpu(name) as ctx:
        got = msm.batch_normalize(name, ctx.msm_batch(bases, scal))
        for j in range(4):
            assert (got[j] == np.asarray(want[j]).reshape(-1)).all(), j
        # resident proving key, host scalars; one MSM falls back to the registered bases
        db = [torch.from_numpy(b.view(np.int64)).cuda() for b in bases]
        ctx.register_bases(bases[3])
        got = msm.batch_normalize(name, ctx.msm_batch([db[0].data_ptr(), db[1].data_ptr(), db[2].data_ptr(), None],
                                                      scal, sizes))
        for j in range(4):
            assert (got[j] == np.asarray(want[j]).reshape(-1)).all(), j
        if msm.device_count() > 1:
            ctx.set_option("devices", min(msm.device_count(), 4))
            got = msm.batch_normalize(name, ctx.msm_batch(bases, scal))
            for j in range(4):
                assert (got[j] == np.asarray(want[j]).reshape(-1)).all(), j


# SURVEY 8f-3 (MSM part): the Groth16 proof from assignments — L, H, A, B1 batched on the G1
# context, B2 concurrently on the G2 context, r / s blinding on the host — against the oracle's
# restatement of zk/r1cs/groth16/prove.h:33-165 (itself pinned to the Python model).
@pytest.mark.parametrize("curve", CURVES)
def test_groth16_prove(oracles, torch_cuda, curve):
    import torch
    from oracle import cpu_oracle
    from tests.groth16_util import make_case
    with msm.MSMGpu(curve) as g1, msm.MSMGpu(curve + "_g2") as g2:
        for blind, h_extra in ((True, 1), (False, 0)):
            pk, r, s, h, witness, full = make_case(oracles, curve, n_full=3000, n_pub=17, h_size=2048 + h_extra,
                                                   seed=70, h_query_size=2048, blind=blind)
            want = cpu_oracle.groth16_prove(curve, pk, r, s, h, witness, full)
            got = msm.groth16_prove(g1, g2, pk, r, s, h, witness, full)
            for a, b in zip(got, want):
                assert (a == b).all(), (blind, h_extra)
        # device-resident proving key (a prover that keeps its zkey loaded)
        dev = {k: torch.from_numpy(v.view(np.int64)).cuda() for k, v in pk.items() if k.endswith("_query")}
        pk_dev = dict(pk)
        for k, t in dev.items():
            pk_dev[k] = (t.data_ptr(), t.shape[0])
        got = msm.groth16_prove(g1, g2, pk_dev, r, s, h, witness, full)
        for a, b in zip(got, want):
            assert (a == b).all()
        if msm.device_count() > 1:
            g1.set_option("devices", min(msm.device_count(), 4))
            got = msm.groth16_prove(g1, g2, pk, r, s, h, witness, full)
            for a, b in zip(got, want):
                assert (a == b).all()
        with pytest.raises(RuntimeError):
            msm.groth16_prove(g1, g2, pk, r, s, h, witness[:-1], full)   # l query / witness size mismatch


# SURVEY 8f-3, file level (vendors/circom/prover_main.cc:81-186): proof straight from a snarkjs
# .zkey and a .wtns — the reference-held fixtures multiplier_3.zkey / multiplier_3.wtns — through
# tachyon_bn254_groth16_prove_from_files_b200, against the oracle's groth16_prove fed by the
# independent Python model of the formats and of the witness map (oracle/circom_model.py); the
# snarkjs JSON it writes must carry the same points as decimal strings.
@pytest.mark.parametrize("blind", [True, False])
def test_groth16_prove_from_files(oracles, torch_cuda, tmp_path, blind):
    import json
    import os
    from oracle import circom_model, cpu_oracle
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    zkey, wtns = os.path.join(golden, "multiplier_3.zkey"), os.path.join(golden, "multiplier_3.wtns")
    o = oracles["bn254"]
    z = circom_model.parse_zkey(zkey)
    _, w = circom_model.parse_wtns(wtns)
    as_fr = lambda ints: o.fr_to_mont(np.array([pymodel.to_limbs(v, 4) for v in ints], dtype=np.uint64))
    pts = lambda raw, limbs: np.frombuffer(raw, dtype=np.uint64).reshape(-1, limbs).copy()
    pk = {"alpha_g1": pts(z["alpha_g1"], 8)[0], "beta_g1": pts(z["beta_g1"], 8)[0], "delta_g1": pts(z["delta_g1"], 8)[0],
          "beta_g2": pts(z["beta_g2"], 16)[0], "delta_g2": pts(z["delta_g2"], 16)[0],
          "a_g1_query": pts(z["a_g1"], 8), "b_g1_query": pts(z["b_g1"], 8), "b_g2_query": pts(z["b_g2"], 16),
          "h_g1_query": pts(z["h_g1"], 8), "l_g1_query": pts(z["c_g1"], 8)}
    n_inst = z["num_public"] + 1
    full, witness = as_fr(w[1:]), as_fr(w[n_inst:])
    h = as_fr(circom_model.witness_map("bn254", z, w))
    rs = o.generate_scalars(171, 2, "uniform")
    r, s_ = (rs[0], rs[1]) if blind else (np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
    want = cpu_oracle.groth16_prove("bn254", pk, r, s_, h, witness, full)
    pj, uj = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    with msm.MSMGpu("bn254") as g1, msm.MSMGpu("bn254_g2") as g2:
        got = msm.groth16_prove_from_files(g1, g2, zkey, wtns, r if blind else None, s_ if blind else None, pj, uj)
    for a, b in zip(got, want):
        assert (np.asarray(a).reshape(-1) == np.asarray(b).reshape(-1)).all()
    c = pymodel.CURVES["bn254"]
    dec = lambda limbs: str(pymodel.from_limbs(limbs) * pow(c.fq_R, -1, c.p) % c.p)
    a, b, cc = (np.asarray(x).reshape(-1, 4) for x in want)
    proof = json.load(open(pj))
    assert proof["pi_a"] == [dec(a[0]), dec(a[1]), "1"] and proof["pi_c"] == [dec(cc[0]), dec(cc[1]), "1"]
    assert proof["pi_b"] == [[dec(b[0]), dec(b[1])], [dec(b[2]), dec(b[3])], ["1", "0"]]
    assert proof["protocol"] == "groth16" and proof["curve"] == "bn128"
    assert json.load(open(uj)) == [str(v) for v in w[1:n_inst]] == ["60"]


# the dump written under TACHYON_MSM_GPU_INPUT_DIR (msm_gpu.h:99-119: u64 count,
# canonical little-endian limbs) and the replay CLI (msm_gpu_replay.cc:40-88: --idx --degree
# --input_dir, prints the time and the affine point as hex without leading zeros).
@pytest.mark.parametrize("name", CURVES)
def test_dump_and_replay_cli(oracles, torch_cuda, name, tmp_path, monkeypatch):
    import os
    import subprocess
    o, c = oracles[name], pymodel.CURVES[name]
    n = 700
    bases, scalars = o.generate_points(95, n), o.generate_scalars(96, n)
    bases[5] = 0
    monkeypatch.setenv("TACHYON_MSM_GPU_INPUT_DIR", str(tmp_path))
    with msm.MSMGpu(name) as ctx:
        ctx.affine_msm(bases, scalars)
        ctx.point2_msm(bases[:40], scalars[:40])
    monkeypatch.delenv("TACHYON_MSM_GPU_INPUT_DIR")
    raw = np.fromfile(tmp_path / "bases0.txt", dtype=np.uint64)
    assert raw[0] == n and raw.size == 1 + n * 2 * c.fq_limbs
    assert (raw[1:].reshape(n, -1) == np.concatenate(
        [o.fq_from_mont(bases[:, :c.fq_limbs]), o.fq_from_mont(bases[:, c.fq_limbs:])], axis=1)).all()
    raw = np.fromfile(tmp_path / "scalars1.txt", dtype=np.uint64)
    assert raw[0] == 40 and (raw[1:].reshape(40, 4) == o.fr_from_mont(scalars[:40])).all()
    exe = os.path.join(os.path.dirname(msm._lib.LIB_PATH), "msm_gpu_replay")
    env = {k: v for k, v in os.environ.items() if k != "TACHYON_MSM_GPU_INPUT_DIR"}
    out = subprocess.run([exe, "--idx", "0,1", "--degree", "10", "--input_dir", str(tmp_path), "--curve", name],
                         capture_output=True, text=True, env=env, check=True).stdout.splitlines()
    points = [ln for ln in out if ln.startswith("(0x")]
    assert len(points) == 2
    for line, k in zip(points, (n, 40)):
        want = o.fq_from_mont(o.msm_affine(bases[:k], scalars[:k]))
        hx = ["0x%x" % pymodel.from_limbs([int(v) for v in want[i]]) for i in range(2)]
        assert line == "(%s, %s)" % tuple(hx)
    # the reference refuses to run with the dump variable set (msm_gpu_replay.cc:41-44)
    bad = subprocess.run([exe, "--idx", "0", "--degree", "10", "--input_dir", str(tmp_path)],
                         capture_output=True, text=True, env=dict(env, TACHYON_MSM_GPU_INPUT_DIR=str(tmp_path)))
    assert bad.returncode == 1


# Full benchmark sizes through a size-independent property: the synthetic bases
# are chains P_(j,d) = 2^d H_j, so MSM(P, s) == MSM(H, fold(s)) with
# fold(s)_j = sum_d s_(j,d) 2^d mod r — a 2^12-times smaller MSM the oracle does in
# milliseconds.
# (bls12_381 at 2^20: above 12 M entries the 12-limb kernel runs its 4-CTAs/SM build, below its 3-CTA one)
@pytest.mark.parametrize("name,logn", [("bn254", 16), ("bn254", 20), ("bls12_381", 18), ("bls12_381", 20),
                                       ("bn254_g2", 16), ("bls12_381_g2", 14)])
def test_msm_full_size_chain_fold(oracles, torch_cuda, name, logn):
    o = oracles[name]
    n = 1 << logn
    bases, scalars = _device_inputs(torch_cuda, name, o, 71, n, "uniform")
    hs = _to_np(scalars)
    assert (hs[:64] == o.generate_scalars(72, 64)).all()
    heads = o.generate_points(71, n)[::4096] if n <= (1 << 16) else np.stack(
        [o.generate_points(71, 1, first=j * 4096)[0] for j in range(n // 4096)])
    want = o.msm_affine(heads, o.fold_chain_scalars(hs))
    with msm.MSMGpu(name) as ctx:
        got = o.jacobian_to_affine(ctx.affine_msm(bases.data_ptr(), scalars.data_ptr(), n))
        assert (got == want).all()
        # linearity: MSM over the two halves adds up to the whole (host XYZZ add)
        half = n // 2
        fq = o.fq_limbs
        a = ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), half)
        b = ctx.msm_xyzz(bases.data_ptr() + half * 2 * fq * 8, scalars.data_ptr() + half * 32, half)
        assert (o.xyzz_to_affine(msm.xyzz_add(name, a, b)) == want).all()


# A forced small window makes entries = points x windows the limiting quantity (u32 offsets):
# the call must split itself into pieces rather than overflow.
def test_msm_small_window_many_entries(oracles, torch_cuda):
    name, n = "bn254", 1 << 16
    o = oracles[name]
    bases, scalars = o.generate_points(191, n), o.generate_scalars(192, n)
    want = o.msm_affine(bases, scalars)
    with msm.MSMGpu(name) as ctx:
        ctx.set_option("window_bits", 4)          # W = 64 windows
        assert (o.jacobian_to_affine(ctx.affine_msm(bases, scalars)) == want).all()
        assert ctx.last_timing()["windows"] == 64


# Maximum sizes: an MSM above the engine's 2^26-point piece limit (u32 index arithmetic) runs as
# independent pieces whose sums are added on the host (the serial chunk loop of
# icicle_msm_bn254_g1.cc:56-73, without its dropped remainder); checked through chain-fold.
def test_msm_beyond_piece_limit(oracles, torch_cuda):
    name, n = "bn254", (1 << 26) + 3 * 4096
    o = oracles[name]
    bases, scalars = _device_inputs(torch_cuda, name, o, 171, n, "uniform")
    hs = _to_np(scalars)
    heads = np.stack([o.generate_points(171, 1, first=j * 4096)[0] for j in range(n // 4096)])
    want = o.msm_affine(heads, o.fold_chain_scalars(hs))
    with msm.MSMGpu(name) as ctx:
        got = o.jacobian_to_affine(ctx.affine_msm(bases.data_ptr(), scalars.data_ptr(), n))
        assert (got == want).all()


# Host inputs large enough for the pipelined paths: pageable numpy memory goes through the pinned
# bounce buffers filled by host threads (>= 8 MB per copy), pinned memory straight to the DMA
# engine; several geometric point ranges; both against the chain-fold value.
@pytest.mark.parametrize("name,logn", [("bn254", 20), ("bls12_381", 19)])
def test_msm_large_host_inputs_pageable_and_pinned(oracles, torch_cuda, name, logn):
    import torch
    o = oracles[name]
    n = (1 << logn) + 4096 * 3 + 0
    bases, scalars = _device_inputs(torch_cuda, name, o, 181, n, "uniform")
    hb, hs = _to_np(bases), _to_np(scalars)                 # pageable
    heads = np.stack([o.generate_points(181, 1, first=j * 4096)[0] for j in range(n // 4096)])
    want = o.msm_affine(heads, o.fold_chain_scalars(hs))
    pb, ps = bases.cpu().pin_memory(), scalars.cpu().pin_memory()
    with msm.MSMGpu(name) as ctx:
        assert (o.jacobian_to_affine(ctx.affine_msm(hb, hs)) == want).all()
        assert ctx.last_timing()["ranges"] > 1
        assert (o.jacobian_to_affine(ctx.affine_msm(pb.data_ptr(), ps.data_ptr(), n)) == want).all()
        # device bases (resident SRS), pageable scalars
        assert (o.jacobian_to_affine(ctx.affine_msm(bases.data_ptr(), hs, n)) == want).all()


# Witness-like (skewed) scalars at a size where the scalar sample switches the sort to the
# two-level shared-memory form: host scalars (sampled in place), one range and the automatic
# pipeline, against the chain-fold value; the forced one-level sort must agree.
def test_msm_skewed_large_host_scalars(oracles, torch_cuda):
    name, n = "bn254", (1 << 21) + 4096
    o = oracles[name]
    bases, scalars = _device_inputs(torch_cuda, name, o, 191, n, "witness")
    hs = _to_np(scalars)
    heads = np.stack([o.generate_points(191, 1, first=j * 4096)[0] for j in range(n // 4096)])
    want = o.msm_affine(heads, o.fold_chain_scalars(hs))
    ps = scalars.cpu().pin_memory()
    with msm.MSMGpu(name) as ctx:
        for ranges, sort_mode in ((1, -1), (0, -1), (1, 0)):
            ctx.set_option("ranges", ranges)
            ctx.set_option("sort_mode", sort_mode)
            got = o.jacobian_to_affine(ctx.affine_msm(bases.data_ptr(), ps.data_ptr(), n))
            assert (got == want).all(), (ranges, sort_mode)


# Threading contract of the reference (msm_gpu.h:26-33): one context per host thread; contexts
# are independent (own stream, workspace, copy stream) and may run concurrently.
def test_concurrent_contexts_on_host_threads(oracles, torch_cuda):
    import threading
    jobs = [("bn254", 5000, 301), ("bls12_381", 3000, 302), ("bn254", 1 << 15, 303), ("bn254_g2", 2000, 304)]
    inputs = [(oracles[c].generate_points(s, n), oracles[c].generate_scalars(s + 50, n)) for c, n, s in jobs]
    want = [oracles[c].msm_affine(b, k) for (c, n, s), (b, k) in zip(jobs, inputs)]
    got = [None] * len(jobs)

    def work(i):
        c = jobs[i][0]
        with msm.MSMGpu(c) as ctx:
            for _ in range(3):
                got[i] = oracles[c].jacobian_to_affine(ctx.affine_msm(*inputs[i]))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(len(jobs)):
        assert got[i] is not None and (got[i] == want[i]).all(), jobs[i]
