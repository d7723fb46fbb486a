"""CPU-only checks of the product library: it loads, exports every symbol the
public header declares, and its host-only helpers agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import pymodel
from tachyon_b200 import _lib, msm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "tachyon_msm_b200.h")).read()
    # per-curve declarations live in the TACHYON_B200_DECLARE_CURVE macro
    macro_names = set(re.findall(r"(tachyon_##C##_[#\w]+)\s*\(", hdr))
    declared = {n.replace("##C##", c).replace("##G##", g) for n in macro_names for c in _lib.CURVES
                for g in (_lib.GROUPS if "##G##" in n else ("",))}
    declared |= set(re.findall(r"\b(tachyon_b200_\w+)\s*\(", hdr))
    assert declared == set(_lib.all_symbols())
    L = _lib.load()
    for name in declared:
        assert getattr(L, name) is not None
    # the reference's five entry points per curve (msm_gpu.h.tpl:26-56, point.h.tpl:117)
    for c in ("bn254", "bls12_381"):
        for n in ("g1_init", "g1_create_msm_gpu", "g1_destroy_msm_gpu", "g1_point2_msm_gpu",
                  "g1_affine_msm_gpu"):
            assert f"tachyon_{c}_{n}" in declared


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "tachyon_msm_b200.h"\n'
                   "_Static_assert(sizeof(struct tachyon_bn254_g1_affine) == 64, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bn254_g1_jacobian) == 96, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bn254_fr) == 32, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bls12_381_g1_affine) == 96, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bls12_381_g1_jacobian) == 144, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bls12_381_g1_xyzz) == 192, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bn254_g2_affine) == 128, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bn254_g2_jacobian) == 192, \"\");\n"
                   "_Static_assert(sizeof(struct tachyon_bls12_381_g2_xyzz) == 384, \"\");\n"
                   "int main(void) { return 0; }\n")
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_window_rule_host():
    # W * c >= bits + 1 so the signed top digit never carries out
    for bits in (254, 255):
        for n in (1, 5, 1 << 10, 1 << 16, 1 << 20, 1 << 24):
            c = msm.window_bits(n, bits)
            W = msm.window_count(bits, c)
            assert 4 <= c <= 22 and W <= 64 and W * c >= bits + 1 and (W - 1) * c < bits + 1
            assert W << (c - 1) <= 1 << 24
    assert msm.window_bits(1 << 24, 254) >= msm.window_bits(1 << 16, 254)


@pytest.mark.parametrize("name", ["bn254", "bls12_381", "bn254_g2", "bls12_381_g2"])
def test_host_point_helpers_vs_oracle(oracles, name):
    o = oracles[name]
    pts = o.generate_points(5, 6)
    ks = o.fr_from_mont(o.generate_scalars(6, 6))
    xs = [o.scalar_mul(pts[i], ks[i]) for i in range(6)]
    zero = o.xyzz_zero()
    for a, b in [(xs[0], xs[1]), (xs[2], xs[2]), (xs[3], zero), (zero, xs[4]), (zero, zero)]:
        got = msm.xyzz_add(name, a, b)
        want = o.xyzz_add(a, b)
        assert (o.xyzz_to_affine(got) == o.xyzz_to_affine(want)).all()
        assert (msm.xyzz_to_jacobian(name, got) == o.xyzz_to_jacobian(got)).all()
    # P + (-P) = identity
    neg = xs[5].copy()
    neg[1] = o.fq_op("neg", xs[5][1:2])[0]
    assert (o.xyzz_to_affine(msm.xyzz_add(name, xs[5], neg)) == 0).all()


@pytest.mark.parametrize("name", ["bn254", "bls12_381", "bn254_g2", "bls12_381_g2"])
def test_batch_normalize_vs_oracle(oracles, name):
    # point_xyzz.h:109-163 BatchNormalize: one inversion for the whole batch, identity -> (0, 0)
    o = oracles[name]
    pts = o.generate_points(15, 9)
    ks = o.fr_from_mont(o.generate_scalars(16, 9))
    xs = np.stack([o.scalar_mul(pts[i], ks[i]) for i in range(9)])
    xs[0] = o.xyzz_zero()
    xs[4] = o.xyzz_zero()
    xs[8] = o.xyzz_zero()
    got = msm.batch_normalize(name, xs)
    for i in range(9):
        assert (got[i] == o.xyzz_to_affine(xs[i]).reshape(-1)).all(), i
    assert msm.batch_normalize(name, xs[:0]).shape[0] == 0
    assert (msm.batch_normalize(name, xs[:1]) == 0).all()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        msm.MSMGpu("bn254")
    with pytest.raises(RuntimeError):
        msm.field_op("bn254", "fq", "mul", np.zeros((1, 4), dtype=np.uint64))


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU,
    times FULL-size MSMs and prints the same `config` as the GPU arm would."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--log-n", "12",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, root)
    import bench
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["config"] == bench.workload_config("bn254", 12, "uniform")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "full 2^12-point MSM per step" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


def test_window_model_is_sane():
    """The exported window rule: defined for every size, never more than 2^24 bucket slots, larger
    windows for larger inputs (up to the occupancy dip between 2^13 and 2^15 points)."""
    from tachyon_b200 import msm
    prev = 0
    for lg in range(4, 27):
        c = msm.window_bits(1 << lg, 254)
        w = msm.window_count(254, c)
        assert 4 <= c <= 22 and w * (1 << (c - 1)) <= (1 << 24) and w * c >= 255
        if lg >= 15:
            assert c >= prev
        prev = c
    assert msm.window_bits(1 << 24, 254) == 20 and msm.window_bits(1 << 21, 254) == 17
