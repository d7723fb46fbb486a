"""Guards on the BUILT library (no GPU needed: cuobjdump reads the embedded SASS).

Round 2 found that the same kernel source came out of nvcc with 64 or 128 registers depending on
the translation unit and the build, halving the occupancy of the sort's
scatter kernels; and that an out-of-line point addition left a fifth of its products as un-fused
IMAD + IMAD.HI pairs.  These tests pin what the measured numbers rely on: resident CTAs per SM
of the throughput kernels (registers per thread), one copy of the sort kernels, fused wide
products in the hot loops."""
import collections
import re
import shutil
import subprocess

import pytest

from tachyon_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def resources():
    try:
        out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True,
                             text=True, timeout=300).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    res = collections.defaultdict(list)
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            res[name].append((int(m.group(1)), int(m.group(2))))
            name = None
    if not res:
        pytest.skip("no resource usage in the library")
    return res


def _find(res, *parts):
    hits = {k: v for k, v in res.items() if all(p in k for p in parts)}
    assert hits, f"no kernel matching {parts}"
    return hits


# (substrings of the mangled name, most registers per thread) — 65536 / (threads x resident CTAs)
CAPS = [
    (("coarse_scatter_kernel",), 64), (("fine_scatter_kernel",), 64), (("fine_hist_kernel",), 32),
    (("digits_coarse_hist_kernel",), 40), (("digits_scatter_kernel",), 32),
    (("17accumulate_kernelINS_10Bn254CurveELb0ELi4",), 128),
    (("accumulate_lockstep_kernelINS_11Bls381CurveELi4",), 128),
    (("17accumulate_kernelINS_11Bls381CurveELb0ELi3",), 168),
    (("accumulate_pair_kernelINS_12Bn254G2CurveELi128ELi3",), 168),
    (("accumulate_pair_kernelINS_13Bls381G2CurveELi128ELi2",), 255),
    (("reduce_blocks_kernelINS_10Bn254Curve", "Li0EEELb1"), 128),
    (("reduce_blocks_kernelINS_11Bls381Curve", "Li1EEELb0"), 128),
    (("reduce_blocks_pair_kernelINS_12Bn254G2CurveELi1ELi3",), 168),
    (("reduce_tree_kernelINS_10Bn254Curve",), 128),
]


@pytest.mark.parametrize("parts,cap", CAPS)
def test_register_budget(resources, parts, cap):
    for name, uses in _find(resources, *parts).items():
        for regs, _stack in uses:
            assert regs <= cap, f"{name}: {regs} registers, occupancy plan assumes <= {cap}"


def test_sort_kernels_defined_once(resources):
    # one definition (msm_sort_kernels.cu), not one static copy per curve
    for k in ("coarse_scatter_kernel", "fine_scatter_kernel", "fine_hist_kernel", "coarse_scan_kernel"):
        copies = sum(len(v) for v in _find(resources, k).values())
        assert copies == 1, f"{k}: {copies} copies in the library"


def _opcodes(function):
    out = subprocess.run([CUOBJDUMP, "-sass", "-fun", function, _lib.LIB_PATH], capture_output=True, text=True,
                         timeout=600).stdout
    ins = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    hist = collections.Counter()
    for line in out.splitlines():
        m = ins.match(line)
        if m:
            parts = m.group(1).split(".")
            hist[parts[0] + ("." + parts[1] if parts[0] == "IMAD" and len(parts) > 1 and parts[1] in ("WIDE", "HI") else "")] += 1
    return hist


@pytest.mark.parametrize("parts", [("17accumulate_kernelINS_10Bn254CurveELb0ELi4",),
                                   ("reduce_blocks_kernelINS_10Bn254Curve", "Li0EEELb1")])
def test_wide_products_are_fused(resources, parts):
    # every 32x32->64 product one IMAD.WIDE; the only IMAD.HI left are the first products of the
    # Montgomery reduction rows (low word known to be zero): 8 per 128 + 8 for 8 limbs = 6-10 %
    name = sorted(_find(resources, *parts))[0]
    hist = _opcodes(name)
    if not hist["IMAD.WIDE"]:
        pytest.skip("no SASS for " + name)
    assert hist["IMAD.HI"] <= 0.12 * hist["IMAD.WIDE"], dict(hist.most_common(6))
