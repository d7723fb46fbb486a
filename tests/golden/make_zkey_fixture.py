"""Generates tests/golden/zkey_multiplier_3_g1.json from the reference tree.

Run HERE (the container that has /root/reference); the GPU box never needs it.
It records (a) the raw Montgomery bytes of the G1 header points alpha1, beta1,
delta1 of vendors/circom/examples/multiplier_3.zkey (the zkey stores affine
points as Montgomery u64 limbs, used zero-copy by the reference:
vendors/circom/circomlib/zkey/zkey.h:176-183) and (b) the decimal coordinates
that the reference's own unit test expects for them
(vendors/circom/circomlib/zkey/zkey_unittest.cc:66-101).  Together they pin
R = 2^256, the limb order and from-Montgomery of the oracle.
"""
import json
import os
import re
import shutil
import struct

REF = "/root/reference/vendors/circom"


def sections(buf):
    assert buf[:4] == b"zkey"
    _ver, nsec = struct.unpack_from("<II", buf, 4)
    off = 12
    out = {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", buf, off)
        off += 12
        out[typ] = buf[off:off + size]
        off += size
    return out


def copy_file_fixtures():
    """The two reference-held binary fixtures the file-level Groth16 tests run on (test-only
    copies: vendors/circom/examples/multiplier_3.zkey, circomlib/wtns/multiplier_3.wtns — the
    files zkey_unittest.cc:56-58 and wtns_unittest.cc:24-26 parse)."""
    here = os.path.dirname(os.path.abspath(__file__))
    shutil.copyfile(f"{REF}/examples/multiplier_3.zkey", os.path.join(here, "multiplier_3.zkey"))
    shutil.copyfile(f"{REF}/circomlib/wtns/multiplier_3.wtns", os.path.join(here, "multiplier_3.wtns"))


def main():
    copy_file_fixtures()
    buf = open(f"{REF}/examples/multiplier_3.zkey", "rb").read()
    hdr = sections(buf)[2]
    off = 0
    n8q, = struct.unpack_from("<I", hdr, off); off += 4
    q = int.from_bytes(hdr[off:off + n8q], "little"); off += n8q
    n8r, = struct.unpack_from("<I", hdr, off); off += 4
    r = int.from_bytes(hdr[off:off + n8r], "little"); off += n8r
    off += 12  # nVars, nPublic, domainSize
    pts = {}
    pts["alpha_g1"] = hdr[off:off + 2 * n8q].hex(); off += 2 * n8q
    pts["beta_g1"] = hdr[off:off + 2 * n8q].hex(); off += 2 * n8q
    pts["beta_g2"] = hdr[off:off + 4 * n8q].hex(); off += 4 * n8q
    pts["gamma_g2"] = hdr[off:off + 4 * n8q].hex(); off += 4 * n8q
    pts["delta_g1"] = hdr[off:off + 2 * n8q].hex(); off += 2 * n8q
    pts["delta_g2"] = hdr[off:off + 4 * n8q].hex(); off += 4 * n8q

    test_src = open(f"{REF}/circomlib/zkey/zkey_unittest.cc").read()
    expected = {}
    for name in ("alpha_g1", "beta_g1", "delta_g1"):
        m = re.search(name + r"_str\[2\]\s*=\s*\{\s*\"(\d+)\",\s*\"(\d+)\"", test_src)
        expected[name] = [m.group(1), m.group(2)]
    # G2 points: x = (c0, c1), y = (c0, c1), stored c0 first (zkey_unittest.cc:84-119)
    for name in ("beta_g2", "gamma_g2", "delta_g2"):
        m = re.search(name + r"_str\[2\]\[2\]\s*=\s*\{\s*\{\s*\"(\d+)\",\s*\"(\d+)\",?\s*\},\s*\{\s*\"(\d+)\",\s*\"(\d+)\"",
                      test_src)
        expected[name] = [m.group(1), m.group(2), m.group(3), m.group(4)]
    out = {
        "source": "vendors/circom/examples/multiplier_3.zkey + circomlib/zkey/zkey_unittest.cc:66-101",
        "q": str(q), "r": str(r), "n8q": n8q,
        "montgomery_bytes_hex": pts,
        "expected_decimal_xy": expected,
    }
    with open(__file__.rsplit("/", 1)[0] + "/zkey_multiplier_3_g1.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
