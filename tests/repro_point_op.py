import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import cpu_oracle
from tachyon_b200 import msm
name = sys.argv[1] if len(sys.argv) > 1 else "bn254"
op = sys.argv[2] if len(sys.argv) > 2 else "add"
o = cpu_oracle.CurveOracle(name)
n = 8
aff = o.generate_points(11, n)
ks = o.fr_from_mont(o.generate_scalars(12, n))
A = np.stack([o.scalar_mul(aff[i], ks[i]) for i in range(n)])
B = np.stack([o.scalar_mul(aff[(i * 7 + 3) % n], ks[(i + 1) % n]) for i in range(n)])
if op == "add_dbl":
    B = A.copy()
    op = "add"
if op == "add_zero":
    A[0] = o.xyzz_zero(); B[1] = o.xyzz_zero()
    op = "add"
if op == "add":
    got = msm.point_op(name, "add", A.reshape(n, -1), B.reshape(n, -1)).reshape(n, 4, -1)
    ok = all((o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_add(A[i], B[i]))).all() for i in range(n))
else:
    got = msm.point_op(name, "madd", A.reshape(n, -1), aff).reshape(n, 4, -1)
    ok = all((o.xyzz_to_affine(got[i]) == o.xyzz_to_affine(o.xyzz_madd(A[i], aff[i]))).all() for i in range(n))
print("OK" if ok else "MISMATCH")
