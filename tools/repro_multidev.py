"""Debug aid: in-process multi-device paths (batch dealing and point-range sharding)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import msm
from oracle import cpu_oracle
print("devices", msm.device_count(), [torch.cuda.get_device_name(i) for i in range(torch.cuda.device_count())])
name = "bn254"
o = cpu_oracle.CurveOracle(name)
m, n_pub = 1 << 12, 10
full = o.generate_scalars(131, m, "witness")
hco = o.generate_scalars(132, m, "uniform")
sizes = [m, m, m - n_pub, m]
bases = [o.generate_points(140 + j, n) for j, n in enumerate(sizes)]
scal = [full, full, full[n_pub:].copy(), hco]
want = [np.asarray(o.msm_affine(b, s)).reshape(-1) for b, s in zip(bases, scal)]
for low in (-1, 0):
    for k in (1, 2, min(4, msm.device_count())):
        if k > msm.device_count():
            continue
        with msm.MSMGpu(name) as ctx:
            ctx.set_option("low_windows", low)
            ctx.set_option("devices", k)
            for rep in range(3):
                got = msm.batch_normalize(name, ctx.msm_batch(bases, scal))
                print("batch low", low, "devices", k, "rep", rep, [bool((got[j] == want[j]).all()) for j in range(4)])
            for rep in range(2):
                jac = ctx.affine_msm(bases[3], scal[3])
                print("  sharded single MSM ok:", bool((np.asarray(o.jacobian_to_affine(jac)).reshape(-1) == want[3]).all()))
