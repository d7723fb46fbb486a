#!/usr/bin/env python3
"""Benchmark of the hot path: BN254 (or BLS12-381) G1 variable-base MSM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--curve bn254|bls12_381] [--log-n 24] [--dist uniform|non_uniform|witness]

One "step" = one whole MSM over synthetic inputs (SURVEY.md §8d generators).
At N > 1 (launched by torchrun, one rank per GPU) the points are partitioned by
contiguous range; every rank computes a partial sum on its GPU and the N
partials (128 B each) are gathered with one NCCL all_gather and added on the
host — strong scaling, the total stays 2^log_n points.

Rank 0 prints ONE JSON line.  `value` is whole-job points/s with inputs resident
in HBM; `e2e` is the same through the reference-shaped C-ABI call with pinned
HOST buffers (H2D inside the timed region).  `--impl reference` times the CPU
oracle (restatement of Tachyon's OpenMP Pippenger; the reference itself cannot
be built in this image) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x7461636879
SCALAR_BITS = {"bn254": 254, "bls12_381": 255, "bn254_g2": 254, "bls12_381_g2": 255}
FQ_LIMBS = {"bn254": 4, "bls12_381": 6, "bn254_g2": 8, "bls12_381_g2": 12}   # u64 limbs of one coordinate
BASE_LIMBS = {"bn254": 4, "bls12_381": 6, "bn254_g2": 4, "bls12_381_g2": 6}  # u64 limbs of Fq


def algorithmic_products(curve, n):
    """W_alg of SURVEY.md §8(d): 32x32->64 products of the reference's own window
    rule c = max(ceil(log2 n) - 4, 1) (icicle_msm_utils.cc:26-28)."""
    lam = SCALAR_BITS[curve]
    limbs = 2 * BASE_LIMBS[curve]
    c = max((n - 1).bit_length() - 4, 1)
    W = -(-lam // c)
    modmuls = n * W * 10 + W * (1 << c) * 14 + W * c * 9 + W * 14
    per_mul = 2 * limbs * limbs + limbs
    if curve.endswith("_g2"):
        per_mul *= 3        # one Fq2 multiplication = 3 Fq multiplications (Karatsuba count)
    # What the accumulation kernel ISSUES per mixed addition (xyzz.cuh xyzz_madd: 6 mul + 2 sqr
    # + 1 fused two-product mul2; fp.cuh: mul 2L^2+L, sqr L(L+1)/2+L^2+L, mul2 3L^2+L wide
    # products) — fewer than the canonical 10 multiplications, which is why `frac` can exceed 1.
    L = limbs
    if not curve.endswith("_g2"):
        issued_per_add = 6 * (2 * L * L + L) + 2 * (L * (L + 1) // 2 + L * L + L) + (3 * L * L + L)
    else:
        # G2 (fp.cuh Fp2Field / Fp2Lanes): an Fq2 multiplication is two fp_mul2 (3L^2+L each), an Fq2
        # squaring two fp_mul (2L^2+L each), the fused Y3 two Fq2 multiplications
        issued_per_add = 6 * 2 * (3 * L * L + L) + 2 * 2 * (2 * L * L + L) + 2 * 2 * (3 * L * L + L)
    return dict(c=c, W=W, modmuls=modmuls, products=modmuls * per_mul,
                accumulate_products=n * W * 10 * per_mul, issued_per_add=issued_per_add)


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region: NVML polled
    every 5 ms from a thread (an MSM step is tens of ms, too short for `nvidia-smi -lms`);
    falls back to one nvidia-smi query if NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, uuid, gpu_index):
        self.uuid, self.gpu = uuid, gpu_index
        self.sm, self.mem, self.power, self.bits = [], [], [], 0
        self.ecc = None
        self.max_sm = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode() if self.uuid else b"")
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            try:  # what kind of box this is: the memory-bound stages vary from box to box
                self.ecc = bool(pynvml.nvmlDeviceGetEccMode(h)[0])
            except Exception:
                self.ecc = None
            self.nvml, self.handle = pynvml, h
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        n, h = self.nvml, self.handle
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mem.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_MEM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
                self.bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self.stop_flag.set()
        self.thread.join(timeout=1)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
                "mem_mhz": statistics.median(self.mem) if self.mem else None, "ecc": self.ecc,
                "reasons": [name for name, bit in self.REASONS if self.bits & bit], "source": "nvml, 5 ms poll"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            r = [x.strip() for x in out.strip().split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "power_w_max": float(r[2]), "samples": 1,
                    "reasons": [n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")],
                    "source": "nvidia-smi, one query after the timed region"}
        except Exception as e:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable: %s" % e], "samples": 0}


def workload_config(curve, log_n, dist):
    """`config` of the JSON line: identical for this repo's arm and for --impl reference (the
    driver compares them); engine-specific detail goes under the top-level key "engine"."""
    group = "G2" if curve.endswith("_g2") else "G1"
    return {"workload": f"{curve} {group} MSM 2^{log_n} points, {dist} scalars",
            "l2": "inputs + workspace exceed the 126 MB L2 every step"}


def run_reference(args, rank, world):
    """--impl reference: the CPU path (oracle restatement of Tachyon's OpenMP Pippenger,
    VariableBaseMSM::Run -> PippengerAdapter kParallelTerm) on all host threads, on the SAME
    workload as this repo's arm: every timed step is one full 2^log_n-point MSM
    (benchmark/msm/msm_benchmark_gpu.cc:56-66 runs CPU and GPU on the same sizes).  Warm-up steps
    run on a 2^18-point prefix.  If K full MSMs would take more than --cpu-budget seconds the step
    count is cut (and reported)."""
    if rank != 0:
        return
    from oracle import cpu_oracle
    cpu_oracle.build()
    o = cpu_oracle.CurveOracle(args.curve)
    n = 1 << args.log_n
    threads = os.cpu_count() or cpu_oracle.max_threads()  # torchrun pins OMP_NUM_THREADS=1
    bases = o.generate_points(SEED + 2, n)
    scalars = o.generate_scalars(SEED + 3, n, args.dist)
    wn = min(n, 1 << 18)
    for _ in range(max(1, args.warmup)):
        o.msm(bases[:wn], scalars[:wn], threads=threads)
    steps = max(1, args.steps)
    times = []
    t_begin = time.perf_counter()
    for k in range(steps):
        t0 = time.perf_counter()
        o.msm(bases, scalars, threads=threads)
        times.append(time.perf_counter() - t0)
        if k + 1 < steps and (time.perf_counter() - t_begin) + times[-1] > args.cpu_budget:
            break
    steps_done = len(times)
    dt = sum(times) / steps_done
    value = n / dt
    sample = (f"full 2^{args.log_n}-point MSM per step, {steps_done} timed steps"
              + ("" if steps_done == steps else f" (cut from {steps}: --cpu-budget {args.cpu_budget:.0f} s)")
              + f"; {max(1, args.warmup)} warm-up MSMs on a 2^18-point prefix")
    print(json.dumps({
        "impl": "reference", "metric": f"{args.curve} G1 MSM throughput", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": steps_done, "warmup": max(1, args.warmup), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": workload_config(args.curve, args.log_n, args.dist),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample,
                         "ms_per_msm": dt * 1e3, "ms_min": min(times) * 1e3, "ms_max": max(times) * 1e3},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def groth16_measure(args, d, steps, with_proof=False, curve="bn254"):
    """The G1 MSM set of one Groth16 proof of a synthetic 2^log_m-constraint circuit
    (BASELINE.json configs[4]) — A, B1 (full assignment), L (witness part) and H (quotient
    coefficients), zk/r1cs/groth16/prove.h:100-131 — dealt over the ranks
    (tachyon_b200.sharding.deal_msms) and run through the batched C-ABI call.  The proving
    key (bases) is device-resident as in a prover that keeps its zkey loaded; scalars are
    device-resident for `ms_per_step` and pinned host memory for `e2e_ms_per_step`."""
    import torch
    from tachyon_b200 import msm, sharding
    rank, world, local_rank = d.rank, d.world, d.local_rank
    fq = FQ_LIMBS[curve]
    m = 1 << args.log_m
    n_pub = 64
    names = ["A", "B1", "L", "H"]
    sizes = [m, m, m - n_pub, m]
    bases, scalars = [], []
    full = torch.empty((m, 4), dtype=torch.int64, device="cuda")          # full assignment, witness-like
    msm.generate_scalars_device(curve, SEED + 30, m, full.data_ptr(), "witness")
    hco = torch.empty((m, 4), dtype=torch.int64, device="cuda")           # h coefficients, uniform
    msm.generate_scalars_device(curve, SEED + 31, m, hco.data_ptr(), "uniform")
    for j, n in enumerate(sizes):
        b = torch.empty((n + 4096, 2 * fq), dtype=torch.int64, device="cuda")[:n]
        msm.generate_bases_device(curve, SEED + 40 + j, n, b.data_ptr())
        bases.append(b)
    scalars = [full, full, full[n_pub:], hco]
    torch.cuda.synchronize()
    h_scalars = [t.cpu().pin_memory() for t in (full, hco)]
    h_scalars = [h_scalars[0], h_scalars[0], h_scalars[0][n_pub:], h_scalars[1]]

    work = sharding.deal_msms(sizes, world, args.groth16_split)[rank]
    stream = torch.cuda.Stream()
    ctx = msm.MSMGpu(curve, degree=args.log_m, device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    zero = np.zeros((4, fq), dtype=np.uint64)
    set_gather = sharding.PartialGather((len(sizes), 4, fq), world) if world > 1 else None

    def step(sc):
        parts = np.stack([zero] * len(sizes))
        if work:
            out = ctx.msm_batch([bases[j].data_ptr() + lo * 2 * fq * 8 for j, lo, hi in work],
                                [sc[j].data_ptr() + lo * 32 for j, lo, hi in work], [hi - lo for j, lo, hi in work])
            for (j, lo, hi), o in zip(work, out):
                parts[j] = o
        if world == 1:
            return parts
        g = set_gather(parts, stream)
        return sharding.combine_set(curve, g) if rank == 0 else parts

    def timed(fn, nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d.barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(nsteps):
            out = fn()
        e1.record(stream)
        d.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), wall)   # host epilogues of a batch overlap device work: wall bounds it
        ms = d.max_float([ms])[0]
        return ms / nsteps, out

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(scalars)
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = ""
    sampler = ClockSampler(gpu_uuid, local_rank)
    sampler.start()
    l0 = msm.kernel_launch_count()
    ms_step, result = timed(lambda: step(scalars), steps)
    launches = msm.kernel_launch_count() - l0
    clocks = sampler.stop()
    step(h_scalars)
    e2e_ms, e2e_out = timed(lambda: step(h_scalars), max(2, min(steps, 5)))
    # ---- the whole proof (adds the G2 MSM B2 and the r / s arithmetic), one process -------------
    proof = None
    if with_proof and world == 1 and not curve.endswith("_g2"):
        g2fq = 2 * fq
        b2 = torch.empty((m + 1, 2 * g2fq), dtype=torch.int64, device="cuda")
        msm.generate_bases_device(curve + "_g2", SEED + 50, m + 1, b2.data_ptr())
        a_q = torch.empty((m + 1, 2 * fq), dtype=torch.int64, device="cuda")
        b1_q = torch.empty((m + 1, 2 * fq), dtype=torch.int64, device="cuda")
        msm.generate_bases_device(curve, SEED + 51, m + 1, a_q.data_ptr())
        msm.generate_bases_device(curve, SEED + 52, m + 1, b1_q.data_ptr())
        torch.cuda.synchronize()
        single = lambda t: t[:1].cpu().numpy().view(np.uint64).reshape(-1)
        pk = {"alpha_g1": single(bases[0]), "beta_g1": single(bases[1]), "delta_g1": single(bases[2]),
              "beta_g2": single(b2), "delta_g2": b2[1:2].cpu().numpy().view(np.uint64).reshape(-1),
              "a_g1_query": (a_q.data_ptr(), m + 1), "b_g1_query": (b1_q.data_ptr(), m + 1),
              "b_g2_query": (b2.data_ptr(), m + 1), "h_g1_query": (bases[3].data_ptr(), m),
              "l_g1_query": (bases[2].data_ptr(), m - n_pub)}
        rs = torch.empty((2, 4), dtype=torch.int64, device="cuda")
        msm.generate_scalars_device(curve, SEED + 53, 2, rs.data_ptr(), "uniform")
        rs = rs.cpu().numpy().view(np.uint64)
        hw, hh, hf = h_scalars[2].numpy().view(np.uint64), h_scalars[3].numpy().view(np.uint64), \
            h_scalars[0].numpy().view(np.uint64)
        g2ctx = msm.MSMGpu(curve + "_g2", degree=args.log_m, device=local_rank)
        run_proof = lambda: msm.groth16_prove(ctx, g2ctx, pk, rs[0], rs[1], hh, hw, hf)
        for _ in range(3):
            run_proof()
        t0 = time.perf_counter()
        reps = max(3, min(steps, 10))
        for _ in range(reps):
            run_proof()
        proof = {"ms": (time.perf_counter() - t0) * 1e3 / reps,
                 "what": "tachyon_%s_groth16_prove_b200: L, H, A, B1 (G1 batch) + B2 (G2, concurrent) + blinding, "
                         "resident proving key, pinned host assignments, wall clock" % curve}
        g2ctx.close()
        del b2, a_q, b1_q

    parity = "skipped (--no-parity)"
    if not args.no_parity:
        ok = True
        if rank == 0:
            from oracle import cpu_oracle
            o = cpu_oracle.CurveOracle(curve)
            for j, n in enumerate(sizes):
                hs = scalars[j].cpu().numpy().view(np.uint64)
                pad = (-n) % 4096
                if pad:
                    hs = np.concatenate([hs, np.zeros((pad, 4), dtype=np.uint64)])
                heads = np.stack([o.generate_points(SEED + 40 + j, 1, first=c * 4096)[0] for c in range(len(hs) // 4096)])
                want = o.msm_affine(heads, o.fold_chain_scalars(hs))
                ok = ok and bool((o.xyzz_to_affine(result[j]) == want).all()) and bool((o.xyzz_to_affine(e2e_out[j]) == want).all())
        parity = "bit-exact vs CPU oracle (chain-fold), all four MSMs" if ok else "MISMATCH"
    d.barrier()
    ctx.close()
    total = sum(sizes)
    return {"workload": f"Groth16 G1 MSM set (A, B1, L, H) of a synthetic 2^{args.log_m}-constraint circuit, "
                        f"{curve}; witness-like assignment (40% 0, 30% 1, 20% <2^32, 10% full), uniform h; {world} GPU(s)",
            "sizes": dict(zip(names, sizes)), "split": args.groth16_split,
            "work_rank0": [[names[j], lo, hi] for j, lo, hi in work],
            "ms_per_step": ms_step, "points_per_s": total / (ms_step * 1e-3), "e2e_ms_per_step": e2e_ms,
            "total_points": total, "launches": int(launches), "clocks": clocks, "warmup": warm,
            "parity": parity, "proof": proof}


def run_groth16(args, rank, world, local_rank):
    """--workload groth16: the set above as the headline of its own JSON line."""
    from tachyon_b200 import msm
    d = Dist(rank, world, local_rank)
    curve = args.curve
    g = groth16_measure(args, d, args.steps, with_proof=True, curve=curve)
    if rank == 0:
        sizes = list(g["sizes"].values())
        total, ms_step, e2e_ms = g["total_points"], g["ms_per_step"], g["e2e_ms_per_step"]
        alg = sum(algorithmic_products(curve, n)["products"] for n in sizes)
        peak = max(msm.imad_peak(local_rank, v, 3) for v in (0, 1, 2))
        print(json.dumps({
            "metric": f"{curve} Groth16 G1 MSM set throughput", "value": total / (ms_step * 1e-3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": g["warmup"], "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": g["workload"], "sizes": g["sizes"], "split": g["split"],
                       "work_rank0": g["work_rank0"],
                       "l2": "inputs + workspace exceed the 126 MB L2 every step"},
            "clocks": g["clocks"], "gpu_launches": g["launches"],
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 32 * total, "d2h_bytes_per_step": 0,
                    "api": f"tachyon_{curve}_g1_msm_gpu_batch_b200: resident proving-key bases, pinned host scalars"},
            "roofline": {"bound": "int32-imad", "achieved": alg / (ms_step * 1e-3) / 1e9, "peak": peak * world / 1e9,
                         "unit": "G products/s (32x32->64)", "frac": alg / (ms_step * 1e-3) / (peak * world),
                         "traffic": None, "note": "W_alg of SURVEY 8d summed over the four MSMs; witness scalars are "
                                                  "mostly 0/1 so far fewer additions are actually needed"},
            "cpu_baseline": None, "parity": g["parity"], "proof": g["proof"]}), flush=True)
    if world > 1:
        d.dist.destroy_process_group()


def run_commit_batch(args, local_rank):
    """--workload commit_batch: the KZG / SHPlonk commitment loop of
    tachyon/crypto/commitments/kzg/kzg.h:91-113, 217-313 — the SRS is registered once, then
    `--batch` MSMs of 2^log_n points with fresh scalars run through
    tachyon_<c>_g1_msm_gpu_commit_batch_b200 and are batch-normalised.  `value`: scalars resident
    in HBM; `e2e`: scalars in pinned host memory (their H2D is pipelined against the previous
    MSM).  One GPU."""
    import torch
    from tachyon_b200 import msm
    torch.cuda.set_device(local_rank)
    curve, fq = args.curve, FQ_LIMBS[args.curve]
    n, count = 1 << args.log_n, args.batch
    bases = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
    msm.generate_bases_device(curve, SEED + 60, n, bases.data_ptr())
    dev, host = [], []
    for i in range(count):
        t = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        msm.generate_scalars_device(curve, SEED + 61 + i, n, t.data_ptr(), args.dist)
        dev.append(t)
        host.append(t.cpu().pin_memory())
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    ctx = msm.MSMGpu(curve, degree=args.log_n, device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    if args.precompute:
        ctx.set_option("precompute", 1)
    t_reg = time.perf_counter()
    ctx.register_bases(bases.data_ptr(), n)
    torch.cuda.synchronize()
    t_reg = (time.perf_counter() - t_reg) * 1e3

    def timed(ptrs, steps):
        for _ in range(3):
            out = ctx.commit_batch(ptrs, [n] * count)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = ctx.commit_batch(ptrs, [n] * count)
            aff = msm.batch_normalize(curve, out)
        return (time.perf_counter() - t0) * 1e3 / steps, aff

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = ""
    sampler = ClockSampler(gpu_uuid, local_rank)
    sampler.start()
    l0 = msm.kernel_launch_count()
    ms_dev, aff_dev = timed([t.data_ptr() for t in dev], args.steps)
    launches = msm.kernel_launch_count() - l0
    clocks = sampler.stop()
    t = ctx.last_timing()     # sums over the commitments of the last batch
    stages = {k: t[k] / count for k in ("sort_ms", "accumulate_ms", "reduce_ms", "host_ms")}
    ms_host, aff_host = timed([t.data_ptr() for t in host], max(2, min(args.steps, 5)))
    parity = "skipped"
    if not args.no_parity:
        from oracle import cpu_oracle
        o = cpu_oracle.CurveOracle(curve)
        heads = np.stack([o.generate_points(SEED + 60, 1, first=c * 4096)[0] for c in range(n // 4096)])
        ok = bool((aff_dev == aff_host).all())
        for i in (0, count - 1):
            want = o.msm_affine(heads, o.fold_chain_scalars(host[i].numpy().view(np.uint64)))
            ok = ok and bool((aff_dev[i] == np.asarray(want).reshape(-1)).all())
        parity = "bit-exact vs CPU oracle (chain-fold), first and last commitment; host == resident" if ok else "MISMATCH"
    alg = algorithmic_products(curve, n)["products"] * count
    peak = max(msm.imad_peak(local_rank, v, 3) for v in (0, 1, 2))
    print(json.dumps({
        "metric": f"{curve} G1 commitment batch throughput", "value": n * count / (ms_dev * 1e-3), "unit": "points/s",
        "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"{count} commitments of 2^{args.log_n} points over registered (device-resident) "
                               f"{curve} bases, {args.dist} scalars, batch-normalised", "ms_per_commitment": ms_dev / count,
                   "precompute": bool(args.precompute), "register_ms": t_reg, "stages_ms_per_commitment": stages,
                   "window_bits": ctx.last_timing()["window_bits"], "windows": ctx.last_timing()["windows"],
                   "l2": "inputs + workspace exceed the 126 MB L2 every step"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": n * count / (ms_host * 1e-3), "unit": "points/s", "ms_per_step": ms_host,
                "h2d_bytes_per_step": 32 * n * count, "d2h_bytes_per_step": 0,
                "api": f"tachyon_{curve}_g1_msm_gpu_commit_batch_b200, pinned host scalars"},
        "roofline": {"bound": "int32-imad", "achieved": alg / (ms_dev * 1e-3) / 1e9, "peak": peak / 1e9,
                     "unit": "G products/s (32x32->64)", "frac": alg / (ms_dev * 1e-3) / peak, "traffic": None},
        "cpu_baseline": None, "parity": parity}), flush=True)
    ctx.close()


def run_table(args):
    """`-k K [-k K ...] [--test_set random|non_uniform] [--check_results]`: the protocol and flags
    of the reference's own benchmark (benchmark/msm/msm_config.cc:39-59, msm_runner.h:46-61,
    msm_benchmark_gpu.cc:47-70): ONE test set of the largest size, prefixes of it for the
    smaller sizes, ONE un-warmed wall-clock sample per size, host (pageable) pointers through
    the C API, seconds per MSM in a table beside the CPU implementation."""
    import torch
    from oracle import cpu_oracle
    from tachyon_b200 import msm
    curve, fq = args.curve, FQ_LIMBS[args.curve]
    ks = sorted(args.k)
    nmax = 1 << ks[-1]
    dist = {"random": "uniform", "uniform": "uniform", "non_uniform": "non_uniform"}[args.test_set]
    torch.cuda.set_device(0)
    b = torch.empty((nmax, 2 * fq), dtype=torch.int64, device="cuda")
    sc = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
    msm.generate_bases_device(curve, SEED + 2, nmax, b.data_ptr())
    msm.generate_scalars_device(curve, SEED + 3, nmax, sc.data_ptr(), dist)
    torch.cuda.synchronize()
    hb, hs = b.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64)   # pageable, like std::vector
    del b, sc
    o = cpu_oracle.CurveOracle(curve)
    threads = os.cpu_count() or cpu_oracle.max_threads()
    ctx = msm.MSMGpu(curve, degree=ks[-1], device=0, banner=True)
    # all GPU samples first, then the CPU column (msm_benchmark_gpu.cc:52-66 also runs one
    # implementation over every size before the next; it also keeps the OpenMP team of the
    # CPU run from competing with this library's copy threads on a small host)
    gpu = {}
    for k in ks:
        n = 1 << k
        t0 = time.perf_counter()
        jac = ctx.affine_msm(hb[:n], hs[:n], n)
        gpu[k] = (time.perf_counter() - t0, jac)
    rows = []
    for k in ks:
        n = 1 << k
        cpu_s, cpu_pt = None, None
        if k <= args.cpu_sample_log:
            t0 = time.perf_counter()
            cpu_pt = o.msm_affine(hb[:n], hs[:n], threads=threads)
            cpu_s = time.perf_counter() - t0
        gpu_s, jac = gpu[k]
        ok = None
        if args.check_results:
            got = o.jacobian_to_affine(jac)
            if cpu_pt is None:   # too large for the CPU column: chain-fold value instead
                pad = (-n) % 4096
                fs = hs[:n] if not pad else np.concatenate([hs[:n], np.zeros((pad, 4), dtype=np.uint64)])
                cpu_pt = o.msm_affine(hb[:n:4096], o.fold_chain_scalars(fs))
            ok = bool((got == cpu_pt).all())
            if not ok:
                raise SystemExit(f"--check_results: GPU result differs from the CPU result at k={k}")
        rows.append((k, cpu_s, gpu_s, ok))
    print(f"| Exponent | Tachyon CPU restatement ({threads} threads) | B200 (this library, first call un-warmed) |")
    print("| :------: | ------------ | ------------ |")
    for k, c, g, ok in rows:
        print(f"|    {k}    | {'-' if c is None else '%.6f' % c} | **{g:.6f}** |")
    for k, c, g, ok in rows:
        print(json.dumps({"k": k, "curve": curve, "test_set": args.test_set, "cpu_s": c, "gpu_s": g,
                          "check_results": ok, "protocol": "one un-warmed sample, pageable host pointers"}), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381", "bn254_g2", "bls12_381_g2"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "non_uniform", "witness"])
    ap.add_argument("--cpu-sample-log", type=int, default=20)
    ap.add_argument("--cpu-budget", type=float, default=240.0,
                    help="--impl reference: seconds of timed full-size CPU MSMs after which the step count is cut")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary 2^20 measurement")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--ranges", type=int, default=0, help="point ranges per MSM (0 = automatic)")
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16", "commit_batch"])
    ap.add_argument("--batch", type=int, default=16, help="commit_batch: number of MSMs per step")
    ap.add_argument("--precompute", action="store_true",
                    help="commit_batch: register the bases with the table of window multiples")
    ap.add_argument("-k", type=int, action="append", default=None,
                    help="table mode with the reference benchmark's flags: exponent(s) of the sizes")
    ap.add_argument("--test_set", default="random", choices=["random", "uniform", "non_uniform"])
    ap.add_argument("--check_results", action="store_true")
    ap.add_argument("--vendor", action="append", default=None, help="accepted and ignored (no third-party MSMs here)")
    ap.add_argument("--log-m", type=int, default=20, help="groth16: log2 of the constraint count")
    ap.add_argument("--groth16-split", default="auto", choices=["auto", "msm", "range"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner with printf on
        # fd 1 (NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr for native code and Python's
        # own stdout keeps a duplicate of the real one
        sys.stdout.flush()
        real = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = real
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.k:
        if rank == 0:
            run_table(args)
        return
    if args.workload == "groth16":
        run_groth16(args, rank, world, local_rank)
        return
    if args.workload == "commit_batch":
        if rank == 0:
            if args.log_n == 24:
                args.log_n = 20
            run_commit_batch(args, local_rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    args.warmup = max(args.warmup, 3)
    run_msm(args, rank, world, local_rank)


class Dist:
    """Process-group plumbing of one bench run: NCCL for device tensors, a gloo group for
    host-side barriers and object gathers (ranks parked on gloo leave their GPU idle)."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.local_rank, self.torch, self.dist = rank, world, local_rank, torch, dist
        torch.cuda.set_device(local_rank)
        self.host = None
        if world > 1:
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.host = dist.new_group(backend="gloo")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.host)

    def max_float(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def all_ok(self, ok):
        if self.world == 1:
            return ok
        t = self.torch.tensor([1 if ok else 0], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def gather_objects(self, obj):
        """list over ranks on rank 0, None elsewhere (gloo)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0, group=self.host)
        return out

    def broadcast_object(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0, group=self.host)
        return box[0]


class MsmCase:
    """One MSM workload on this rank: its point range resident in HBM (and in pinned host
    memory for the end-to-end leg), a context that — at N > 1 — has joined the ranks, so that
    every call returns the whole MSM: the only exchange, one ncclAllGather of the partial sums,
    is issued by the library on the context's stream (tachyon_*_msm_gpu_join_ranks_b200)."""

    def __init__(self, d, curve, log_n, dist_name, seed_shift=0, window_bits=0, ranges=0):
        from tachyon_b200 import msm, sharding
        torch = d.torch
        self.d, self.msm, self.curve, self.log_n, self.dist_name = d, msm, curve, log_n, dist_name
        self.fq = FQ_LIMBS[curve]
        self.n_total = 1 << log_n
        lo, hi = sharding.shard_range(self.n_total, d.rank, d.world)
        self.lo, self.n_local = lo, hi - lo
        self.bases = torch.empty((self.n_local, 2 * self.fq), dtype=torch.int64, device="cuda")
        self.scalars = torch.empty((self.n_local, 4), dtype=torch.int64, device="cuda")
        msm.generate_bases_device(curve, SEED + 2 + seed_shift, self.n_local, self.bases.data_ptr(), first=lo)
        msm.generate_scalars_device(curve, SEED + 3 + seed_shift, self.n_local, self.scalars.data_ptr(), dist_name,
                                    first=lo)
        torch.cuda.synchronize()
        self.stream = torch.cuda.Stream()
        self.ctx = msm.MSMGpu(curve, degree=log_n, device=d.local_rank)
        self.ctx.set_stream(self.stream.cuda_stream)
        if window_bits:
            self.ctx.set_option("window_bits", window_bits)
        if ranges:
            self.ctx.set_option("ranges", ranges)
        if d.world > 1:
            uid = d.broadcast_object(msm.nccl_unique_id() if d.rank == 0 else None)
            self.ctx.join_ranks(uid, d.rank, d.world)
        self.h_bases = self.h_scalars = None
        self.host_memory = None

    def close(self):
        self.ctx.close()
        for b in (self.h_bases, self.h_scalars):
            if b is not None:
                b.free()
        self.bases = self.scalars = self.h_bases = self.h_scalars = None

    def resident(self):
        return self.ctx.msm_xyzz(self.bases.data_ptr(), self.scalars.data_ptr(), self.n_local)

    def pin_host(self):
        """The rank's inputs in page-locked HOST memory.  With several ranks on one box the pages
        are write-combined (tachyon_b200_alloc_host): the bare copy probe (tools/probe/h2d_probe.cu,
        profiles/r2_h2d_probe_8gpu.txt) shows ordinary pinned memory dropping to ~30 GB/s per GPU
        when 4-8 GPUs copy at once, write-combined memory holding 39-49."""
        torch = self.d.torch
        if self.h_bases is None:
            wc = self.d.world >= 4
            self.h_bases = self.msm.HostBuffer(self.n_local, 2 * self.fq, write_combined=wc)
            self.h_scalars = self.msm.HostBuffer(self.n_local, 4, write_combined=wc)
            torch.from_numpy(self.h_bases.array.view(np.int64)).copy_(self.bases)
            torch.from_numpy(self.h_scalars.array.view(np.int64)).copy_(self.scalars)
            torch.cuda.synchronize()
            self.host_memory = "page-locked, write-combined" if wc else "page-locked"

    def e2e(self):
        """The reference-shaped call (tachyon_<c>_g1_affine_msm_gpu) with pinned HOST buffers."""
        return self.ctx.affine_msm(self.h_bases.ptr, self.h_scalars.ptr, self.n_local)

    def timed(self, fn, steps, stage=None):
        """K steps bracketed by barrier + synchronize, CUDA events on the engine's stream; max over ranks."""
        torch = self.d.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.d.barrier()
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            out = fn()
            if stage is not None:
                t = self.ctx.last_timing()
                for k in stage:
                    stage[k] += t[k]
        e1.record(self.stream)
        self.d.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms, wall_ms = self.d.max_float([e0.elapsed_time(e1), wall_ms])
        if stage is not None:
            for k in stage:
                stage[k] /= steps
        return ms / steps, wall_ms / steps, out

    def parity(self, results_xyzz=(), results_jacobian=()):
        """Chain-fold property against the CPU oracle (outside timing): the synthetic bases are
        chains P(j, d) = 2^d H_j, so MSM(P, s) = MSM(H, fold(s)) with a 4096x smaller MSM the
        oracle does in milliseconds.  Every rank checks its own partial; rank 0 checks the
        combined point(s) against the sum of the CPU partials."""
        from oracle import cpu_oracle
        o = cpu_oracle.CurveOracle(self.curve)
        hs = self.scalars.cpu().numpy().view(np.uint64)
        pad = (-self.n_local) % 4096
        if pad:
            hs = np.concatenate([hs, np.zeros((pad, 4), dtype=np.uint64)])
        heads = self.bases[::4096].cpu().numpy().view(np.uint64)
        want_local = o.msm(heads, o.fold_chain_scalars(hs))
        # this rank's own partial, through a context that has NOT joined the ranks
        with self.msm.MSMGpu(self.curve, device=self.d.local_rank) as solo:
            mine = solo.msm_xyzz(self.bases.data_ptr(), self.scalars.data_ptr(), self.n_local)
        ok = bool((o.xyzz_to_affine(mine) == o.xyzz_to_affine(want_local)).all())
        ok = self.d.all_ok(ok)
        gathered = self.d.gather_objects(want_local)
        if self.d.rank == 0:
            tot = gathered[0]
            for g in range(1, self.d.world):
                tot = o.xyzz_add(tot, gathered[g])
            want = o.xyzz_to_affine(tot)
            for r in results_xyzz:
                ok = ok and bool((o.xyzz_to_affine(np.asarray(r)) == want).all())
            for r in results_jacobian:
                ok = ok and bool((o.jacobian_to_affine(np.asarray(r)) == want).all())
        return "bit-exact vs CPU oracle (chain-fold)" if ok else "MISMATCH"


def measure_extra(d, curve, log_n, dist_name, steps, peak, seed_shift, e2e=False):
    """A secondary configuration measured like the headline (device-resident, strong scaling
    over the ranks) with its own parity check; returns the `extra_*` block of the JSON line."""
    case = MsmCase(d, curve, log_n, dist_name, seed_shift=seed_shift)
    for _ in range(3):
        case.resident()
    stage = {"sort_ms": 0.0, "accumulate_ms": 0.0, "reduce_ms": 0.0, "total_ms": 0.0}
    ms, wall, result = case.timed(case.resident, steps, stage)
    t = case.ctx.last_timing()
    out = {"workload": f"{curve} MSM 2^{log_n} points, {dist_name} scalars, {d.world} GPU(s), device-resident",
           "ms_per_msm": ms, "points_per_s": (1 << log_n) / (ms * 1e-3), "window_bits": t["window_bits"],
           "windows": t["windows"], "low_windows": t["low_windows"], "stages_ms": stage}
    alg = algorithmic_products(curve, 1 << log_n)
    if peak:
        out["imad_frac"] = alg["products"] / (ms * 1e-3) / (peak * d.world)
    jac = []
    if e2e:
        case.pin_host()
        case.e2e()
        e_ms, e_wall, e_out = case.timed(case.e2e, max(2, min(steps, 5)))
        out["e2e_ms_per_msm"] = max(e_ms, e_wall)
        jac = [e_out]
    out["parity"] = case.parity([result], jac)
    case.close()
    return out


def measure_in_process_devices(d, curve, log_n, dist_name, steps):
    """The product's own multi-GPU path, the one a drop-in caller of the reference's C API gets
    (TACHYON_B200_MSM_DEVICES / set_option("devices", k), msm_api_common.cuh MsmGpuContext::Run):
    ONE process, ONE call of tachyon_<c>_g1_affine_msm_gpu with host buffers, fanned out to k
    engines on k GPUs from k host threads, partials added on the host.  Run by rank 0 after the
    timed regions while the other ranks wait on a host-side (gloo) barrier."""
    out = None
    if d.rank == 0:
        import torch
        from oracle import cpu_oracle
        from tachyon_b200 import msm
        fq, n = FQ_LIMBS[curve], 1 << log_n
        k = min(d.world, msm.device_count())
        b = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
        sc = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        msm.generate_bases_device(curve, SEED + 2, n, b.data_ptr())
        msm.generate_scalars_device(curve, SEED + 3, n, sc.data_ptr(), dist_name)
        torch.cuda.synchronize()
        wc = k >= 4
        hbuf, sbuf = msm.HostBuffer(n, 2 * fq, write_combined=wc), msm.HostBuffer(n, 4, write_combined=wc)
        torch.from_numpy(hbuf.array.view(np.int64)).copy_(b)
        torch.from_numpy(sbuf.array.view(np.int64)).copy_(sc)
        heads = b[::4096].cpu().numpy().view(np.uint64)
        folded_in = sc.cpu().numpy().view(np.uint64)      # ordinary memory for the CPU-side check
        del b, sc
        torch.cuda.synchronize()
        ctx = msm.MSMGpu(curve, degree=log_n, device=d.local_rank)
        ctx.set_option("devices", k)
        for _ in range(3):
            jac = ctx.affine_msm(hbuf.ptr, sbuf.ptr, n)
        t0 = time.perf_counter()
        for _ in range(steps):
            jac = ctx.affine_msm(hbuf.ptr, sbuf.ptr, n)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        t = ctx.last_timing()
        o = cpu_oracle.CurveOracle(curve)
        want = o.msm_affine(heads, o.fold_chain_scalars(folded_in))
        ok = bool((o.jacobian_to_affine(jac) == want).all())
        ctx.close()
        hbuf.free()
        sbuf.free()
        out = {"workload": f"{curve} MSM 2^{log_n} points through ONE tachyon_{curve}_g1_affine_msm_gpu call, "
                           f"option devices = {k}, page-locked{' write-combined' if wc else ''} host buffers "
                           "(H2D inside the timed region)",
               "devices": t["devices"], "ms_per_msm": ms, "points_per_s": n / (ms * 1e-3), "timing": "host wall clock",
               "parity": "bit-exact vs CPU oracle (chain-fold)" if ok else "MISMATCH"}
    d.host_barrier()
    return out


def run_msm(args, rank, world, local_rank):
    import torch
    from tachyon_b200 import msm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    d = Dist(rank, world, local_rank)
    curve, fq = args.curve, FQ_LIMBS[args.curve]
    n_total = 1 << args.log_n
    case = MsmCase(d, curve, args.log_n, args.dist, window_bits=args.window_bits, ranges=args.ranges)
    ctx, n_local = case.ctx, case.n_local

    # ---- device-resident measurement ------------------------------------------------
    for _ in range(args.warmup):
        result = case.resident()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = ""
    sampler = ClockSampler(gpu_uuid, local_rank)
    sampler.start()
    launches0 = msm.kernel_launch_count()
    stage = {"sort_ms": 0.0, "accumulate_ms": 0.0, "reduce_ms": 0.0, "host_ms": 0.0, "total_ms": 0.0,
             "acc_kernel_ms": 0.0, "combine_ms": 0.0}
    ms_step, wall_step, result = case.timed(case.resident, args.steps, stage)
    launches = (msm.kernel_launch_count() - launches0)
    clocks = sampler.stop()
    timing = ctx.last_timing()
    value = n_total / (ms_step * 1e-3)

    # ---- end to end through the reference-shaped C-ABI call, pinned host buffers ----
    case.pin_host()
    e2e_steps = max(2, min(args.steps, 5))
    case.e2e()
    e2e_ms, e2e_wall, e2e_out = case.timed(case.e2e, e2e_steps)
    e2e_timing = ctx.last_timing()
    h2d = n_local * (2 * fq * 8 + 32) * world
    d2h = (4 * fq * 8 * (world if world > 1 else 1) + 40 * e2e_timing["ranges"]) * world
    e2e_value = n_total / (max(e2e_ms, e2e_wall) * 1e-3)

    parity = "skipped (--no-parity)" if args.no_parity else case.parity([result], [e2e_out])

    # ---- IMAD roofline ----------------------------------------------------------------
    alg = algorithmic_products(curve, n_total)
    peaks = [msm.imad_peak(local_rank, v, 3) for v in (0, 1, 2)] if rank == 0 else [0.0]
    peak = d.broadcast_object(max(peaks))
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 0.0

    extras = {}
    cpu_baseline = None
    if not args.no_extra:
        ex_steps = max(5, min(args.steps, 10))
        if args.log_n != 20 and world == 1:
            extras["extra_2p20"] = measure_extra(d, curve, 20, args.dist, max(args.steps, 10), peak, 10, e2e=True)
        if world == 1:
            extras["extra_2p16"] = measure_extra(d, curve, 16, args.dist, max(args.steps, 20), peak, 11)
        if curve == "bn254" and args.log_n == 24:
            # BASELINE.json configs[3]: BLS12-381 G1 2^22 on 1 / N GPUs (the 381-bit limb path)
            extras["extra_bls12_381_2p22"] = measure_extra(d, "bls12_381", 22, args.dist, ex_steps, peak, 20,
                                                           e2e=(world == 1))
            # BASELINE.json configs[4]: the Groth16 G1 MSM set, dealt over the ranks
            g = groth16_measure(args, d, ex_steps)
            extras["extra_groth16"] = g
        if world > 1:
            extras["extra_in_process_devices"] = measure_in_process_devices(d, curve, args.log_n, args.dist,
                                                                            max(3, min(args.steps, 5)))
    if rank == 0 and not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only
        from oracle import cpu_oracle
        o = cpu_oracle.CurveOracle(curve)
        threads = os.cpu_count() or cpu_oracle.max_threads()  # torchrun pins OMP_NUM_THREADS=1
        sl = min(args.log_n, args.cpu_sample_log)
        cb = o.generate_points(SEED + 2, 1 << sl)
        cs = o.generate_scalars(SEED + 3, 1 << sl, args.dist)
        o.msm(cb, cs, threads=threads)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            o.msm(cb, cs, threads=threads)
        cdt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": (1 << sl) / cdt, "unit": "points/s", "cores": threads, "kind": "port",
                        "ms_per_msm": cdt * 1e3,
                        "sample": f"2^{sl}-point prefix of the workload, oracle OpenMP Pippenger "
                                  f"(kParallelTerm), mean of {reps}; `bench.py --impl reference` times the full size"}

    d.barrier()
    if rank == 0:
        acc_ms, acc_entries = stage["acc_kernel_ms"], timing["acc_kernel_entries"]
        # DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get(f"{curve}:{args.log_n}:{world}", {})
        except OSError:
            pass
        issued = alg["issued_per_add"]
        canonical = 10 * (2 * (2 * BASE_LIMBS[curve]) ** 2 + 2 * BASE_LIMBS[curve]) * (3 if curve.endswith("_g2") else 1)
        rate = lambda per_add: acc_entries * per_add / (acc_ms * 1e-3) if acc_ms and per_add else None
        theoretical = sm_count * 32 * sm_mhz * 1e6
        line = {
            "metric": f"{curve} G1 MSM throughput", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": workload_config(curve, args.log_n, args.dist),
            "engine": {"partition": f"point range over {world} GPU(s), {n_local} points per GPU; partial sums "
                                    "exchanged by one ncclAllGather issued by the library on its stream" if world > 1
                       else f"one GPU, {n_local} points",
                       "arithmetic": "254/381-bit Montgomery field elements as u32 limbs (IMAD.WIDE carry chains)",
                       "window_bits": timing["window_bits"], "windows": timing["windows"],
                       "low_windows": timing["low_windows"]},
            "wall_ms_per_step": wall_step,
            "stages_ms": stage,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": max(e2e_ms, e2e_wall),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ranges": e2e_timing["ranges"], "h2d_ms": e2e_timing["h2d_ms"],
                    "h2d_gbs_per_gpu": (h2d / world) / (e2e_timing["h2d_ms"] * 1e-3) / 1e9 if e2e_timing["h2d_ms"] else None,
                    "host_memory": case.host_memory,
                    "api": "tachyon_%s_%s_affine_msm_gpu, pinned host buffers; H2D of point range k+1 "
                           "overlaps sort/accumulate of range k" % (curve.replace("_g2", ""),
                                                                    "g2" if curve.endswith("_g2") else "g1")},
            "gpu_launches": int(launches),
            # frac = wide products the dominant kernel ISSUED (mixed additions the engine reports x
            # 32x32->64 products per addition in xyzz_madd) per second of that kernel's own launch
            # time, over the measured pipe peak.  frac_alg uses the canonical 10 multiplications of
            # 2 L^2 + L products per addition instead and can exceed 1.
            "roofline": {"bound": "int32-imad",
                         "kernel": "accumulate_pair_kernel" if curve.endswith("_g2") else
                                   ("accumulate_lockstep_kernel" if curve == "bls12_381" and n_local * timing["windows"] >= (12 << 20)
                                    else "accumulate_kernel"),
                         "achieved": rate(issued) / 1e9 if rate(issued) else None,
                         "peak": peak / 1e9, "unit": "G products/s (32x32->64)",
                         "frac": rate(issued) / peak if rate(issued) and peak else None,
                         "frac_alg": rate(canonical) / peak if rate(canonical) and peak else None,
                         "issued_products_per_madd": issued, "canonical_products_per_madd": canonical,
                         "kernel_ms": acc_ms, "kernel_madds": acc_entries,
                         "traffic": traffic.get("bytes"), "traffic_source": traffic.get("source"),
                         "peak_source": "measured live: tachyon_b200_imad_peak, best of IMAD.WIDE.X chains / "
                                        "IMAD.WIDE acc64 / IMAD+IMAD.HI",
                         "peak_variants": [p / 1e9 for p in peaks],
                         "peak_theoretical": theoretical / 1e9,
                         "peak_theoretical_source": f"{sm_count} SMs x 32 lanes/clk (half-rate wide IMAD) x {sm_mhz:.0f} MHz",
                         "whole_msm_frac": alg["products"] / (ms_step * 1e-3) / (peak * world) if peak else None,
                         "algorithmic": {"c": alg["c"], "W": alg["W"], "products": alg["products"]}},
            "hbm": {"algorithmic_bytes": n_total * (2 * fq * 8 + 32),
                    "achieved_gbs": n_total * (2 * fq * 8 + 32) / (ms_step * 1e-3) / 1e9},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    case.close()
    if world > 1:
        d.dist.destroy_process_group()


if __name__ == "__main__":
    main()
