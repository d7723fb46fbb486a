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
    issued_per_add = None
    if not curve.endswith("_g2"):
        L = limbs
        issued_per_add = 6 * (2 * L * L + L) + 2 * (L * (L + 1) // 2 + L * L + L) + (3 * L * L + L)
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
        self.sm, self.power, self.bits = [], [], 0
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


def run_reference(args, rank, world):
    """--impl reference: the CPU path (oracle restatement of Tachyon's OpenMP
    Pippenger, VariableBaseMSM::Run) on all host threads; a step is an MSM over a
    bounded prefix of the same workload."""
    if rank != 0:
        return
    from oracle import cpu_oracle
    cpu_oracle.build()
    o = cpu_oracle.CurveOracle(args.curve)
    sample_log = min(args.log_n, args.cpu_sample_log)
    n = 1 << sample_log
    threads = os.cpu_count() or cpu_oracle.max_threads()  # torchrun pins OMP_NUM_THREADS=1
    bases = o.generate_points(SEED + 2, n)
    scalars = o.generate_scalars(SEED + 3, n, args.dist)
    for _ in range(max(1, min(args.warmup, 1))):
        o.msm(bases, scalars, threads=threads)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        o.msm(bases, scalars, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    value = n / dt
    sample = f"2^{sample_log}-point prefix of the 2^{args.log_n} workload, {steps} timed MSMs"
    print(json.dumps({
        "impl": "reference", "metric": f"{args.curve} G1 MSM throughput", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"{args.curve} G1 MSM 2^{args.log_n} points, {args.dist} scalars", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_groth16(args, rank, world, local_rank):
    """--workload groth16 (BASELINE.json configs[4]): the G1 MSM set of one Groth16 proof of a
    synthetic 2^log_m-constraint circuit — A, B1 (full assignment), L (witness part) and H
    (quotient coefficients), zk/r1cs/groth16/prove.h:100-131 — dealt over the ranks
    (tachyon_b200.sharding.deal_msms) and run through the batched C-ABI call.  The proving
    key (bases) is device-resident as in a prover that keeps its zkey loaded; scalars are
    device-resident for `value` and pinned host memory for `e2e`."""
    import torch
    import torch.distributed as dist
    from tachyon_b200 import msm, sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    curve, fq = args.curve, FQ_LIMBS[args.curve]
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), wall)   # host epilogues of a batch overlap device work: wall bounds it
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, out

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
    ms_step, result = timed(lambda: step(scalars), args.steps)
    launches = msm.kernel_launch_count() - l0
    clocks = sampler.stop()
    step(h_scalars)
    e2e_ms, e2e_out = timed(lambda: step(h_scalars), max(2, min(args.steps, 5)))
    # ---- the whole proof (adds the G2 MSM B2 and the r / s arithmetic), one process -------------
    proof = None
    if world == 1 and not curve.endswith("_g2"):
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
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            run_proof()
        proof = {"ms": (time.perf_counter() - t0) * 1e3 / reps,
                 "what": "tachyon_%s_groth16_prove_b200: L, H, A, B1 (G1 batch) + B2 (G2, concurrent) + blinding, "
                         "resident proving key, pinned host assignments, wall clock" % curve}
        g2ctx.close()
        del b2, a_q, b1_q

    parity = "skipped"
    if not args.no_parity and rank == 0:
        from oracle import cpu_oracle
        o = cpu_oracle.CurveOracle(curve)
        ok = True
        for j, n in enumerate(sizes):
            hs = scalars[j].cpu().numpy().view(np.uint64)
            pad = (-n) % 4096
            if pad:
                hs = np.concatenate([hs, np.zeros((pad, 4), dtype=np.uint64)])
            heads = np.stack([o.generate_points(SEED + 40 + j, 1, first=c * 4096)[0] for c in range(len(hs) // 4096)])
            want = o.msm_affine(heads, o.fold_chain_scalars(hs))
            ok = ok and bool((o.xyzz_to_affine(result[j]) == want).all()) and bool((o.xyzz_to_affine(e2e_out[j]) == want).all())
        parity = "bit-exact vs CPU oracle (chain-fold), all four MSMs" if ok else "MISMATCH"
    if world > 1:
        dist.barrier()
    if rank == 0:
        total = sum(sizes)
        alg = sum(algorithmic_products(curve, n)["products"] for n in sizes)
        peak = max(msm.imad_peak(local_rank, v, 3) for v in (0, 1, 2))
        print(json.dumps({
            "metric": f"{curve} Groth16 G1 MSM set throughput", "value": total / (ms_step * 1e-3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"Groth16 G1 MSM set (A, B1, L, H) of a synthetic 2^{args.log_m}-constraint circuit, "
                                   f"{curve}; witness-like assignment (40% 0, 30% 1, 20% <2^32, 10% full), uniform h",
                       "sizes": dict(zip(names, sizes)), "split": args.groth16_split,
                       "work_rank0": [[names[j], lo, hi] for j, lo, hi in work],
                       "l2": "inputs + workspace exceed the 126 MB L2 every step"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 32 * total, "d2h_bytes_per_step": 0,
                    "api": f"tachyon_{curve}_g1_msm_gpu_batch_b200: resident proving-key bases, pinned host scalars"},
            "roofline": {"bound": "int32-imad", "achieved": alg / (ms_step * 1e-3) / 1e9, "peak": peak * world / 1e9,
                         "unit": "G products/s (32x32->64)", "frac": alg / (ms_step * 1e-3) / (peak * world),
                         "traffic": None, "note": "W_alg of SURVEY 8d summed over the four MSMs; witness scalars are "
                                                  "mostly 0/1 so far fewer additions are actually needed"},
            "cpu_baseline": None, "parity": parity, "proof": proof}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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
    ctx.register_bases(bases.data_ptr(), n)

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
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary 2^20 measurement")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--ranges", type=int, default=0, help="point ranges per MSM (0 = automatic)")
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16", "commit_batch"])
    ap.add_argument("--batch", type=int, default=16, help="commit_batch: number of MSMs per step")
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

    import torch
    import torch.distributed as dist
    from tachyon_b200 import msm, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    curve, fq = args.curve, FQ_LIMBS[args.curve]
    n_total = 1 << args.log_n
    lo, hi = sharding.shard_range(n_total, rank, world)
    n_local = hi - lo

    def make_inputs(first, count, seed_shift=0):
        b = torch.empty((count, 2 * fq), dtype=torch.int64, device="cuda")
        s = torch.empty((count, 4), dtype=torch.int64, device="cuda")
        msm.generate_bases_device(curve, SEED + 2 + seed_shift, count, b.data_ptr(), first=first)
        msm.generate_scalars_device(curve, SEED + 3 + seed_shift, count, s.data_ptr(), args.dist, first=first)
        torch.cuda.synchronize()
        return b, s

    bases, scalars = make_inputs(lo, n_local)
    stream = torch.cuda.Stream()
    ctx = msm.MSMGpu(curve, degree=args.log_n, device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    if args.window_bits:
        ctx.set_option("window_bits", args.window_bits)
    if args.ranges:
        ctx.set_option("ranges", args.ranges)

    gather = sharding.PartialGather((4, fq), world) if world > 1 else None

    def step(b_ptr, s_ptr, count):
        """One whole MSM: local partial on this GPU, then (N > 1) gather + host add on rank 0."""
        part = ctx.msm_xyzz(b_ptr, s_ptr, count)
        if world == 1:
            return part
        parts = gather(part, stream)
        if rank != 0:
            return part
        return sharding.combine_partials(curve, list(parts))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the engine's stream; max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = t.tolist()
        return ms / steps, wall_ms / steps, out

    # ---- device-resident measurement ------------------------------------------------
    resident = lambda: step(bases.data_ptr(), scalars.data_ptr(), n_local)
    for _ in range(args.warmup):
        result = resident()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = ""
    sampler = ClockSampler(gpu_uuid, local_rank)
    sampler.start()
    launches0 = msm.kernel_launch_count()
    stage = {"sort_ms": 0.0, "accumulate_ms": 0.0, "reduce_ms": 0.0, "host_ms": 0.0, "total_ms": 0.0}

    def resident_with_stages():
        r = resident()
        t = ctx.last_timing()
        for k in stage:
            stage[k] += t[k]
        return r

    ms_step, wall_step, result = timed(resident_with_stages, args.steps)
    launches = (msm.kernel_launch_count() - launches0)
    clocks = sampler.stop()
    timing = ctx.last_timing()
    for k in stage:
        stage[k] /= args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- end to end through the reference-shaped C-ABI call, pinned host buffers ----
    h_bases = torch.empty((n_local, 2 * fq), dtype=torch.int64).pin_memory()
    h_scalars = torch.empty((n_local, 4), dtype=torch.int64).pin_memory()
    h_bases.copy_(bases)
    h_scalars.copy_(scalars)
    torch.cuda.synchronize()

    def e2e_step():
        if world == 1:
            return ctx.affine_msm(h_bases.data_ptr(), h_scalars.data_ptr(), n_local)  # tachyon_*_g1_affine_msm_gpu
        return step(h_bases.data_ptr(), h_scalars.data_ptr(), n_local)

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    e2e_ms, e2e_wall, e2e_out = timed(e2e_step, e2e_steps)
    e2e_timing = ctx.last_timing()
    h2d = n_local * (2 * fq * 8 + 32) * world
    d2h = (2 * e2e_timing["windows"] * 4 * fq * 8 + 16) * world
    e2e_value = n_total / (max(e2e_ms, e2e_wall) * 1e-3)

    # ---- parity: chain-fold property against the CPU oracle (outside timing) --------
    parity = "skipped"
    if not args.no_parity:
        from oracle import cpu_oracle
        o = cpu_oracle.CurveOracle(curve)
        hs = h_scalars.numpy().view(np.uint64)
        hb = h_bases.numpy().view(np.uint64)
        folded = o.fold_chain_scalars(hs)
        want_local = o.msm(hb[::4096], folded)        # this rank's partial sum, CPU
        mine = ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), n_local)
        ok = bool((o.xyzz_to_affine(mine) == o.xyzz_to_affine(want_local)).all())
        if world > 1:
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
            if rank == 0:
                # and the combined point equals the sum of the CPU partials
                gathered = [None] * world
                dist.gather_object(want_local, gathered, dst=0)
                tot = gathered[0]
                for g in range(1, world):
                    tot = o.xyzz_add(tot, gathered[g])
                ok = ok and bool((o.xyzz_to_affine(result) == o.xyzz_to_affine(tot)).all())
            else:
                dist.gather_object(want_local, None, dst=0)
        else:
            ok = ok and bool((o.xyzz_to_affine(result) == o.xyzz_to_affine(want_local)).all())
            jac = np.asarray(e2e_out)
            ok = ok and bool((o.jacobian_to_affine(jac) == o.xyzz_to_affine(want_local)).all())
        parity = "bit-exact vs CPU oracle (chain-fold)" if ok else "MISMATCH"

    # ---- IMAD roofline ----------------------------------------------------------------
    alg = algorithmic_products(curve, n_total)
    peak = max(msm.imad_peak(local_rank, v, 3) for v in (0, 1, 2)) if rank == 0 else 0.0

    extra = None
    cpu_baseline = None
    if rank == 0:
        # secondary: 2^20 on this one GPU (BASELINE.json configs[1])
        if not args.no_extra and world == 1 and args.log_n != 20:
            nb, ns = make_inputs(0, 1 << 20, seed_shift=10)
            f20 = lambda: ctx.msm_xyzz(nb.data_ptr(), ns.data_ptr(), 1 << 20)
            for _ in range(3):
                f20()
            ms20, _, _ = timed(f20, max(args.steps, 10))
            a20 = algorithmic_products(curve, 1 << 20)
            extra = {"workload": f"{curve} G1 MSM 2^20 points, 1 GPU, device-resident", "ms_per_msm": ms20,
                     "points_per_s": (1 << 20) / (ms20 * 1e-3),
                     "imad_frac": a20["products"] / (ms20 * 1e-3) / peak if peak else None}
            del nb, ns
        if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only
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
                                      f"(kParallelTerm), mean of {reps}"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        acc_ms = stage["accumulate_ms"]
        # DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get(f"{curve}:{args.log_n}:{world}", {})
        except OSError:
            pass
        line = {
            "metric": f"{curve} G1 MSM throughput", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{curve} G1 MSM 2^{args.log_n} points, {args.dist} scalars",
                       "partition": f"point range over {world} GPU(s), {n_local} points per GPU",
                       "arithmetic": "254/381-bit Montgomery field elements as u32 limbs (IMAD.WIDE carry chains)",
                       "window_bits": timing["window_bits"], "windows": timing["windows"],
                       "l2": "inputs + workspace exceed the 126 MB L2 every step"},
            "wall_ms_per_step": wall_step,
            "stages_ms": stage,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": max(e2e_ms, e2e_wall),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ranges": e2e_timing["ranges"], "h2d_ms": e2e_timing["h2d_ms"],
                    "api": "tachyon_%s_%s_affine_msm_gpu, pinned host buffers; H2D of point range k+1 "
                           "overlaps sort/accumulate of range k" % (curve.replace("_g2", ""),
                                                                    "g2" if curve.endswith("_g2") else "g1")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "int32-imad", "kernel": "accumulate_kernel",
                         "achieved": alg["accumulate_products"] / world / (acc_ms * 1e-3) / 1e9 if acc_ms else None,
                         "peak": peak / 1e9, "unit": "G products/s (32x32->64)",
                         "frac": (alg["accumulate_products"] / world / (acc_ms * 1e-3) / peak) if acc_ms and peak else None,
                         "traffic": traffic.get("bytes"), "traffic_source": traffic.get("source"),
                         "peak_source": "measured live: tachyon_b200_imad_peak, best of IMAD.WIDE.X chains / IMAD.WIDE acc64 / IMAD+IMAD.HI",
                         "whole_msm_frac": alg["products"] / (ms_step * 1e-3) / (peak * world) if peak else None,
                         # pipe occupancy: wide products the kernel actually issued (entries the
                         # engine reports x issued products per mixed addition) / time / peak
                         "issued_frac": (timing["entries"] * alg["issued_per_add"] / (acc_ms * 1e-3) / peak)
                         if acc_ms and peak and alg["issued_per_add"] else None,
                         "algorithmic": {"c": alg["c"], "W": alg["W"], "products": alg["products"]}},
            "hbm": {"algorithmic_bytes": n_total * (2 * fq * 8 + 32),
                    "achieved_gbs": n_total * (2 * fq * 8 + 32) / (ms_step * 1e-3) / 1e9},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "extra_2p20": extra,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
