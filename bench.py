#!/usr/bin/env python
"""bench.py — AES-128 block witnesses/s (Fr advice + lookup helper columns) on B200.

One "step" = one whole halo2 circuit of the reference's bench shape (K=20, N=5 column sets,
filled to its capacity of 3853 AES blocks under one key; BASELINE config 4's per-GPU unit):
    witness   : key schedule + 3853 x encrypt -> 16 advice columns (512 MiB)
    permuted  : the 25 lookups' A', S' columns (1.56 GiB)      [theta]
    product   : the 25 lookups' grand products Z (0.78 GiB)   [beta, gamma]
With --gpus N every rank runs its own independent circuits (no data-path collective; weak
scaling).  Output: ONE JSON line on rank 0 (see the contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x68326165  # "h2ae"
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
K_BENCH, N_BENCH = 20, 5
BLOCKS_BENCH = 3853  # capacity of the shape: 769 + 4 * 771 (src/aes128.rs:303-325)
METRIC = "AES-128 block witnesses/sec (Fr advice + lookup columns A',S',Z)"
UNIT = "blocks/s"


def splitmix_bytes(seed: int, n: int) -> np.ndarray:
    idx = np.arange((n + 7) // 8, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)[:n].copy()


def seeded_fr(seed: int) -> np.ndarray:
    v = int.from_bytes(splitmix_bytes(seed, 64).tobytes(), "little") % R_MOD
    m = v * ((1 << 256) % R_MOD) % R_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def measured_peak():
    """HBM GB/s denominator: MEASURED_PEAKS.json if present, else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_circuit(oracle, seed: int, threads: int, n_lookups: int = -1):
    """ONE circuit of the benchmarked shape on the CPU oracle: K=20, N=5, 3853 blocks under one
    key.  synthesize runs on one thread (the reference's synthesize is sequential,
    src/main.rs:42-54); the 25 lookup arguments (A, S, A', S', Z each) are independent and run
    under OpenMP on `threads` threads.  Returns (synthesize s, lookups s)."""
    th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
    nb = BLOCKS_BENCH
    key = splitmix_bytes(seed, 16)
    pts = splitmix_bytes(seed + 1, 16 * nb).reshape(nb, 16)
    ts, tl, _ = oracle.bench_circuit(oracle.KIND_AES, K_BENCH, N_BENCH, key, pts, th, be, ga,
                                     threads=threads, n_lookups=n_lookups)
    return ts, tl


def cpu_one_thread(oracle):
    """the single-thread figure.  Measured: synthesize of the whole circuit and the
    five lookups of ONE column set (range, xor, sbox, mul2, mul3) on one thread; the lookups of the
    other four sets are the same work on other columns, so their time is taken as 5x (stated)."""
    ts, tl = cpu_circuit(oracle, SEED + 40, 1, n_lookups=5)
    total = ts + tl * (N_BENCH * 5) / 5
    return {"value": BLOCKS_BENCH / total, "unit": UNIT, "cores": 1, "kind": "port",
            "synthesize_s": ts, "lookups_of_one_column_set_s": tl,
            "sample": f"one K={K_BENCH},N={N_BENCH} circuit of {BLOCKS_BENCH} blocks on ONE thread: "
                      f"synthesize {ts:.2f} s measured, the 5 lookups of one column set {tl:.2f} s "
                      f"measured and counted {N_BENCH}x for the {N_BENCH * 5} lookups"}


def cpu_baseline_sample(threads: int, circuits: int = 2):
    """The oracle port on the host cores, on the SAME workload as the GPU arm: `circuits` whole
    K=20,N=5 circuits of 3853 blocks (bounded sample: ~10-20 s)."""
    import oracle
    oracle.build()
    t0 = time.perf_counter()
    parts = [cpu_circuit(oracle, SEED + 20 + 2 * i, threads) for i in range(circuits)]
    sec = time.perf_counter() - t0
    blocks = BLOCKS_BENCH * circuits
    return {"value": blocks / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "seconds": sec, "synthesize_s": sum(p[0] for p in parts), "lookups_s": sum(p[1] for p in parts),
            "one_thread": cpu_one_thread(oracle),
            "prover_side": cpu_prover_side_sample(oracle, threads),
            "sample": f"{circuits} whole K={K_BENCH},N={N_BENCH} circuits of {BLOCKS_BENCH} blocks (the GPU "
                      f"arm's unit): synthesize on one thread as in the reference, the 25 lookups "
                      f"(A,S,A',S',Z) under OpenMP on {threads} threads; CPU restatement (oracle/), "
                      f"{blocks} blocks in {sec:.2f} s"}


def cpu_prover_side_sample(oracle, threads: int):
    """The oracle's restatements of the prover-side steps timed on the host
    cores on ONE column of the bench shape's length, beside the GPU numbers of `prover_side`:
    lagrange_to_coeff and coeff_to_extended of a 2^20 column (butterflies under OpenMP), one
    2^20-point multi-scalar multiplication (bucket method, windows under OpenMP)."""
    k = K_BENCH
    n = 1 << k
    col = splitmix_bytes(SEED + 300, n * 32).view(np.uint64).reshape(n, 4).copy()
    col[:, 3] %= np.uint64(0x30644E72E131A029)
    t0 = time.perf_counter()
    coeff = oracle.domain_transform(oracle.LAGRANGE_TO_COEFF, k, col)
    t1 = time.perf_counter()
    oracle.domain_transform(oracle.COEFF_TO_EXTENDED, k, coeff)
    t2 = time.perf_counter()
    bases = oracle.g1_test_bases(n, oracle.fr_from_int(0x1D2C3B4A5968778695A4B3C2D1E0F))
    t3 = time.perf_counter()
    oracle.g1_msm(col, bases)
    t4 = time.perf_counter()
    return {"cores": threads, "kind": "port", "sample": f"one column of 2^{k} uniform field elements",
            "lagrange_to_coeff_ms_per_column": 1e3 * (t1 - t0),
            "coeff_to_extended_ms_per_column": 1e3 * (t2 - t1),
            "msm_ms_per_column": 1e3 * (t4 - t3)}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port: the Rust crate cannot be
    compiled here — no cargo, un-vendored halo2_proofs) on all host threads, on the workload the
    GPU arm runs: whole K=20, N=5 circuits of 3853 blocks.  One circuit per step (a bounded
    sample of the GPU arm's step, which is `circuits_per_step_per_gpu` such circuits)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    for i in range(args.warmup):
        cpu_circuit(oracle, SEED + 100 + 2 * i, threads)
    t0 = time.perf_counter()
    ts = tl = 0.0
    for i in range(args.steps):
        a, b = cpu_circuit(oracle, SEED + 100 + 2 * (args.warmup + i), threads)
        ts += a
        tl += b
    total = time.perf_counter() - t0
    value = BLOCKS_BENCH * args.steps / total
    sample = (f"each step = ONE whole K={K_BENCH},N={N_BENCH} circuit of {BLOCKS_BENCH} blocks (1/{args.circuits} of "
              f"the GPU arm's step): synthesize on one thread as in the reference ({ts / args.steps:.2f} s), the "
              f"25 lookups A,S,A',S',Z under OpenMP on {threads} threads ({tl / args.steps:.2f} s); CPU "
              f"restatement of the reference (oracle/), the Rust crate itself is unbuildable here")
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 -> bn256 Fr (4x u64 Montgomery)", "data": "synthetic",
        "config": workload_config(args.gpus, args.circuits),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "cpu_baseline_1t": cpu_one_thread(oracle),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus, circuits_per_step):
    return {"workload": "BASELINE config 4 unit: independent K=20,N=5 circuits filled to capacity "
                        "(3853 AES-128 blocks under one key each): 16 advice columns + A',S',Z of "
                        "all 25 lookups; random keys/plaintexts (SplitMix64 seed 0x68326165), "
                        "seeded theta/beta/gamma",
            "k": K_BENCH, "column_sets": N_BENCH, "blocks_per_circuit": BLOCKS_BENCH,
            "circuits_per_step_per_gpu": circuits_per_step,
            "partitioning": f"independent circuits x{n_gpus}",
            "l2_policy": "outputs are 2.86 GiB per circuit (>> 126 MB L2) and never re-read"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class Pipe:
    """One circuit pipeline: its own stream, device circuit and output buffers."""

    def __init__(self, local, rank, index, n_in):
        import torch

        from halo2_aes_b200 import _ffi
        from halo2_aes_b200.backend import Circuit, Context
        self.ctx = Context(local)
        self.stream = torch.cuda.Stream(device=local)
        self.ctx.use_stream(self.stream)
        self.cir = Circuit(self.ctx, _ffi.KIND_AES, K_BENCH, N_BENCH)
        lay = self.cir.layout
        n, nb, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
        self.keys = splitmix_bytes(SEED + 1000 * (rank + 1) + 37 * index, 16 * n_in).reshape(n_in, 16)
        with torch.cuda.stream(self.stream):
            self.pts = torch.from_numpy(splitmix_bytes(
                SEED + 7000 * (rank + 1) + 91 * index, 16 * nb * 2).reshape(2, nb, 16)).cuda()
            self.advice = torch.empty((nadv, n, 4), dtype=torch.int64, device="cuda")
            self.bufs = {"Ap": torch.empty((L, n, 4), dtype=torch.int64, device="cuda"),
                         "Sp": torch.empty((L, n, 4), dtype=torch.int64, device="cuda")}
            self.Z = torch.empty((L, n, 4), dtype=torch.int64, device="cuda")
        self.stream.synchronize()

    def free(self):
        self.cir.close()
        self.advice = self.bufs = self.Z = self.pts = None


PHASES = ("witness", "lookup_prep", "permuted", "product_prep", "product")


def bind_to_gpu_numa(local):
    """Run this rank on the CPUs next to its GPU's PCIe root, so the pinned host buffers of the
    end-to-end leg (first touch) live on that NUMA node.  Returns {"node", "cpus", "why"}: on the
    pool's boxes sysfs reports numa_node -1 for the GPUs (one NUMA domain exposed to the
    container), in which case there is nothing to bind to and the reason is recorded."""
    info = {"node": None, "cpus": None, "why": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            if lo:
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info["node"] = node
        if node < 0:
            info["why"] = f"sysfs numa_node of {bdf} is {node}: no NUMA locality exposed"
        elif not cpus:
            info["why"] = "local_cpulist does not intersect this process's affinity mask"
        elif cpus == allowed:
            info["why"] = "the GPU's local CPUs are already the whole affinity mask"
            info["cpus"] = len(cpus)
        else:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as e:
        info["why"] = repr(e)
    return info


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    P = max(1, args.pipelines)
    CPS = max(1, args.circuits)  # circuits per step: the timed region is steps * CPS circuits
    n_in = 64
    pipes = [Pipe(local, rank, j, n_in) for j in range(P)]
    lay = pipes[0].cir.layout
    n, nb, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
    th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)

    def circuit(i, pipe, ev=None):
        """one circuit: fresh key, theta nudged per circuit so the per-theta bin build always runs"""
        th_i = th.copy()
        th_i[0] ^= np.uint64(i + 1)
        s, cir = pipe.stream, pipe.cir
        with torch.cuda.stream(s):
            if ev: ev[0].record(s)
            cir.witness(pipe.keys[i % n_in].tobytes(), pipe.pts[i & 1], out=pipe.advice,
                        want_ciphertexts=False)
            if ev: ev[1].record(s)
            cir.lookup_permuted(th_i, first=0, count=0, want_inputs=False, out=pipe.bufs)
            if ev: ev[2].record(s)
            cir.lookup_permuted(th_i, first=0, count=L, want_inputs=False, out=pipe.bufs)
            if ev: ev[3].record(s)
            cir.lookup_product(be, ga, first=0, count=0, out=pipe.Z)
            if ev: ev[4].record(s)
            cir.lookup_product(be, ga, first=0, count=L, out=pipe.Z)
            if ev: ev[5].record(s)

    one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e,
                        0x0e0a77c19a07df2f], dtype=torch.int64, device="cuda")

    def circuit_with_blocks(i, pipe, blocks):
        """the same call sequence for a circuit holding `blocks` <= capacity AES blocks"""
        th_i = th.copy()
        th_i[0] ^= np.uint64(i + 1)
        with torch.cuda.stream(pipe.stream):
            pipe.cir.witness(pipe.keys[i % n_in].tobytes(), pipe.pts[i & 1][:blocks], out=pipe.advice,
                             want_ciphertexts=False)
            pipe.cir.lookup_permuted(th_i, first=0, count=L, want_inputs=False, out=pipe.bufs)
            pipe.cir.lookup_product(be, ga, first=0, count=L, out=pipe.Z)

    def check_closed():
        if args.no_check:
            return
        for pipe in pipes:
            pipe.cir.lookup_check()
            # every Z column closes at 1: the lookup argument is satisfied
            assert bool((pipe.Z[:, n - 6] == one).all()), "Z[n-6] != 1"

    def step(s, ev=None):
        """one step = CPS circuits dealt to the P pipelines (events around the first one)"""
        for c in range(CPS):
            i = s * CPS + c
            circuit(i, pipes[i % P], ev if c == 0 else None)

    for s in range(args.warmup):
        step(s)
    torch.cuda.synchronize()
    check_closed()

    sampler = ClockSampler(local)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    launches0 = sum(p.ctx.kernel_launches for p in pipes)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(pipes[0].stream)
    for pipe in pipes[1:]:
        pipe.stream.wait_event(t_start)
    for s in range(args.steps):
        step(args.warmup + s, evs[s])
    for pipe in pipes[1:]:
        done = torch.cuda.Event()
        done.record(pipe.stream)
        pipes[0].stream.wait_event(done)
    t_end.record(pipes[0].stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(p.ctx.kernel_launches for p in pipes) - launches0
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    check_closed()

    # ---- phase times of the timed region (kernels of the P circuits in flight share the SMs) ---
    def phase_table(events):
        return {p: statistics.mean(e[j].elapsed_time(e[j + 1]) for e in events)
                for j, p in enumerate(PHASES)}

    ph_ms = phase_table(evs)

    # ---- per-kernel launch durations, live, with CUDA events around every launch --------------
    # (h2aes_ctx_profile) in an un-overlapped pass right after the timed region: with several
    # circuits in flight a kernel's own duration is not observable.
    peak, peak_src = measured_peak()
    NPROF = 4
    iso = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(NPROF)]
    pipes[0].ctx.profile(True)
    for j in range(NPROF):
        circuit((args.warmup + args.steps) * CPS + j, pipes[0], iso[j])
    prof = pipes[0].ctx.profile_read()
    pipes[0].ctx.profile(False)
    nc = int(pipes[0].cir.lookup_stats().sum())  # rows whose grand-product factor is not one
    # SURVEY 8(f).1, outside the step (the metric is advice + lookup columns): the permutation
    # argument's grand products for the same circuit, timed the same way
    perm = None
    if not args.no_perm:
        cirp = pipes[0].cir
        pipes[0].ctx.profile(True)
        with torch.cuda.stream(pipes[0].stream):
            cirp.witness(pipes[0].keys[0].tobytes(), pipes[0].pts[0], out=pipes[0].advice,
                         want_ciphertexts=False)
            cirp.perm_product(be, ga, pipes[0].advice)   # first call also builds the keygen-side mapping
            pipes[0].ctx.profile_read()
            Zp = None
            for _ in range(3):
                Zp = cirp.perm_product(be, ga, pipes[0].advice)
        pp = pipes[0].ctx.profile_read()
        pipes[0].ctx.profile(False)
        assert bool((Zp[-1, n - 6] == one).all()), "Z_perm does not close at 1"
        ns = Zp.shape[0]
        perm = {"n_sets": ns, "ms": sum(ms for _, ms in pp.values()) / 3,
                "kernels_ms": {k: ms / 3 for k, (cnt, ms) in pp.items()},
                "algorithmic_bytes": ns * n * 32,
                "note": "h2aes_perm_product_dev on the benchmarked circuit; not part of `value`"}
        del Zp
    ph_iso = phase_table(iso)
    alg = {  # algorithmic bytes per launch; None: not a streaming kernel
        "k_advice": nadv * n * 32 + nb * 16,      # every advice cell written once (+ plaintexts)
        "k_permuted": 2 * L * n * 32,             # A', S'
        "k_z_write": L * n * 32 + nc * 4 + L * n // 8,   # Z written; codes of the non-constant rows
                                                          # and the masks read
        "k_z_totals": nc * 4 + L * n // 8 + L * (n // 40) * 32,
        "k_lookup_hist": 2 * (nadv // 3) * n * 4, # trace words read, bin words written
    }
    # Fr-arithmetic kernels: Montgomery multiplies per launch where the count is exact, against the
    # multiply throughput measured on this chip (profiles/microbench/RESULTS.md: 67 Gmul/s)
    MUL_PEAK_G = 67.0
    muls = {"k_z_totals": nc, "k_z_write": nc}  # one multiply per non-constant row in each
    kernels = []
    for name, (cnt, ms) in prof.items():
        per_launch = ms / cnt
        row = {"kernel": name, "launches_per_step": cnt / NPROF, "ms": per_launch,
               "ms_per_step": ms / NPROF}
        if name in alg:
            gbs = alg[name] / (per_launch * 1e-3) / 1e9
            row.update({"algorithmic_bytes": alg[name], "achieved_gbs": gbs, "frac": gbs / peak})
        if name in muls:
            gm = muls[name] / (per_launch * 1e-3) / 1e9
            row.update({"fr_multiplies": muls[name], "achieved_gmul_s": gm,
                        "imad_frac_of_measured_mul_peak": gm / MUL_PEAK_G})
        kernels.append(row)
    # DRAM bytes per launch from the committed `ncu --set full` capture of the same workload
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ncu_traffic = json.load(f)["kernels"]
    except Exception:
        ncu_traffic = {}
    for row in kernels:
        if row["kernel"] in ncu_traffic:
            row["traffic"] = ncu_traffic[row["kernel"]]["traffic"]
    # the HBM roofline is reported for the dominant HBM-bound kernel; the dominant multiply-bound
    # kernel (the grand product's k_z_write / k_z_totals) gets its own object below
    dom = max((x for x in kernels if x["kernel"] in ("k_advice", "k_permuted")),
              key=lambda x: x["ms_per_step"])
    mdom = max((x for x in kernels if "achieved_gmul_s" in x), key=lambda x: x["ms_per_step"], default=None)
    roofline_multiply = None if mdom is None else {
        "bound": "imad (bn256 Montgomery multiply = 128 IMAD.WIDE)", "kernel": mdom["kernel"],
        "achieved": mdom["achieved_gmul_s"], "peak": MUL_PEAK_G, "unit": "Gmul/s",
        "frac": mdom["imad_frac_of_measured_mul_peak"], "launch_ms": mdom["ms"],
        "peak_source": "profiles/microbench/RESULTS.md (fr_mul chains, this chip)",
        "note": "useful multiplies only (one per non-constant row); executed multiplies incl. idle "
                "lanes and the scans are ~20 % more"}
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom.get("achieved_gbs"),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom.get("frac"),
                "traffic": dom.get("traffic"), "launch_ms": dom["ms"],
                "algorithmic_bytes": dom.get("algorithmic_bytes"),
                "measured": "CUDA events around every launch (h2aes_ctx_profile), un-overlapped "
                            "pass after the timed region",
                "note": "the grand-product kernels (k_z_*) are bound by integer-multiply issue, not "
                        "HBM: one bn256 Montgomery multiply = 128 IMAD.WIDE (quarter rate); measured "
                        "peak 67 Gmul/s on B200 (profiles/microbench). The HBM-bound kernels are "
                        "k_advice and k_permuted, see `kernels`."}
    step_bytes = nadv * n * 32 + 3 * L * n * 32

    # ---- end to end through the C ABI with HOST buffers (copies inside the timed region) -------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pipes[0].ctx, pipes[0].cir, pipes[0].stream, pipes[0].keys, th, be, ga,
                      rank, world)
    e2e_committed = None
    if rank == 0 and not args.no_e2e and not args.no_prover_side:
        try:
            e2e_committed = run_e2e_committed(pipes[0].ctx, pipes[0].cir, pipes[0].stream, pipes[0].keys,
                                              th, be, ga)
        except Exception as e:
            e2e_committed = {"error": repr(e)}

    # ---- BASELINE config 4 as specified: 2^20 blocks = 273 circuits dealt round-robin -----------
    config4 = None
    if not args.no_config4:
        config4 = run_config4(pipes, circuit_with_blocks, rank, world, n, one)

    # ---- BASELINE config 5: ONE K=22, N=5 circuit row-sharded over the ranks (NCCL inside the lib)
    sharded_rec = None
    if not args.no_sharded:
        for p_ in pipes[1:]:
            p_.free()
        try:
            sharded_rec = run_sharded(args, local, rank, world, th, be, ga, tiled=True)
            if world > 1:
                sharded_rec["contiguous_ownership"] = run_sharded(args, local, rank, world, th, be, ga,
                                                                  tiled=False)
        except Exception as e:  # never lose the headline line to the extra measurement
            sharded_rec = {"error": repr(e)}

    # ---- what follows the witness path in the prover; reported, not timed
    prover_side = None
    if rank == 0 and world == 1 and not args.no_prover_side:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location(
                "prover_side", os.path.join(ROOT, "profiles", "prover_side.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            with torch.cuda.stream(pipes[0].stream):
                prover_side = mod.measure(pipes[0].ctx, lay["k"], lay["n_sets"])
        except Exception as e:  # never lose the headline line to the extra measurement
            prover_side = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            cpu = cpu_baseline_sample(os.cpu_count() or 1)
        blocks = nb * CPS * args.steps * world
        cfg = workload_config(world, CPS)  # (exactly the reference arm's dict)
        out = {
            "metric": METRIC, "value": blocks / (ms_total * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 -> bn256 Fr (4x u64 Montgomery)",
            "data": "synthetic", "config": cfg,
            "circuits_in_flight_per_gpu": P, "host_numa_rank0": numa,
            "roofline": roofline, "roofline_multiply": roofline_multiply, "kernels": kernels,
            "phase_ms": ph_ms, "phase_ms_isolated": ph_iso,
            "circuit_algorithmic_bytes": step_bytes,
            "ms_per_circuit": ms_total / args.steps / CPS,
            "step_effective_gbs": step_bytes * CPS / (ms_total / args.steps * 1e-3) / 1e9,
            "witness_only_blocks_per_s": nb / (ph_iso["witness"] * 1e-3),
            "permutation_product": perm,
            "nonconstant_rows": nc, "nonconstant_row_fraction": nc / (L * n),
            "prover_side": prover_side,
            "config4_stream": config4, "sharded": sharded_rec,
            "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "e2e_committed": e2e_committed,
            "cpu_baseline": cpu,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config4(pipes, circuit_with_blocks, rank, world, n, one):
    """BASELINE config 4 as written: 2^20 independent AES-128 blocks = 272 circuits at capacity
    (3853 blocks) + one of 560 blocks, dealt round-robin to the ranks and streamed through the
    pipelines' buffers (127.5 GiB of advice in total never exists at once).  No data-path
    collective; time = max over ranks."""
    import torch
    import torch.distributed as dist

    from halo2_aes_b200 import sharded
    total_blocks = 1 << 20
    circuits = sharded.split_into_circuits(total_blocks, BLOCKS_BENCH)
    mine = sharded.circuits_of_rank(len(circuits), rank, world)
    P = len(pipes)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(pipes[0].stream)
    for p in pipes[1:]:
        p.stream.wait_event(t0)
    for j, ci in enumerate(mine):
        circuit_with_blocks(ci, pipes[j % P], circuits[ci][1])
    for p in pipes[1:]:
        done = torch.cuda.Event()
        done.record(p.stream)
        pipes[0].stream.wait_event(done)
    t1.record(pipes[0].stream)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    ok = all(bool((p.Z[:, n - 6] == one).all()) for p in pipes)
    if world > 1:
        t = torch.tensor([ms, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ok = float(t[0]), float(t[1]) == 0.0
    assert ok, "config 4: a Z column does not close at 1"
    return {"blocks": total_blocks, "circuits": len(circuits), "circuits_rank0": len(mine),
            "last_circuit_blocks": circuits[-1][1], "ms": ms, "value": total_blocks / (ms * 1e-3),
            "unit": UNIT, "check": "Z[n-6] = 1 on the last circuit of every pipeline",
            "note": "2^20 blocks as 273 independent K=20,N=5 circuits, round-robin over the ranks"}


def run_sharded(args, local, rank, world, th, be, ga, tiled=True):
    """BASELINE config 5: ONE K=22, N=5 circuit filled to capacity (15418 blocks), row-sharded over
    the ranks; both exchanges (bin-histogram all-reduce, partial-product all-gather) are issued by
    the library over its own NCCL communicator inside the timed region.  With one rank it is the
    same circuit on one GPU (the denominator of the sharded speed-up)."""
    import torch
    import torch.distributed as dist

    from halo2_aes_b200 import _ffi, sharded
    from halo2_aes_b200.backend import Context
    k, N = args.sharded_k, N_BENCH
    ctx = Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.use_stream(stream)
    sharded.library_comm_init(ctx)
    cir = sharded.RowShardedCircuit(ctx, _ffi.KIND_AES, k, N, library_comm=True, tiled=tiled)
    lay = cir.circuit.layout
    n, nb, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
    rows = cir.circuit.rows
    key = splitmix_bytes(SEED + 80, 16).tobytes()
    one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e,
                        0x0e0a77c19a07df2f], dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        pts = torch.from_numpy(splitmix_bytes(SEED + 81, 16 * nb).reshape(-1, 16)).cuda()
        advice = torch.empty((nadv, rows, 4), dtype=torch.int64, device="cuda")
        out = {nm: torch.empty((L, rows, 4), dtype=torch.int64, device="cuda") for nm in ("Ap", "Sp", "Z")}
    iters, times = 9, []
    for it in range(iters + 2):
        stream.synchronize()
        if world > 1:
            dist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        th_i = th.copy()
        th_i[0] ^= np.uint64(1000 + it)
        with torch.cuda.stream(stream):
            e[0].record(stream)
            cir.witness(key, pts, out=advice)
            e[1].record(stream)
            cir.lookup_columns(th_i, be, ga, out=out)
            e[2].record(stream)
        stream.synchronize()
        cir.circuit.lookup_check()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[0].elapsed_time(e[2])],
                         dtype=torch.float64, device="cuda")
        mine_lk = float(t[1])
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 2:
            times.append(t.tolist() + [mine_lk])
    # size-independent checks on the device: Z[0] = 1, Z[n-6] = 1, blinding rows zero,
    # (A'_i - S'_i)(A'_i - A'_{i-1}) = 0 on the shard's usable rows (inside its row ranges)
    grow = torch.from_numpy(cir.circuit.global_rows()).cuda()
    ok = True
    with torch.cuda.stream(stream):
        Z, Ap, Sp = out["Z"], out["Ap"], out["Sp"]
        at = lambda r: torch.nonzero(grow == r).flatten()
        i0, iu = at(0), at(n - 6)
        if i0.numel():
            ok &= bool((Z[:, i0[0]] == one).all())
        if iu.numel():
            ok &= bool((Z[:, iu[0]] == one).all())
        ok &= not bool(Z[:, grow > n - 6].any())
        usable = grow < n - 6
        inner = usable[1:] & (grow[1:] == grow[:-1] + 1)      # rows whose predecessor is local too
        for l0 in range(0, L, 5):  # (a column set at a time: bounded temporaries)
            a_, s_ = Ap[l0:l0 + 5], Sp[l0:l0 + 5]
            same_s = (a_[:, 1:] == s_[:, 1:]).all(dim=-1)
            same_prev = (a_[:, 1:] == a_[:, :-1]).all(dim=-1)
            ok &= bool((same_s | same_prev)[:, inner].all())
        flag = torch.tensor([1 if ok else 0], device="cuda")
    stream.synchronize()
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag) == 1, "sharded lookup columns violate an invariant"
    # median over the iterations (each already the max over ranks): a single NCCL hiccup of a few
    # hundred microseconds would move the mean of these ~1 ms steps by tens of percent
    w, lk, tot, my_lk = (float(np.median([t[j] for t in times])) for j in range(4))
    tot_mean, tot_max = float(np.mean([t[2] for t in times])), float(np.max([t[2] for t in times]))
    per_rank = [my_lk]
    if world > 1:
        g = torch.zeros(world, dtype=torch.float64, device="cuda")
        g[rank] = my_lk
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        per_rank = g.tolist()
    hist_bytes = L * 66568 * 4
    own = ("ownership tiles of 43520 rows dealt round-robin (h2aes_circuit_set_shard_tiled)" if tiled
           else "contiguous n/G rows per shard (h2aes_circuit_set_shard)")
    rec = {"config": f"ONE K={k}, N={N} circuit at capacity ({nb} blocks) row-sharded over {world} GPU(s)",
           "row_ownership": own if world > 1 else "one shard",
           "n_gpus": world, "blocks": nb, "rows_per_gpu": rows, "witness_ms": w, "lookup_columns_ms": lk,
           "ms": tot, "value": nb / (tot * 1e-3), "unit": UNIT,
           "timing": {"iterations": iters, "statistic": "median of the per-iteration max over ranks",
                      "ms_mean": tot_mean, "ms_max": tot_max},
           "bytes_per_gpu": (nadv + 3 * L) * rows * 32,
           "nvlink_bytes_per_gpu": {"histogram_allreduce": hist_bytes if world > 1 else 0,
                                    "partial_allgather": world * L * 32 if world > 1 else 0},
           "collectives": "ncclAllReduce (u32 sum, all 25 lookups in one call) + ncclAllGather (one Fr "
                          "per lookup per shard), issued by h2aes_lookup_columns_sharded_dev on the "
                          "compute stream" if world > 1 else "none (one shard)",
           "lookup_columns_ms_per_rank": per_rank,
           "bound": ("the per-theta bin order and the per-(beta,gamma) inverse tables, replicated on every "
                     "rank (~0.25 ms, do not shrink with G), ~25 kernel launches and the two collectives; "
                     "the multiplies of the grand product are balanced by the dealt tiles" if tiled else
                     "load balance of the grand product: with contiguous ownership the shards at the END "
                     "of the columns hold the rows where A' is not zero, every one of which costs two "
                     "multiplies (see lookup_columns_ms_per_rank); then the replicated per-theta / "
                     "per-(beta,gamma) tables and the two collectives") + ".  The bulk columns never "
                    "cross NVLink",
           "checks": "Z[0]=1, Z[n-6]=1, blinding rows zero, permuted-pair constraint on every shard: ok"}
    del advice, out, pts
    cir.circuit.close()
    ctx.comm_destroy()
    return rec


def host_d2h_ceiling(stream, rank, world):
    """What the box's host side can absorb: every rank copies 1 GiB device -> pinned host 4 times
    with plain cudaMemcpyAsync (one copy per buffer), all ranks at once; GB/s per rank (min) and
    summed over the ranks.  The end-to-end leg copies 3.05 GB per circuit, so this is its bound."""
    import torch
    import torch.distributed as dist
    nbytes = 1 << 30
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    with torch.cuda.stream(stream):
        host.copy_(dev, non_blocking=True)
    stream.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(4):
            host.copy_(dev, non_blocking=True)
    stream.synchronize()
    dt = time.perf_counter() - t0
    mine = 4 * nbytes / dt / 1e9
    lo, total = mine, mine
    if world > 1:
        t = torch.tensor([mine, -mine], dtype=torch.float64, device="cuda")
        s = t[:1].clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, lo = float(s[0]), -float(t[1])
    return {"gbs_per_rank_min": lo, "gbs_all_ranks": total, "ranks": world,
            "how": "4 x 1 GiB cudaMemcpyAsync device->pinned host per rank, all ranks concurrently"}


def run_e2e_committed(ctx, cir, stream, keys, th, be, ga):
    """The device-resident alternative the (f) rows enable: plaintexts from pinned host memory,
    witness -> A', S', Z on the device, every witness-dependent column committed on the device
    (h2aes_msm_dev against synthetic bases), only the 64-byte commitments come
    back.  Separate from `e2e` (which returns every column); bound by the commitments."""
    import torch

    from halo2_aes_b200.backend import Msm, synthetic_bases
    lay = cir.layout
    n, nb, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
    with torch.cuda.stream(stream):
        bases = synthetic_bases(ctx, n, seeded_fr(SEED + 9))
        msm = Msm(ctx, bases).enable_runs()
        pts_h = torch.from_numpy(splitmix_bytes(SEED + 33, 16 * nb).reshape(nb, 16)).pin_memory()
        pts_d = torch.empty((nb, 16), dtype=torch.uint8, device="cuda")
        advice = torch.empty((nadv, n, 4), dtype=torch.int64, device="cuda")
        bufs = {nm: torch.empty((L, n, 4), dtype=torch.int64, device="cuda") for nm in ("Ap", "Sp")}
        Z = torch.empty((L, n, 4), dtype=torch.int64, device="cuda")
        com_d = torch.empty((nadv + 3 * L, 2, 4), dtype=torch.int64, device="cuda")
    com_h = torch.empty((nadv + 3 * L, 2, 4), dtype=torch.int64).pin_memory()

    def step(i):
        th_i = th.copy()
        th_i[0] ^= np.uint64(i + 177)
        with torch.cuda.stream(stream):
            pts_d.copy_(pts_h, non_blocking=True)
            cir.witness(keys[i % len(keys)].tobytes(), pts_d, out=advice, want_ciphertexts=False)
            msm.commit(advice, out=com_d[:nadv])
            cir.lookup_permuted(th_i, first=0, count=L, want_inputs=False, out=bufs)
            msm.commit(bufs["Ap"], out=com_d[nadv:nadv + L])
            msm.commit(bufs["Sp"], out=com_d[nadv + L:nadv + 2 * L])
            cir.lookup_product(be, ga, first=0, count=L, out=Z)
            msm.commit(Z, out=com_d[nadv + 2 * L:])
            com_h.copy_(com_d, non_blocking=True)
        stream.synchronize()

    step(0)
    steps = 2
    t0 = time.perf_counter()
    for i in range(steps):
        step(1 + i)
    dt = time.perf_counter() - t0
    assert bool(com_h.any()), "no commitment came back"
    msm.close()
    return {"value": nb * steps / dt, "unit": UNIT, "steps": steps, "ms_per_circuit": 1e3 * dt / steps,
            "h2d_bytes_per_step": nb * 16 + 16 + 3 * 32, "d2h_bytes_per_step": (nadv + 3 * L) * 64,
            "columns_committed": nadv + 3 * L,
            "api": "h2aes_witness_dev + h2aes_lookup_permuted_dev + h2aes_lookup_product_dev + "
                   "h2aes_msm_dev (synthetic bases), commitments only copied back",
            "note": "rank 0 only; bound by the G1 multi-scalar multiplications, not by PCIe"}


def run_e2e(args, ctx, cir, stream, keys, th, be, ga, rank, world):
    """Same step through the reference-facing _host entry points: plaintexts start in pinned
    host memory, every produced column ends in pinned host memory."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from halo2_aes_b200 import _ffi
    L_ = _ffi.lib()
    lay = cir.layout
    n, nb, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
    pts = torch.from_numpy(splitmix_bytes(SEED + 31 + rank, 16 * nb)).pin_memory()
    adv = torch.empty((nadv, n, 4), dtype=torch.int64).pin_memory()
    Ap = torch.empty((L, n, 4), dtype=torch.int64).pin_memory()
    Sp = torch.empty((L, n, 4), dtype=torch.int64).pin_memory()
    Zh = torch.empty((L, n, 4), dtype=torch.int64).pin_memory()
    cts = torch.empty((nb, 16), dtype=torch.uint8).pin_memory()
    u64p = C.POINTER(C.c_uint64)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def step(i):
        key = np.ascontiguousarray(keys[i % len(keys)])
        th_i = th.copy()
        th_i[0] ^= np.uint64(i + 77)
        _ffi.check(L_.h2aes_witness_host(cir._h, key.ctypes.data_as(C.POINTER(C.c_uint8)), vp(pts),
                                         nb, vp(adv), vp(cts)), ctx._h)
        _ffi.check(L_.h2aes_lookup_permuted_host(cir._h, th_i.ctypes.data_as(u64p), 0, L, None,
                                                 None, vp(Ap), vp(Sp)), ctx._h)
        _ffi.check(L_.h2aes_lookup_product_host(cir._h, be.ctypes.data_as(u64p),
                                                ga.ctypes.data_as(u64p), 0, L, vp(Zh)), ctx._h)

    steps = max(1, min(args.steps, args.e2e_steps))
    step(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(1 + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    one = [-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f]
    assert Zh[:, n - 6].tolist() == [one] * L, "e2e: Z[n-6] != 1"
    d2h = (nadv + 3 * L) * n * 32 + nb * 16
    ceiling = host_d2h_ceiling(stream, rank, world)
    return {"value": nb * steps * world / dt, "unit": UNIT, "steps": steps,
            "achieved_d2h_gbs_all_ranks": world * d2h * steps / dt / 1e9,
            "host_d2h_ceiling": ceiling,
            "ms_per_step": 1e3 * dt / steps, "h2d_bytes_per_step": nb * 16 + 16 + 3 * 32,
            "d2h_bytes_per_step": d2h,
            "api": "h2aes_witness_host + h2aes_lookup_permuted_host + h2aes_lookup_product_host "
                   "(pinned host buffers; every column copied back)"}


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line there.
    Everything but emit() goes to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--circuits", type=int, default=12,
                    help="circuits per step per GPU (one step = this many whole K=20,N=5 circuits)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-perm", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the 273-circuit stream of config 4")
    ap.add_argument("--no-sharded", action="store_true", help="skip the row-sharded K=22 circuit (config 5)")
    ap.add_argument("--sharded-k", type=int, default=22)
    ap.add_argument("--no-check", action="store_true",
                    help="tuning only: skip the Z[n-6] == 1 assertions (diagnostic library variants)")
    ap.add_argument("--no-prover-side", action="store_true",
                    help="skip the NTT / lookup-quotient measurement (SURVEY 8f items 3/4)")
    ap.add_argument("--pipelines", type=int, default=4,
                    help="independent circuits in flight per GPU (each on its own stream)")
    args = ap.parse_args()
    protect_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
