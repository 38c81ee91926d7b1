#!/usr/bin/env python
"""bench.py — sample-pairs/s of the all-pairs FunUniFrac hot path (push-up + weighted-L1 pairwise) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples 10000] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "all-pairs FunUniFrac (10K samples, KO tree)"): a synthetic full-size KO tree
(fununifrac_b200.synth, seed 0: 30,002 nodes / 30,001 edges / 26,000 leaves) and 10,000 synthetic profiles
(seed 1), weighted L1.  A step = one pass of the hot path over all samples: push-up of every profile, the
N(N-1)/2 distances, the assembled symmetric float64 matrix on rank 0.  Total work is fixed as GPUs are added
(samples and tile rows are sharded) -> "scaling": "strong".

  value     pairs/s with the un-pushed profiles resident in HBM (float64, the reference's dtype), device-timed
  e2e       the same through the host-buffer entry (`fuf_all_pairs_host` at 1 GPU; the sharded host path at N>1):
            pinned host profiles -> device -> ... -> distance matrix back in host memory, copies inside the timing
  roofline  the pairwise kernel against the FP32 pipe (2 lane-ops per (pair, basis entry)), measured
            live with CUDA events around its launches; `roofline_pushup` the push-up against measured HBM copy
  cpu_baseline / --impl reference : the CPU oracle port (oracle/oracle.c, OpenMP, all host threads) on a bounded
            sample of the same workload.  The reference itself is pure Python (nothing to compile into
            oracle/_ref) and /root/reference does not exist on the GPU box, so kind = "port".
  workloads the default run also carries BASELINE's other two device workloads on the same JSON line, each with its
            own value / roofline / e2e / parity / cpu_baseline: `workloads.l2` (config 4: 50,000 samples, --L2 on the
            tcgen05 tensor cores, roofline against the int8 MMA rate MEASURED in the same run) and `workloads.diffab`
            (config 5: 500 samples, difference vectors streamed to pinned host memory, roofline against the measured
            PCIe rate).  `--workload l1|l2|diffab|cli` runs one of them alone.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sample-pairs/sec, all-pairs FunUniFrac (10K samples, KO tree)"
UNIT = "pairs/s"
CPU_SAMPLE = 3072     # rows of the same synthetic set timed on the host cores (4,717,056 pairs, ~10 s on 16 cores)


_REAL_STDOUT = None


def claim_stdout():
    """Everything libraries print to stdout (e.g. NCCL's version banner) goes to stderr; the one JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def profile_traffic(name):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the committed ``ncu --set full`` capture
    ``profiles/<name>`` (metric,unit,value rows), or None.  Captured on the default workload (N = 10,000, 1 GPU)."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", name.replace("r2_", "r1_"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        with open(path) as f:
            for line in f:
                parts = [x.strip().strip('"') for x in line.split(",")]
                if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(parts[2]) * scale[parts[1]]
                    seen += 1
    except (OSError, KeyError, ValueError):
        return None
    return total if seen == 2 else None


def build_tree():
    from fununifrac_b200 import synth
    from fununifrac_b200.factory import make_emd_input, make_tree

    with tempfile.TemporaryDirectory() as tmp:
        path = synth.synth_ko_tree(os.path.join(tmp, "ko_tree.txt"), seed=0)
        inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
    return inp, synth.leaves_of(inp.parent)


def _rows_job(args):
    from fununifrac_b200 import synth

    leaves, n, r0, r1 = args
    return r0, synth.synth_profile_rows(leaves, n, range(r0, r1), seed=1)


def make_profiles(leaves, n, lo, hi, out):
    """Rows lo..hi-1 of the seeded synthetic set into out[hi - lo, n] (float64), generated by a process pool."""
    from concurrent.futures import ProcessPoolExecutor

    out[:] = 0
    jobs = [(leaves, n, r, min(hi, r + 128)) for r in range(lo, hi, 128)]
    workers = max(1, min(16, (os.cpu_count() or 2) - 1))
    with ProcessPoolExecutor(workers) as ex:
        for r0, (indptr, indices, values) in ex.map(_rows_job, jobs):
            for i in range(len(indptr) - 1):
                s, e = indptr[i], indptr[i + 1]
                out[r0 - lo + i, indices[s:e]] = values[s:e]
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self.nv = pynvml
            try:    # maps the CUDA ordinal to the NVML device (CUDA_VISIBLE_DEVICES aware)
                self.h = torch.cuda._get_pynvml_handler(index)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.02)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_oracle_run(inp, X, repeat=1):
    """Push-up + all pairs of X on the host cores with the oracle port; returns (pairs/s, seconds, threads, D)."""
    from oracle import oracle_c

    oracle_c.set_threads(os.cpu_count() or 1)     # all host threads, whatever OMP_NUM_THREADS the launcher exported
    n_pairs = X.shape[0] * (X.shape[0] - 1) // 2
    best, D = None, None
    for _ in range(repeat):
        t0 = time.perf_counter()
        P = oracle_c.push_up(inp.parent, inp.edge_length, X, False)
        D = oracle_c.pairwise(P, False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_pairs / best, best, oracle_c.num_threads(), D


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port, all host threads), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp, leaves = build_tree()
    n = len(inp.basis)
    Ns = min(args.samples, CPU_SAMPLE)
    X = make_profiles(leaves, n, 0, Ns, np.zeros((Ns, n)))
    for _ in range(args.warmup):
        cpu_oracle_run(inp, X[:256])
    t0 = time.perf_counter()
    threads = 1
    for _ in range(args.steps):
        _, _, threads, _ = cpu_oracle_run(inp, X)
    dt = (time.perf_counter() - t0) / args.steps
    pairs = Ns * (Ns - 1) // 2
    value = pairs / dt
    sample = f"first {Ns} of the {args.samples} synthetic samples ({pairs} pairs), push-up + all pairs, float64"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic full KO tree (n={n} nodes), {args.samples} samples, weighted L1",
                   "timed_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def device_profiles(leaves, n, N, seed=1):
    """Random profiles generated on the device: uniform abundances on a random 20 % of the leaves, normalised."""
    import torch

    from fununifrac_b200 import engine

    g = torch.Generator(device="cuda").manual_seed(seed)
    lv = torch.from_numpy(leaves).cuda()
    X = engine.profile_matrix(N, n, torch.float64, "cuda")
    for b in range(0, N, 4096):
        e = min(N, b + 4096)
        v = torch.rand((e - b, len(leaves)), generator=g, device="cuda", dtype=torch.float64)
        v *= torch.rand((e - b, len(leaves)), generator=g, device="cuda") < 0.2
        X[b:e, lv] = v / v.sum(1, keepdim=True)
    return X


class Env:
    """What every workload needs: ranks, device, process group, the synthetic tree, measured peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.args = torch, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = f"cuda:{self.local_rank}"
        self.dist = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            self.dist = dist
        self.peaks = load_peaks()
        self.inp, self.leaves = build_tree()
        self.n = len(self.inp.basis)
        self.sm_count = torch.cuda.get_device_properties(self.local_rank).multi_processor_count

    def context(self):
        from fununifrac_b200 import engine
        return engine.TreeContext(self.inp.parent, self.inp.edge_length, device=self.local_rank)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def gather_rows(self, X_local, lo, sub):
        """Rows ``sub`` (global sample indices, the same array on every rank) of the row-sharded device matrix
        X_local = X[lo:lo + len] collected on rank 0 as a host array (None elsewhere)."""
        torch = self.torch
        mine = [i for i, s in enumerate(sub) if lo <= s < lo + X_local.shape[0]]
        buf = torch.zeros((len(sub), X_local.shape[1]), dtype=X_local.dtype, device=self.dev)
        if mine:
            buf[mine] = X_local[torch.tensor([int(sub[i]) - lo for i in mine], device=self.dev)]
        if self.dist is not None:
            self.dist.reduce(buf, dst=0)               # every row is non-zero on exactly one rank
        return buf.cpu().numpy() if self.rank == 0 else None

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def timed_steps(env, step, steps, warmup):
    """W warm-up steps, then K steps bracketed by barrier + synchronize; device time (CUDA events), max over ranks."""
    for _ in range(warmup):
        step()
    env.barrier()
    t0, t1 = env.event(), env.event()
    t0.record()
    for _ in range(steps):
        out = step()
    t1.record()
    env.barrier()
    return env.max_over_ranks([t0.elapsed_time(t1)])[0] / steps, out


def host_timed_steps(env, step, steps, warmup):
    """The same for calls that block the host (host-buffer API): the larger of device time and wall clock."""
    for _ in range(warmup):
        step()
    env.barrier()
    w0 = time.perf_counter()
    t0, t1 = env.event(), env.event()
    t0.record()
    for _ in range(steps):
        out = step()
    t1.record()
    env.barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    return env.max_over_ranks([max(t0.elapsed_time(t1), wall_ms)])[0] / steps, out


def measure_host_links(env):
    """Host-link rates of this box as the job sees them: every rank copies 4 x 256 MiB between pinned host memory and its
    GPU at the same time (the ranks' links may share PCIe switches and root ports), best of three trials, summed over the
    ranks.  GB/s, per direction (each direction measured alone)."""
    torch = env.torch
    hp = torch.empty(1 << 26, dtype=torch.float32, pin_memory=True)
    dp = torch.empty(1 << 26, dtype=torch.float32, device=env.dev)
    out = {}
    for name, (dst, src) in (("h2d_gbs", (dp, hp)), ("d2h_gbs", (hp, dp))):
        dst.copy_(src, non_blocking=True)
        best = 0.0
        for _ in range(3):
            env.barrier()
            e0, e1 = env.event(), env.event()
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 4 * hp.numel() * 4 / e0.elapsed_time(e1) / 1e6)
        t = torch.tensor([best], dtype=torch.float64, device=env.dev)
        if env.dist is not None:
            env.dist.all_reduce(t)
        out[name] = float(t[0])
    out["note"] = "pinned cudaMemcpy, all ranks at once, summed over the ranks"
    return out


def rel_err_block(D, Dref):
    iu = np.triu_indices(Dref.shape[0], 1)
    den = np.where(Dref[iu] > 0, Dref[iu], 1.0)
    return float((np.abs(D[iu] - Dref[iu]) / den).max()), len(iu[0])


# =====================================================================================================================
# config 5: --diffab
# =====================================================================================================================
def run_diffab(env, args, N=500, steps=None):
    """BASELINE config 5: --diffab.  Every pair's signed difference vector P_i - P_j (n entries; subtracted in float64
    from the float64 pushed profiles, rounded once to fp32 for the wire) is produced by K4 in itertools.combinations
    order in blocks of 1,024 pairs and moved to pinned host memory by the copy engine, double-buffered on both sides.
    Under torchrun every rank pushes all samples up and streams its own contiguous share of the pairs over its own PCIe
    link (no exchange at all).  Bound: the device-to-host link (measured with a pinned cudaMemcpy in the same run)."""
    torch = env.torch
    from fununifrac_b200 import engine

    world, rank, dev, n = env.world, env.rank, env.dev, env.n
    steps = steps or args.steps
    warmup = max(1, min(args.warmup, 3))
    ctx = env.context()
    X = device_profiles(env.leaves, n, N)               # same seed on every rank: the same samples
    job = ctx.push_up_job(X, engine.MODE_L1)
    pairs = N * (N - 1) // 2
    p0, p1 = engine.diffab_pair_range(N, world, rank)
    # link peak: pinned device-to-host copy of 1 GiB (all ranks at once, like the stream itself)
    hp = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)
    dp = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    hp.copy_(dp, non_blocking=True)
    link = 0.0
    for _ in range(3):          # best of three trials of 4 GiB each (ranks in step: the links may share a PCIe switch)
        env.barrier()
        e0, e1 = env.event(), env.event()
        e0.record()
        for _ in range(4):
            hp.copy_(dp, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        link = max(link, 4 * hp.numel() * 4 / e0.elapsed_time(e1) / 1e6)
    del hp, dp

    def step():
        chk = 0.0
        for b, e, rows in engine.iter_diffab_blocks(ctx, job.P64T, N, engine.MODE_L1, pairs_per_block=1024,
                                                    out_dtype=torch.float32, pair_begin=p0, pair_end=p1):
            chk += float(rows[0, -2])          # touch the block on the host (it is reused two blocks later)
        return chk

    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ms, _ = host_timed_steps(env, step, steps, warmup)
    clocks = sampler.finish()
    launches = (ctx.launch_count() - launches0) // (steps + warmup) * steps
    lk = torch.tensor([link], dtype=torch.float64, device=dev)
    if env.dist is not None:
        env.dist.all_reduce(lk)                            # aggregate link rate
    link_sum = float(lk[0])
    nbytes = pairs * n * 4
    out = None
    if rank == 0:
        # parity: the first 499 rows (pairs (0, 1..N-1)) and the last block against float64 differences of the ORACLE's
        # pushed profiles
        from oracle import oracle_c
        P = oracle_c.push_up(env.inp.parent, env.inp.edge_length, X.cpu().numpy(), False)
        b, e, rows = next(engine.iter_diffab_blocks(ctx, job.P64T, N, engine.MODE_L1, pairs_per_block=N - 1,
                                                    out_dtype=torch.float32))
        want = P[0][None, :] - P[1:N]
        scale = np.abs(P).max()
        err = float(np.abs(rows - want).max() / np.abs(want).max())
        assert np.all(np.abs(rows - want) <= 6.1e-8 * np.abs(want) + 1e-13 * scale), err
        b, e, rows = next(engine.iter_diffab_blocks(ctx, job.P64T, N, engine.MODE_L1, pairs_per_block=N - 1,
                                                    out_dtype=torch.float32, pair_begin=pairs - 3, pair_end=pairs))
        want = np.stack([P[N - 3] - P[N - 2], P[N - 3] - P[N - 1], P[N - 2] - P[N - 1]])
        assert np.all(np.abs(rows - want) <= 6.1e-8 * np.abs(want) + 1e-13 * scale)
        out = {
            "metric": "sample-pairs/sec, --diffab per-pair difference vectors streamed to host", "value": pairs / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 subtract, f32 on the wire",
            "data": "synthetic",
            "config": {"workload": f"synthetic full KO tree (n={n} nodes), {N} samples, --diffab, dense fp32 rows of "
                                   f"{n} entries per pair to pinned host memory", "pairs_per_block": 1024,
                       "parallelism": f"contiguous pair ranges over {world} GPU(s), one PCIe link each"},
            "roofline": {"bound": "pcie", "kernel": "diffab_kernel<double,float,L1> + async device-to-host copies",
                         "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": link_sum, "unit": "GB/s",
                         "frac": nbytes / (ms * 1e-3) / 1e9 / link_sum, "traffic": None,
                         "peak_source": "pinned cudaMemcpy device-to-host, 4 x 1 GiB on every rank at once (best of three "
                                        "trials), summed over the ranks, same run", "algorithmic_bytes": nbytes},
            "e2e": {"value": pairs / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": nbytes,
                    "note": "the stream IS the end-to-end path: rows land in host memory"},
            "gpu_launches": int(launches), "clocks": clocks,
            "parity": {"max_err_vs_oracle_differences_rel_to_max": err, "tolerance": "6.1e-8 per entry (one fp32 rounding)"},
        }
    del job, X
    ctx.close()
    torch.cuda.empty_cache()
    return out


# =====================================================================================================================
# config 4: --L2 on the tensor cores
# =====================================================================================================================
def measure_i8_peak(env, ctx):
    """Tensor-pipe rate of the Gram kernel's own instruction (tcgen05.mma kind::i8, M = 128 per CTA, N = 128, K = 32) with
    operands resident in shared memory: burst = best of several ~3 ms launches after an idle gap, sustained = 1.5 s of
    back-to-back launches (power-capped clocks).  TOP/s, for both CTA groupings."""
    res = {}
    for group in (1, 2):
        ms, ops = ctx.i8_mma_peak(group, 2048, 1)              # warm-up / calibration
        iters = max(256, int(2048 * 3.0 / max(ms, 1e-3)))
        burst = 0.0
        for _ in range(5):
            time.sleep(0.05)
            ms, ops = ctx.i8_mma_peak(group, iters, 1)
            burst = max(burst, ops / (ms * 1e-3) / 1e12)
        launches = max(1, int(1500.0 / max(ms, 1e-3)))
        ms, ops = ctx.i8_mma_peak(group, iters, launches)
        res[group] = (burst, ops / (ms * 1e-3) / 1e12)
    best = max(res, key=lambda g: res[g][1])
    return {"burst": max(v[0] for v in res.values()), "sustained": res[best][1],
            "by_cta_group": {str(g): {"burst": v[0], "sustained": v[1]} for g, v in res.items()}}


def run_l2(env, args, N=50000, steps=None):
    """BASELINE config 4: synthetic full KO tree, --L2 through the tcgen05 Gram path, 1/2/4/8 GPUs.
    Profiles are generated on the device (uniform random abundances on a random 20 % of the leaves, normalised;
    rank r generates its own shard with seed 1 + r): 50,000 samples in the host generator would take minutes.
    Parity is asserted on a 512-sample block drawn over all samples (every rank's shard and tile rows)."""
    torch = env.torch
    from fununifrac_b200 import engine, sharded

    world, rank, dev, n = env.world, env.rank, env.dev, env.n
    steps = steps or args.steps
    warmup = max(args.warmup, 3)
    ctx = env.context()
    peak = measure_i8_peak(env, ctx)
    job = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, engine.MODE_L2, world, rank, env.dist)
    assert job.use_gram
    X = device_profiles(env.leaves, n, job.hi - job.lo, seed=1 + rank)
    pairs = N * (N - 1) // 2
    gram_ms = []

    def step():
        D = job.step(X)
        if job.rows:
            gram_ms.append(ctx.gram_last_kernel())
        return D

    for _ in range(warmup):
        step()
    gram_ms.clear()
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ms_step, D = timed_steps(env, step, steps, 0)
    clocks = sampler.finish()
    launches = ctx.launch_count() - launches0
    k_ms = float(np.mean([m for m, _ in gram_ms])) if gram_ms else 0.0
    k_ms = env.max_over_ranks([k_ms])[0]
    flops = gram_ms[-1][1] if gram_ms else 0.0
    buckets, bypassed = ctx.gram_last_info() if job.rows else (0, 0)
    flagged = job.last_flagged

    # ---- parity block over ALL samples: rows from every rank's shard, distances from every rank's tile rows ----------
    Ns = min(N, 512)
    sub = np.sort(np.random.default_rng(11).choice(N, Ns, replace=False))
    Xs = env.gather_rows(X, job.lo, sub)
    D_blk = None
    if rank == 0:
        idx = torch.from_numpy(sub).to(dev)
        D_blk = D[idx][:, idx].cpu().numpy()
        assert bool(torch.equal(D[:4096, :4096], D[:4096, :4096].T)) and not bool(torch.diagonal(D).any())
    del D

    # ---- end to end through the host-buffer API: pinned float64 profiles in, float64 matrix in host memory out --------
    e2e = None
    need = (job.hi - job.lo) * n * 8 + N * N * 8
    import psutil
    fits = psutil.virtual_memory().available > need * (1 if world == 1 else 2) + (16 << 30) and not args.no_e2e
    if env.dist is not None:        # one decision for all ranks (the leg is full of collectives)
        ok = torch.tensor([1 if fits else 0], dtype=torch.int32, device=dev)
        env.dist.all_reduce(ok, op=env.dist.ReduceOp.MIN)
        fits = bool(int(ok[0]))
    if fits:
        e2e_steps = max(1, min(steps, 3))
        X_pin = torch.empty((max(job.hi - job.lo, 1), n), dtype=torch.float64, pin_memory=True)[:job.hi - job.lo]
        X_pin.copy_(X)
        if world == 1:
            job.release_device_matrix()
            D_pin = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
            Xh, Dh = X_pin.numpy(), D_pin.numpy()

            def host_step():
                ctx.all_pairs_host(Xh, engine.MODE_L2, Dh)
                return Dh
        else:
            shared = sharded.SharedHostMatrix(N, world, rank, env.dist)

            def host_step():
                return job.step_host(X_pin, shared)
        e2e_ms, Dh_out = host_timed_steps(env, host_step, e2e_steps, 1)
        e2e = {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_bytes_per_step": int(N) * n * 8, "d2h_bytes_per_step": int(N) * int(N) * 8,
               "api": "fuf_all_pairs_host (C ABI, pinned host float64 profiles in, host float64 matrix out)" if world == 1
               else "sharded.ShardedAllPairs.step_host over the C ABI: pinned host shards in on every rank, one shared "
                    "host matrix out (each rank stores its tile rows over its own PCIe link)"}
        e2e["host_links"] = measure_host_links(env)
        e2e["transfer_ms_at_link_rate"] = (e2e["h2d_bytes_per_step"] / e2e["host_links"]["h2d_gbs"] +
                                           e2e["d2h_bytes_per_step"] / e2e["host_links"]["d2h_gbs"]) / 1e6
        if rank == 0:
            e2e["block_vs_device_matrix_max_rel_diff"] = rel_err_block(Dh_out[np.ix_(sub, sub)], D_blk)[0]
            assert e2e["block_vs_device_matrix_max_rel_diff"] < 1e-9
        del Dh_out
        if world > 1:
            env.barrier()
            shared.close()
    out = None
    if rank == 0:
        from oracle import oracle_c
        threads = os.cpu_count() or 1
        oracle_c.set_threads(threads)
        t0c = time.perf_counter()
        Dref = oracle_c.pairwise(oracle_c.push_up(env.inp.parent, env.inp.edge_length, Xs, True), True)
        cpu_s = time.perf_counter() - t0c
        err, n_checked = rel_err_block(D_blk, Dref)
        assert err < 1e-6, err
        # B200_PROFILING.md: the burst figure for a kernel timed alone, the sustained one for a kernel timed inside a long
        # step.  Here both are MEASURED with the kernel's own MMA (fuf_i8_mma_peak), not derived from bf16.
        long_run = ms_step * steps > 1000.0
        i8_peak = peak["sustained"] if long_run else peak["burst"]
        achieved = flops / (k_ms * 1e-3) / 1e12 if k_ms else 0.0
        out = {
            "metric": "sample-pairs/sec, all-pairs FunUniFrac --L2 (tensor-core Gram path)", "value": pairs / (ms_step * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8x4 digits (exact int32) + f64",
            "data": "synthetic",
            "config": {"workload": f"synthetic full KO tree (n={n} nodes), {N} samples, --L2, device-generated profiles "
                                   f"(20% of the leaves present)", "pairs_flagged_for_float64": flagged,
                       "parallelism": f"samples and 128-row tile rows sharded over {world} GPU(s)",
                       "scale_buckets": buckets, "columns_in_float64": bypassed},
            "roofline": {"bound": "tensor", "kernel": ("gram_tile_kernel" if os.environ.get("FUF_GRAM_2CTA") == "0"
                                                       else "gram_pair_kernel (cta_group::2)")
                                   + " (tcgen05.mma kind::i8, 10 digit products per term)",
                         "achieved": achieved, "peak": i8_peak, "unit": "TOP/s", "frac": achieved / i8_peak if i8_peak else None,
                         "traffic": profile_traffic(PROFILE_GRAM) if (world == 1 and N == 50000) else None,
                         "traffic_source": f"profiles/{PROFILE_GRAM} (ncu --set full, N = 50,000, one GPU)",
                         "ms_per_launch": k_ms,
                         "peak_source": "measured in this run: fuf_i8_mma_peak, the kernel's own tcgen05.mma kind::i8 "
                                        "(operands resident in shared memory, every SM), "
                                        + ("sustained over 1.5 s (timed region > 1 s, power-capped)" if long_run else "burst"),
                         "peak_measured": peak,
                         "frac_of_burst": achieved / peak["burst"], "frac_of_sustained": achieved / peak["sustained"],
                         "executed_ops_per_launch_rank0": flops,
                         "algorithmic_tflops": 2.0 * pairs * n / world / (k_ms * 1e-3) / 1e12 if k_ms else None},
            "kernel_share": {"gram_kernel": k_ms / ms_step},
            "gpu_launches": int(launches), "clocks": clocks,
            "parity": {"max_rel_err_vs_oracle": err, "pairs_checked": n_checked, "tolerance": 1e-6,
                       "block": f"{Ns} samples drawn over all {N} (every rank's shard and tile rows)"},
            "cpu_baseline": {"value": n_checked / cpu_s, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{Ns} of the samples ({n_checked} pairs), --L2 push-up + all pairs, float64, "
                                       f"{cpu_s:.2f} s", "host_cpus": os.cpu_count()},
        }
        if e2e is not None:
            out["e2e"] = e2e
    del job, X
    ctx.close()
    torch.cuda.empty_cache()
    return out


# =====================================================================================================================
# the headline: configs 2/3, weighted L1
# =====================================================================================================================
def run_l1(env, args):
    torch = env.torch
    from fununifrac_b200 import engine, sharded, synth

    world, rank, dev, n, N = env.world, env.rank, env.dev, env.n, args.samples
    peaks, inp, leaves = env.peaks, env.inp, env.leaves
    ctx = env.context()
    ops = sharded.CudaOps(ctx)
    job = sharded.ShardedAllPairs(ops, N, engine.MODE_L1, world, rank, env.dist)
    lo, hi = job.lo, job.hi
    # un-pushed profiles of this rank: pinned host copy (e2e source) and HBM-resident copy (value source)
    X_pin = torch.zeros((max(hi - lo, 1), n), dtype=torch.float64, pin_memory=True)
    make_profiles(leaves, n, lo, hi, X_pin.numpy()[:hi - lo])
    X_pin = X_pin[:hi - lo]
    # HBM-resident copy with 128-byte aligned rows (engine.profile_matrix): the layout the push-up's TMA boxes like
    X_dev = engine.profile_matrix(hi - lo, n, torch.float64, dev)
    X_dev.copy_(X_pin)
    pairs = N * (N - 1) // 2
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---------------- device-resident steps (value) ----------------
    k_push, k_pair = [], []

    def device_step(timed=True):
        """One pass (sharded.ShardedAllPairs.step); per-kernel CUDA events on the launching (current) stream."""
        if not timed:
            return job.step(X_dev)
        marks = {}

        def mark(name):
            marks[name] = env.event()
            marks[name].record()
        D = job.step(X_dev, mark)
        k_push.append((marks["push_begin"], marks["push_end"]))
        k_pair.append((marks["pair_begin"], marks["pair_end"]))
        return D

    for _ in range(warmup):
        device_step(False)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ms_step, D = timed_steps(env, device_step, steps, 0)
    launches = ctx.launch_count() - launches0
    refined = job.last_flagged if world > 1 else ctx.refine_count()
    clocks = sampler.finish()
    ms_push = sum(a.elapsed_time(b) for a, b in k_push) / steps
    ms_pair = sum(a.elapsed_time(b) for a, b in k_pair) / steps
    ms_push_max, ms_pair_max = env.max_over_ranks([ms_push, ms_pair])
    value = pairs / (ms_step * 1e-3)

    # ---------------- end to end through the host-buffer API ----------------
    if world == 1:
        D_pin = torch.zeros((N, N), dtype=torch.float64, pin_memory=True)
        Xh, Dh = X_pin.numpy(), D_pin.numpy()

        def host_step():
            ctx.all_pairs_host(Xh, engine.MODE_L1, Dh)
            return Dh
    else:
        # one host matrix shared by the ranks of the node: every rank DMA-writes its own tile rows (and their mirror)
        shared = sharded.SharedHostMatrix(N, world, rank, env.dist)

        def host_step():
            return job.step_host(X_pin, shared)
    e2e_ms, Dh_out = host_timed_steps(env, host_step, steps, 2)
    e2e = {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(N) * n * 8, "d2h_bytes_per_step": int(N) * int(N) * 8,
           "api": "fuf_all_pairs_host (C ABI, pinned host float64 profiles in, host float64 matrix out)" if world == 1
           else "sharded.ShardedAllPairs.step_host over the C ABI: pinned host shards in on every rank, one shared "
           "host matrix out (each rank stores its tile rows over its own PCIe link)"}

    e2e["host_links"] = measure_host_links(env)
    hl = e2e["host_links"]
    e2e["transfer_ms_at_link_rate"] = (e2e["h2d_bytes_per_step"] / hl["h2d_gbs"] + e2e["d2h_bytes_per_step"] / hl["d2h_gbs"]) / 1e6

    line = None
    if rank == 0:
        # ---------------- parity + CPU baseline (outside every timed region) ----------------
        D_host = D.cpu().numpy()
        assert np.array_equal(D_host, D_host.T) and not np.any(np.diag(D_host))
        with np.errstate(divide="ignore", invalid="ignore"):   # the host path computes D band by band: same fp32 partial
            e2e_diff = float(np.nanmax(np.abs(Dh_out - D_host) / D_host))   # sums, other order of the float64 additions
        out = {}
        if not args.no_cpu_baseline:
            parity = {"tolerance": 1e-6, "e2e_vs_device_matrix_max_rel_diff": e2e_diff}
            # (a) a block drawn over ALL samples: rows from every rank's shard, distances from every rank's tile rows.  The
            #     profiles are a seeded set, so rank 0 regenerates the rows it does not hold.
            Nb = min(N, 384)
            sub = np.sort(np.random.default_rng(17).choice(N, Nb, replace=False))
            Xb = np.zeros((Nb, n))
            indptr, indices, values = synth.synth_profile_rows(leaves, n, [int(s) for s in sub], seed=1)
            for i in range(Nb):
                Xb[i, indices[indptr[i]:indptr[i + 1]]] = values[indptr[i]:indptr[i + 1]]
            _, _, _, Dref_b = cpu_oracle_run(inp, Xb)
            err_b, cnt_b = rel_err_block(D_host[np.ix_(sub, sub)], Dref_b)
            owners = sorted({sharded.tile_row_owner(int(s) // 128, world) for s in sub})
            parity.update(random_block={"max_rel_err_vs_oracle": err_b, "pairs_checked": cnt_b,
                                        "samples": f"{Nb} drawn over all {N}", "tile_row_owners_covered": owners})
            assert err_b < 1e-6, f"GPU result differs from the oracle on the random block: max rel err {err_b}"
            assert len(owners) == world
            err, cnt = err_b, cnt_b
            if world == 1:
                # (b) the CPU baseline's own block: the first 3,072 samples
                Ns = min(N, CPU_SAMPLE)
                cpu_value, cpu_s, threads, Dref = cpu_oracle_run(inp, X_pin.numpy()[:Ns])
                err, cnt = rel_err_block(D_host[:Ns, :Ns], Dref)
                assert err < 1e-6, f"GPU result differs from the oracle: max rel err {err}"
                sample = (f"first {Ns} of the {N} synthetic samples ({Ns * (Ns - 1) // 2} pairs), push-up + all pairs, "
                          f"float64, {cpu_s:.2f} s")
                out["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                                       "sample": sample, "host_cpus": os.cpu_count(),
                                       "note": "C/OpenMP restatement on all cores; the reference's own pairwise loop is "
                                               "single-core Python/numpy by construction (~3.2e4 pairs/s, SURVEY.md par. 6)"}
            parity.update(max_rel_err_vs_oracle=max(err, err_b), pairs_checked=cnt + (cnt_b if world == 1 else 0))
            out["parity"] = parity
            assert e2e_diff < 1e-12, e2e_diff

        lane_ops = 2.0 * pairs * n                      # one subtract + one |.|-accumulate per (pair, entry)
        fp32_peak = env.sm_count * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12      # Tlane-op/s at the max SM clock
        achieved = lane_ops / world / (ms_pair_max * 1e-3) / 1e12         # per GPU (each rank runs 1/world of the tiles)
        default_workload = world == 1 and N == 10000
        traffic_pair = profile_traffic(PROFILE_PAIRWISE) if default_workload else None
        traffic_push = profile_traffic(PROFILE_PUSHUP) if default_workload else None
        push_alg = (hi - lo) * n * (4 + 4)              # N n (4 in + 4 out)
        push_req = (hi - lo) * n * job.pushup_bytes_per_entry(8)   # bytes this job's push-up has to move (float64 in)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic full KO tree (n={n} nodes, 26000 leaves), {N} samples, weighted L1, "
                                   f"median_abund-style integer abundances normalised in float64",
                       "accumulation": "fp32 partial sums over 256 basis entries, combined in float64; operands centred "
                                       "before rounding to fp32; float64 refinement pass over uncertified pairs",
                       "pairs_refined_in_float64": refined,
                       "parallelism": f"samples and 128-row tile rows sharded over {world} GPU(s)",
                       "l2_flush": "none needed: per step 2.4 GB of profiles + 1.2 GB pushed + 0.8 GB matrix >> 126 MB L2",
                       "seed": {"tree": 0, "profiles": 1}},
            "roofline": {"bound": "fp32", "kernel": "pairwise_tile_kernel<L1>", "achieved": achieved, "peak": fp32_peak,
                         "unit": "Tlane-op/s", "frac": achieved / fp32_peak, "traffic": traffic_pair,
                         "traffic_source": f"profiles/{PROFILE_PAIRWISE} (ncu --set full, same workload; the kernel is "
                                           "FP32-pipe bound: 2.0 GB algorithmic)",
                         "ms_per_launch": ms_pair_max, "peak_source": f"{env.sm_count} SMs x 128 FP32 lanes x "
                         f"{peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} max SM clock)",
                         "algorithmic_work": "2 FP32 lane-ops per (pair, basis entry)"},
            "roofline_pushup": {"bound": "hbm", "kernel": "push-up kernels (chunk + span)",
                                "achieved": push_req / (ms_push_max * 1e-3) / 1e9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": push_req / (ms_push_max * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                "bytes_per_entry": job.pushup_bytes_per_entry(8),
                                "frac_survey_8Nn": push_alg / (ms_push_max * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                "traffic": traffic_push,
                                "traffic_source": f"profiles/{PROFILE_PUSHUP}",
                                "ms_per_launch": ms_push_max, "peak_source": peaks["source"]},
            "kernel_share": {"pairwise": ms_pair_max / ms_step, "pushup": ms_push_max / ms_step},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(out)
    del D, Dh_out, X_dev, job
    if world > 1:
        env.barrier()
        shared.close()
    ctx.close()
    torch.cuda.empty_cache()
    return line


def run_cli(env, args):
    """The whole user-visible run through the drop-in CLI with gather CSVs on disk (BASELINE config 2 at --samples 1000;
    any size): argument parsing -> tree -> native ingest into page-locked memory -> push-up + all pairs -> np.save, checked
    against the oracle on the same files.  One GPU."""
    import tempfile as tf

    import importlib.util
    spec = importlib.util.spec_from_file_location("fuf_config_runs", os.path.join(ROOT, "tests", "test_gpu_configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    run_config2 = mod.run_config2
    N = 1000 if args.samples == 10000 and not args.cli_full else args.samples
    with tf.TemporaryDirectory() as tmp:
        res = run_config2(N, tmp)
    assert res["max_rel_err_vs_oracle"] < 1e-6, res
    res.update(metric="sample-pairs/sec through the CLI, files on disk to .npy on disk", value=res["pairs_per_second_whole_cli"],
               unit=UNIT, n_gpus=1, higher_is_better=True, data="synthetic gather CSVs on disk")
    return res


PROFILE_PAIRWISE = "r2_ncu_full_pairwise.csv"
PROFILE_PUSHUP = "r2_ncu_full_pushup.csv"
PROFILE_GRAM = "r2_ncu_full_gram_pair_N50000.csv"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="--workload l2: skip the host-buffer leg (32 GB of pinned memory)")
    ap.add_argument("--cli-full", action="store_true", help="--workload cli: use --samples as given (default 1,000 files)")
    ap.add_argument("--l2-samples", type=int, default=50000)
    ap.add_argument("--workload", default="all", choices=["all", "l1", "l2", "diffab", "cli"],
                    help="all (default): the headline L1 line (10K samples, weighted L1) carrying workloads.l2 (BASELINE "
                         "config 4: --L2 tensor-core Gram path, 50,000 samples) and workloads.diffab (config 5: 500 samples, "
                         "per-pair difference vectors streamed to pinned host memory); l1 / l2 / diffab: that line alone "
                         "(--samples applies); cli: config 2 through the CLI with files on disk")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args)
    try:
        if args.workload == "cli":
            line = run_cli(env, args) if env.rank == 0 else None
        elif args.workload == "l2":
            line = run_l2(env, args, N=args.samples if args.samples != 10000 else args.l2_samples)
        elif args.workload == "diffab":
            line = run_diffab(env, args, N=500 if args.samples == 10000 else args.samples)
        else:
            line = run_l1(env, args)
            if args.workload == "all":
                # the other two north_star workloads ride on the same line (bounded step counts: the default run must
                # finish within minutes); their own metric / roofline / e2e / parity / cpu_baseline inside
                extra = {"l2": run_l2(env, args, N=args.l2_samples, steps=max(1, min(args.steps, 5))),
                         "diffab": run_diffab(env, args, N=500, steps=max(1, min(args.steps, 3)))}
                if line is not None:
                    line["workloads"] = extra
                    line["gpu_launches_all_workloads"] = int(line["gpu_launches"] + sum(w["gpu_launches"] for w in extra.values()))
        if env.rank == 0 and line is not None:
            emit(line)
    finally:
        env.close()


if __name__ == "__main__":
    main()
