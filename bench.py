#!/usr/bin/env python
"""Benchmark of the ANNCUR test-time search path (score + top-100) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload n1m|c2|c4] [--precision f32r|f32x3|bf16]
                    [--shard items|queries] [--exchange p2p|nccl|allgather]

One JSON line on stdout (rank 0).  A *step* is one pass of the hot path over one batch of B queries:
``CURApprox.topk_in_row`` = ``torch.topk(Q @ E, k, dim=1)`` (eval/matrix_approx_zeshel.py:109-126 of the
reference).  Workloads (BASELINE.json):

  n1m  N = 1 000 000 items, k_i = 500, B = 4096, top-100   <- default: the size north_star quotes its target on
  c2   N = 100 000 items,   k_i = 500, B = 4096, top-100      (configs[1]; carried as an extra key of the default run at N = 1)
  c4   N = 10 000 000 items, k_i = 500, B = 4096, top-100     (configs[3]; carried as an extra key of every default run)
  (c3  adaptive multi-round ANNCUR, N = 100 000, 4 rounds x 125 anchors (configs[2]) is an extra key of the default run at N = 1)

N = 1: the whole index on one GPU.  N > 1 (torchrun, one rank per GPU): **items sharded** -- rank p holds E[:, lo_p:hi_p],
every rank scores the same batch against its slice, then ONE exchange step: each row's P candidate lists travel to the rank
that owns the row (contiguous row blocks) and are merged there.  ``--exchange p2p`` (default): keys are stored straight into
the owner's buffer over NVLink peer memory by our own kernel (csrc/peer_exchange.cu); ``nccl``: all_to_all_single;
``allgather``: all-gather + merge of every row on every rank (the form north_star words).  The exchange is inside both timed
regions; total work is fixed as N grows (``scaling: "strong"``).  In every multi-GPU run rank 0 also builds the whole index
and asserts that the sharded answer equals the single-GPU answer on a row sample.  ``--shard queries`` = replicas of the
index, no collective (weak scaling; kept for comparison).

``value``  : device-resident throughput (queries already in HBM), CUDA events, max over ranks; N = 1 one stream in order, N > 1
             one stream or two rotating streams (the exchange of batch j overlaps the kernels of batch j + 1), whichever an
             untimed calibration pass finds faster on this box (``device_streams``); the other figure is printed beside it as
             ``value_with_1_stream`` / ``value_with_2_streams``.
``e2e``    : the same through the host-facing call: N = 1 ``anncur_search_host`` (C ABI: pinned host Q -> H2D -> kernels ->
             D2H of values + indices every step); N > 1 ``ShardedIndex.search_owned``: every rank uploads ITS block of the
             batch, the blocks are all-gathered over NVLink, and every rank downloads the merged rows it owns.
``roofline``: dominant kernel (fused tcgen05 score + top-k = MAIN), per-launch CUDA events recorded inside the library on
             the launching stream; algorithmic flops 2*B*k_i*N_local counted once; peak = MEASURED_PEAKS.json (burst figure
             unless the timed loop is >= 1 s or ran power-capped below 85 % of the max SM clock; both fractions are printed).  ``step_frac`` is
             the same ratio for the whole step (all kernels + exchange).
``cpu_baseline`` / ``--impl reference``: the reference's OWN ``CURApprox.topk_in_row`` from oracle/_ref (a verbatim copy of
             eval/matrix_approx_zeshel.py made by oracle/build_ref.py; kind "reference") on all host threads, or the oracle
             port when oracle/_ref is absent (kind "port") -- the only places bench executes oracle/.
"""
import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #        N items,   k_i, n_train, B,    k
    "c2": (100_000, 500, 2000, 4096, 100),
    "n1m": (1_000_000, 500, 2000, 4096, 100),
    "c4": (10_000_000, 500, 2000, 4096, 100),
}
DEFAULT_STEPS = {"c2": 200, "n1m": 200, "c4": 30}
RANK_LOW, NOISE = 64, 0.05
N_BATCHES = 4            # distinct query batches rotated through the steps
DEV_STREAMS = 2          # multi-GPU device-resident loop: rotating streams (see main)
CHECK_ROWS = 256         # rows of the sharded == single-GPU assertion


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) are sent to stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields of the profiling recipe, read through NVML)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting", 0x10: "sync_boost",
    }

    def __init__(self, dev_index, period_s=0.025):        # 40 Hz: three NVML queries per sample, and NVML polling at 200 Hz
        # showed up as a 4 % slower timed loop on some boxes (the same loop without the sampler ran at the side-measurement rate)
        self.samples, self.reason_bits, self.power = [], 0, []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = dev_index
            if vis:
                toks = [t.strip() for t in vis.split(",") if t.strip()]
                if dev_index < len(toks) and toks[dev_index].isdigit():
                    phys = int(toks[dev_index])
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:                                     # no NVML: report that, do not fake numbers
            self._nv = None
            self.err = str(exc)

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self._period)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(2.0)

    def summary(self):
        if self._nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable or region too short"}
        reasons = sorted(name for bit, name in self.REASONS.items() if self.reason_bits & bit)
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------
# synthetic workload: A = X.Y^T/sqrt(r) + noise*G; only R = A_train (its anchor columns C) and
# Q = A_test[:, anchors] are ever formed.  Plain torch on `device` (cuda or cpu), chunk-keyed seeds, so that any item range
# of the same workload can be regenerated anywhere (a shard on its rank, the whole index on rank 0, the reference arm).
# ---------------------------------------------------------------------------------------------------
CHUNK = 250_000


class Synthetic:
    def __init__(self, name, device, seed=0):
        import numpy as np
        import torch
        self.torch = torch
        self.name, self.device, self.seed = name, device, seed
        self.N, self.k_i, self.n_train, self.B, self.k = WORKLOADS[name]
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.X_train = torch.randn((self.n_train, RANK_LOW), generator=g, device=device)
        self.anc = torch.as_tensor(np.sort(np.random.default_rng(seed).choice(self.N, size=self.k_i, replace=False)), device=device)
        self.chunks = [(a, min(a + CHUNK, self.N)) for a in range(0, self.N, CHUNK)]
        # anchor-item factors / columns: taken from the chunk they live in, so that C == R[:, anchors] exactly
        self.C = torch.empty((self.n_train, self.k_i), device=device)
        self.Y_anc = torch.empty((self.k_i, RANK_LOW), device=device)
        for a, b in self.chunks:
            sel = (self.anc >= a) & (self.anc < b)
            if not bool(sel.any()):
                continue
            Y = self._item_factors(a, b)
            Rc = self._rows_chunk(Y, a, b)
            self.C[:, sel] = Rc[:, self.anc[sel] - a]
            self.Y_anc[sel] = Y[self.anc[sel] - a]
            del Rc, Y

    def _item_factors(self, a, b):
        gy = self.torch.Generator(device=self.device)
        gy.manual_seed(self.seed * 1_000_003 + a)
        return self.torch.randn((b - a, RANK_LOW), generator=gy, device=self.device)

    def _rows_chunk(self, Y, a, b):
        gn = self.torch.Generator(device=self.device)
        gn.manual_seed(self.seed * 7_000_003 + a * 31 + 1)
        return self.X_train @ Y.t() / math.sqrt(RANK_LOW) + self.torch.randn((self.n_train, b - a), generator=gn, device=self.device) * NOISE

    def anchor_rows(self, lo, hi):
        """R[:, lo:hi] = the anchor queries' exact scores of items lo..hi, chunk by chunk: yields (a2, b2, R_chunk)."""
        for a, b in self.chunks:
            a2, b2 = max(a, lo), min(b, hi)
            if a2 >= b2:
                continue
            Rc = self._rows_chunk(self._item_factors(a, b), a, b)[:, a2 - a:b2 - a].contiguous()
            yield a2, b2, Rc

    def query_batch(self, j, salt=0):
        gq = self.torch.Generator(device=self.device)
        gq.manual_seed(self.seed * 13 + 1000 + j + salt)
        Xq = self.torch.randn((self.B, RANK_LOW), generator=gq, device=self.device)
        Q = Xq @ self.Y_anc.t() / math.sqrt(RANK_LOW) + self.torch.randn((self.B, self.k_i), generator=gq, device=self.device) * NOISE
        return Q.contiguous()


def build_index(syn, lo, hi):
    """E[:, lo:hi] with the engine's own K1 (pinv) and K2 (U . R) -- outside every timed region."""
    import torch
    from anncur_b200 import engine
    dev = syn.device
    t0 = time.perf_counter()
    U = engine.pinv(syn.C)                                                   # K1: k_i x n_train
    torch.cuda.synchronize(dev)
    t_pinv = time.perf_counter() - t0
    E = torch.empty((syn.k_i, hi - lo), device=dev)
    t_gemm = 0.0
    for a2, b2, Rc in syn.anchor_rows(lo, hi):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        E[:, a2 - lo:b2 - lo] = engine.gemm(U, Rc)                            # K2: E = U . R
        torch.cuda.synchronize(dev)
        t_gemm += time.perf_counter() - t0
        del Rc
    return E, {"pinv": t_pinv, "U@R": t_gemm}


def cpu_sample_rows(N, B):
    """Bounded CPU sample: rows of one batch such that one pass is ~1e12 flop (seconds on a host's cores)."""
    return int(max(64, min(B, 1_000_000_000 // N)))


def reference_topk_fn():
    """(callable(Q, E, k) -> topk, kind): the reference's own CURApprox.topk_in_row when oracle/_ref is there, else the port."""
    from oracle.ref_shim import load_reference_curapprox
    CUR = load_reference_curapprox()
    if CUR is None:
        from oracle import cur_oracle as O
        return O.score_topk, "port"

    def run(Q, E, k):
        ap = CUR.__new__(CUR)                  # E was built outside: attach it as the constructor would (:63-65)
        ap.latent_cols, ap.approx_preference = E, "rows"
        return ap.topk_in_row(Q, k)            # eval/matrix_approx_zeshel.py:121-126, verbatim
    return run, "reference"


def cpu_topk_throughput(fn, E_host, Q_host, k, repeats, warmup=1):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        fn(Q_host, E_host, k)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return Q_host.shape[0] / med, times


def pin_rank_to_cores(local_rank, world):
    """One process per GPU: give each rank its own slice of the cores NVML reports as local to its GPU (all ranks of a box
    usually report the same set -- then the set is split evenly), so that the launch threads of the ranks do not migrate
    over each other.  Returns the core list, or None when affinity cannot be set."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
    except Exception:
        local = []
    allowed = sorted(os.sched_getaffinity(0))
    cores = [c for c in local if c in allowed] or allowed
    per = max(1, len(cores) // world)
    mine = cores[(local_rank * per) % len(cores):(local_rank * per) % len(cores) + per] or cores
    try:
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only).

    Same synthetic workload as the GPU arm (same generator code and seeds; generated with torch on the GPU when one is
    visible -- plain torch ops, none of our kernels -- else on the CPU).  The index is built the reference's way: its own
    constructor (np.linalg.pinv + U @ R, eval/matrix_approx_zeshel.py:21-69) when R fits in host memory, its formulas chunk by
    chunk otherwise.  Timed: CURApprox.topk_in_row on a bounded row sample of one batch."""
    if rank != 0:
        return
    import numpy as np
    import torch
    N, k_i, n_train, B, k = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count() or 1)
    fn, kind = reference_topk_fn()
    gen_dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    syn = Synthetic(args.workload, gen_dev, seed=0)
    rows = cpu_sample_rows(N, B)
    t0 = time.perf_counter()
    from oracle.ref_shim import load_reference_curapprox
    CUR = load_reference_curapprox()
    C_host = syn.C.cpu()
    if CUR is not None and N <= 2_000_000:
        R_host = torch.empty((n_train, N))
        for a2, b2, Rc in syn.anchor_rows(0, N):
            R_host[:, a2:b2] = Rc.cpu()
        ap = CUR(rows=R_host, cols=C_host, row_idxs=list(range(n_train)), col_idxs=syn.anc.cpu().tolist(), approx_preference="rows")
        E = ap.latent_cols
        built = "reference constructor (np.linalg.pinv + U @ R)"
        del R_host
    else:
        U = torch.from_numpy(np.linalg.pinv(C_host.numpy()))                    # :49
        E = torch.empty((k_i, N))
        for a2, b2, Rc in syn.anchor_rows(0, N):
            E[:, a2:b2] = U @ Rc.cpu()                                          # :65
        built = "np.linalg.pinv + U @ R chunk by chunk (the reference's formulas :49, :65)"
    t_build = time.perf_counter() - t0
    Q = syn.query_batch(0)[:rows].cpu()
    del syn
    for _ in range(args.warmup):
        fn(Q, E, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(Q, E, k)
    dt = time.perf_counter() - t0
    qps = rows * args.steps / dt
    sample = (f"{rows} of the {B} queries of a batch per step, all {N} items, k_i={k_i}, top-{k}; "
              f"{'the reference CURApprox.topk_in_row (oracle/_ref)' if kind == 'reference' else 'oracle port'}: torch.matmul + torch.topk fp32; "
              f"index built by {built} in {t_build:.1f} s (untimed)")
    shard = args.shard or "items"
    line = {
        "impl": "reference", "metric": "queries/sec (ANNCUR score+top-100)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if shard == "items" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        # same config object as the GPU arm prints for these flags (the arm itself runs on the host threads of rank 0)
        "config": workload_config(args, world, shard),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world, shard):
    N, k_i, n_train, B, k = WORKLOADS[args.workload]
    if world == 1:
        par = "1 GPU: the whole index on one device"
    elif shard == "items":
        how = {"p2p": "candidate keys stored into the row owner's buffer over NVLink peer memory (own kernel) + merge of owned rows",
               "nccl": "NCCL all_to_all_single of candidate keys by row block + merge of owned rows",
               "allgather": "NCCL all-gather of candidate keys + merge of every row on every rank"}[args.exchange]
        par = f"items sharded over {world} ranks (E[:, N/{world}] per GPU), same batch on every rank; exchange: {how}"
        if args.exchange == "p2p" and getattr(args, "local_k", "full") != "full":
            par += "; rank-budgeted: each shard ships its best k/P + 6 sigma + 8 candidates per row, every merged row certified exact on the device (full-k fallback)"
    else:
        par = f"dp{world}: E replicated, queries sharded, no collective"
    return {"workload": f"{args.workload}: ANNCUR score+top-{k}, N={N} items, k_i={k_i}, batch {B} queries/step",
            "n_items": N, "k_i": k_i, "batch": B, "top_k": k, "precision": args.precision, "parallelism": par,
            "l2": f"inputs larger than L2: the packed index is streamed every step ({2 * k_i * N / world / 1e6:.0f} MB per rank "
                  f"for the one-pass kinds) and {N_BATCHES} distinct query batches rotate"}


# ---------------------------------------------------------------------------------------------------
class Harness:
    """One workload on this rank: index (whole or item slice), query batches, the step functions."""

    def __init__(self, args, name, rank, local_rank, world, shard):
        import torch
        import torch.distributed as dist
        from anncur_b200 import engine
        from anncur_b200.sharded import ShardedIndex, shard_bounds
        self.torch, self.dist, self.engine = torch, dist, engine
        self.args, self.name, self.rank, self.world, self.shard = args, name, rank, world, shard
        self.device = torch.device("cuda", local_rank)
        self.N, self.k_i, self.n_train, self.B, self.k = WORKLOADS[name]
        self.sharded = shard == "items" and world > 1
        self.syn = Synthetic(name, self.device, seed=0)
        self.lo, self.hi = shard_bounds(self.N, world)[rank] if self.sharded else (0, self.N)
        self.E, self.build_s = build_index(self.syn, self.lo, self.hi)
        t0 = time.perf_counter()
        self.packed = engine.PackedItems(self.E, args.precision)
        torch.cuda.synchronize()
        self.build_s["pack"] = time.perf_counter() - t0
        salt = 0 if (self.sharded or world == 1) else 97 * rank                       # replicas: every rank its own queries
        self.batches = [self.syn.query_batch(j, salt) for j in range(N_BATCHES)]
        self.index = None
        self.local_k = None
        if self.sharded:
            from anncur_b200.sharded import suggest_local_k
            self.index = ShardedIndex(self.E, self.lo, self.N, precision=args.precision, packed=self.packed,
                                      exchange="nccl" if args.exchange == "allgather" else args.exchange)
            self.row_lo, self.row_hi = self.index.row_block(self.B)
            if args.exchange == "p2p" and args.local_k != "full":
                lk = suggest_local_k(self.k, world) if args.local_k == "auto" else int(args.local_k)
                self.local_k = lk if lk < self.k else None
        self.out_v = torch.empty((self.B, self.k), dtype=torch.float32, device=self.device)
        self.out_i = torch.empty((self.B, self.k), dtype=torch.int64, device=self.device)

    # one step with the batch already in HBM
    def step_device(self, j):
        Q = self.batches[j % N_BATCHES]
        if self.index is None:
            return self.engine.score_topk(Q, self.packed, self.k, idx_offset=self.lo, out=(self.out_v, self.out_i))
        if self.args.exchange == "allgather":
            return self.index.search(Q, self.k)
        return self.index.search_rowblock(Q, self.k, self.local_k)

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def units_per_step(self):
        return self.B * (self.world if (self.world > 1 and not self.sharded) else 1)

    def time_device(self, steps, warmup, streams=1, sample_clocks=False, local_rank=0):
        """(ms_total max over ranks, clocks summary | None, MAIN ms sum, MAIN launches, launches counted)."""
        torch, engine = self.torch, self.engine
        pool = [torch.cuda.current_stream(self.device)] if streams == 1 else [torch.cuda.Stream(device=self.device) for _ in range(streams)]

        def issue(j):
            if streams == 1:
                self.step_device(j)
            else:
                with torch.cuda.stream(pool[j % streams]):
                    self.step_device(j)
        for j in range(max(warmup, 2 * streams)):
            issue(j)
        self.sync_all()
        engine.profile_enable(True)
        engine.profile_read()
        engine.reset_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        self.sync_all()
        ev0.record()
        if streams > 1:
            for s in pool:
                s.wait_event(ev0)
        for j in range(steps):
            issue(j)
        if streams > 1:
            for s in pool:
                torch.cuda.current_stream(self.device).wait_stream(s)
        ev1.record()
        self.sync_all()
        if sampler:
            sampler.__exit__()
        ms_total = self.max_over_ranks(ev0.elapsed_time(ev1))
        launches = engine.launch_count()
        fused_ms, fused_n = engine.profile_read()
        engine.profile_enable(False)
        return ms_total, (sampler.summary() if sampler else None), fused_ms, fused_n, launches

    def close(self):
        if self.index is not None:
            self.index.close()
        self.engine.WORKSPACE.clear()
        for name in ("E", "packed", "batches", "index", "syn", "out_v", "out_i"):
            setattr(self, name, None)
        self.torch.cuda.empty_cache()


def quick_extra(args, name, rank, local_rank, world, shard, steps):
    """Device-resident figure of another BASELINE config inside the same run (no e2e / CPU legs)."""
    h = Harness(args, name, rank, local_rank, world, shard)
    ms_total, _, fused_ms, fused_n, _ = h.time_device(steps, 3, streams=DEV_STREAMS if h.index is not None else 1)
    cert = None
    if h.local_k is not None:
        cert = {"local_k": h.local_k, "certificate_failures": h.index.certificate_failures(reset=True)}
        if cert["certificate_failures"]:
            h.local_k = None
            ms_total, _, fused_ms, fused_n, _ = h.time_device(steps, 3, streams=DEV_STREAMS if h.index is not None else 1)
            cert["action"] = "certificate failed: timed again with local_k = k"
    out = {"workload": workload_config(argparse.Namespace(**{**vars(args), "workload": name}), world, shard)["workload"],
           "value": h.units_per_step() * steps / (ms_total * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps,
           "ms_per_step": ms_total / steps, "main_kernel_ms": fused_ms / max(fused_n, 1),
           "parallelism": workload_config(argparse.Namespace(**{**vars(args), "workload": name}), world, shard)["parallelism"],
           "index_build_s": h.build_s}
    if cert is not None:
        out["rank_budgeted_exchange"] = cert
    h.close()
    return out


def run_c3(device, steps=3, B=4096, N=100_000, k_q=500, rounds=4, per_round=125, k=100, world=1, rank=0):
    """BASELINE configs[2]: adaptive multi-round ANNCUR (NOT in the reference, parity unpinned): per query 4
    rounds of 125 anchor items chosen by re-solving e_q = c_q . pinv(R_anc[:, I_t]) and re-scoring all items; the exact-score
    matrix stands in for the cross-encoder calls and stays on the device.  A step = the whole procedure for one batch of B
    queries.  world > 1 (strong scaling of the same batch), two forms timed:
      * queries split over the ranks by row block, every rank holds the whole R_anc (200 MB at this size) -- no collective on the
        data path: the per-query solves are 4/5 of the work and queries are independent.  This is ``value``.
      * the same split of the solves with the RE-SCORE item-sharded and one ShardedIndex.search_owned exchange per round (the form
        SURVEY 8e words; it pays off when R_anc is too large to replicate) -- ``item_sharded_rescore``."""
    import torch
    import torch.distributed as dist
    from anncur_b200 import adaptive_anncur
    from anncur_b200.adaptive import AdaptiveIndex
    from anncur_b200.sharded import ShardedIndex, shard_bounds
    g = torch.Generator(device=device)
    g.manual_seed(0)
    r = RANK_LOW
    Y = torch.randn((N, r), generator=g, device=device)
    R = torch.randn((k_q, r), generator=g, device=device) @ Y.t() / math.sqrt(r) + NOISE * torch.randn((k_q, N), generator=g, device=device)
    X = torch.randn((B, r), generator=g, device=device) @ Y.t() / math.sqrt(r) + NOISE * torch.randn((B, N), generator=g, device=device)
    first = torch.randperm(N, generator=g, device=device)[:per_round].sort().values
    lo, hi = shard_bounds(B, world)[rank]
    Xb = X[lo:hi] if world > 1 else X

    def timed(index, total):
        t0 = time.perf_counter()
        adaptive_anncur(R, Xb, first, rounds, per_round, k, index=index, n_rows_total=total)
        torch.cuda.synchronize()
        first_s = time.perf_counter() - t0            # includes anncur_adaptive_prepare (once per index + first anchors)
        adaptive_anncur(R, Xb, first, rounds, per_round, k, index=index, n_rows_total=total)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0.record()
        for _ in range(steps):
            out = adaptive_anncur(R, Xb, first, rounds, per_round, k, index=index, n_rows_total=total)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, first_s, out

    t0 = time.perf_counter()
    index = AdaptiveIndex(R)                      # packed R_anc + its item-major copy: built once per index, outside the step
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    ms, first_s, (anc, idx, val) = timed(index, None)
    n_chk = min(512, Xb.shape[0])
    exact = torch.topk(Xb[:n_chk], k, dim=1).indices
    recall = (idx[:n_chk].unsqueeze(2) == exact.unsqueeze(1)).any(2).float().mean().item()
    out = {"workload": f"c3: adaptive ANNCUR, N={N} items, k_q={k_q}, {rounds} rounds x {per_round} anchors, batch {B} queries/step, top-{k} by exact score",
           "value": B / (ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "solver": "incremental (anncur_adaptive_begin / _extend: factor carried across rounds) + fused tensor-core re-score",
           "index_build_s": build_s, "first_call_s_incl_prepare": first_s, "recall_at_k_vs_exact": recall,
           "parity": "unpinned (no reference implementation; checked against our own CPU restatement)"}
    if world > 1:
        out["parallelism"] = (f"queries split over {world} ranks by row block ({hi - lo} per rank), R_anc replicated, no collective on the data path; "
                              "strong scaling of one batch")
        sharded = AdaptiveIndex(R, sharded=ShardedIndex.from_full(R))
        ms2, _, (anc2, idx2, _) = timed(sharded, B)
        out["item_sharded_rescore"] = {
            "value": B / (ms2 * 1e-3), "ms_per_step": ms2,
            "parallelism": f"solves split by query block, re-score item-sharded over {world} ranks, one search_owned exchange per round ({sharded.sharded.exchange})",
            "anchors_equal_replicated_form_frac": float((anc2 == anc).float().mean().item()),
            "answer_equal_replicated_form_frac": float((idx2 == idx).float().mean().item())}
        sharded.sharded.close()
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        adaptive_anncur(R, X, first, rounds, per_round, k, index=index, solver="full")
        ev1.record()
        torch.cuda.synchronize()
        out["ms_per_step_full_resolve_every_round"] = ev0.elapsed_time(ev1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="n1m", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f32r", choices=["f32r", "f32x3", "bf16"])
    ap.add_argument("--shard", default=None, choices=["queries", "items"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl", "allgather"])
    ap.add_argument("--local-k", default="auto", help="item-sharded p2p exchange: candidates a shard re-scores and ships per row: 'auto' "
                    "(k/P + 6 sigma + 8, certified on the device, full-k fallback), 'full' (= k), or an integer")
    ap.add_argument("--pin-cores", action="store_true", help="multi-GPU: pin each rank to its own slice of the GPU-local host cores")
    ap.add_argument("--extras", default=None, help="comma list of other workloads measured in the same run (default: c2,c4 at N=1, c4 at N>1; 'none')")
    ap.add_argument("--no-extra", action="store_true", help="skip the other-precision / recall / pipelined side measurements and the extras")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args, rank, world)
        return
    args.steps = args.steps if args.steps is not None else DEFAULT_STEPS[args.workload]
    args.warmup = max(3, args.warmup if args.warmup is not None else 10)
    shard = args.shard or "items"

    import torch
    import torch.distributed as dist
    from anncur_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    # measured on a 4-GPU box (round 2): pinning each rank to its own 1/P slice of the cores made nothing faster and the
    # host-timed legs slightly slower (NCCL's proxy / watchdog threads then share the slice with the launch thread) -> opt-in
    pinned_cores = pin_rank_to_cores(local_rank, world) if args.pin_cores else None
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    N, k_i, n_train, B, k = WORKLOADS[args.workload]

    h = Harness(args, args.workload, rank, local_rank, world, shard)
    packed, batches, index = h.packed, h.batches, h.index
    lo = h.lo

    # ---- device-resident timing (one stream, in order) -------------------------------------------------
    # N = 1: one stream, in order.  N > 1: the same steps issued over DEV_STREAMS rotating streams (each with its own exchange
    # channel), so that a rank waiting for the slowest sender of batch j already runs the kernels of batch j + 1 -- the GPUs of
    # a box drift apart by several % per step under their power caps, and an in-order loop pays the slowest rank every step.
    # Whether that pays depends on the box (8 GPUs: +10 %; a 4-GPU box with a lower power cap: -12 %), so an untimed calibration
    # pass picks 1 or DEV_STREAMS streams (max-over-ranks times, hence the same choice on every rank); the timed region below
    # then runs exactly args.steps steps with that choice.
    dev_streams, stream_calibration = 1, None
    settle_steps = 0
    if index is None:
        # the first loop after the index build runs through a power / clock transient of the capped GPU (observed: 1215 MHz
        # in the first loop, 1320-1400 MHz in every later one); an untimed pass lets the clocks settle before W + K
        settle_steps = max(10, min(100, args.steps))
        h.time_device(settle_steps, 3, streams=1)
    if index is not None:
        n_cal = max(10, min(40, args.steps))
        cal = {ns: h.time_device(n_cal, 3, streams=ns)[0] / n_cal for ns in (1, DEV_STREAMS)}
        dev_streams = min(cal, key=cal.get)
        stream_calibration = {f"ms_per_step_with_{ns}_stream{'s' if ns > 1 else ''}": v for ns, v in cal.items()}
        if h.local_k is not None:
            index.certificate_failures(reset=True)         # calibration rows are not part of the count reported below
    ms_total, clocks, fused_ms, fused_n, launches = h.time_device(args.steps, args.warmup, streams=dev_streams, sample_clocks=True, local_rank=local_rank)
    if index is not None and args.exchange == "p2p" and index.exchange != "p2p":
        args.exchange, h.local_k = "nccl", None           # peer memory was not available on this box: the line says what ran
    cert = None
    if h.local_k is not None:
        # rank-budgeted exchange: every merged row was certified on the device inside the timed region; rows that failed are
        # counted.  A run with failures is NOT reported: the loop is timed again with the full-k exchange.
        n_fail = index.certificate_failures(reset=True)
        cert = {"local_k": h.local_k, "k": k, "rows_certified": h.B * (args.steps + max(args.warmup, 2)), "certificate_failures": n_fail}
        if n_fail:
            h.local_k = None
            ms_total, clocks, fused_ms, fused_n, launches = h.time_device(args.steps, args.warmup, streams=dev_streams, sample_clocks=True, local_rank=local_rank)
            cert["action"] = "certificate failed: timed again with local_k = k"
    qps = h.units_per_step() * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the host-facing call ---------------------------------------------------------
    # Rotating streams, each with its own workspace / exchange channel and pinned result buffers, so that the H2D copy of one
    # batch and the D2H copy of another overlap the kernels of a third (the calls are asynchronous; every step still moves
    # its own query rows in and its own result rows out inside the timed region).
    N_SLOTS = 3          # measured (tools/e2e_probe.py, 200 steps at C2): 1 stream 0.712 ms/step, 2: 0.525, 3: 0.478, 4: 0.492
    rows_lo, rows_hi = (h.row_lo, h.row_hi) if (index is not None and args.exchange != "allgather") else (0, B)
    n_own = rows_hi - rows_lo
    Qh = [b[rows_lo:rows_hi].cpu().pin_memory() for b in batches]
    vh = [torch.empty((n_own, k), dtype=torch.float32).pin_memory() for _ in range(N_SLOTS)]
    ih = [torch.empty((n_own, k), dtype=torch.int64).pin_memory() for _ in range(N_SLOTS)]
    streams = [torch.cuda.Stream(device=device) for _ in range(N_SLOTS)]
    q_stage = [torch.empty((n_own, k_i), dtype=torch.float32, device=device) for _ in range(N_SLOTS)]

    def step_e2e(j):
        s = j % N_SLOTS
        with torch.cuda.stream(streams[s]):
            if index is None:
                engine.search_host(Qh[j % N_BATCHES], packed, k, vh[s], ih[s], idx_offset=lo, ws_key=f"search_host{s}")
                return
            q_stage[s].copy_(Qh[j % N_BATCHES], non_blocking=True)
            if args.exchange == "allgather":        # replicated batch in, full answer out on every rank
                v, i = index.search(q_stage[s], k)
            else:                                   # own block of rows in, NVLink all-gather of the blocks, own rows out
                v, i = index.search_owned(q_stage[s], B, k, h.local_k)
            vh[s].copy_(v, non_blocking=True)
            ih[s].copy_(i, non_blocking=True)

    e2e_steps = args.steps
    for j in range(2 * N_SLOTS):
        step_e2e(j)
    h.sync_all()
    e2e_times = []
    for _ in range(3):                     # host-timed (the copies are part of it): median of three passes of `steps` steps
        t0 = time.perf_counter()
        for j in range(e2e_steps):
            step_e2e(j)
        torch.cuda.synchronize()
        e2e_times.append(h.max_over_ranks(time.perf_counter() - t0))
        h.sync_all()
    t_e2e = statistics.median(e2e_times)
    e2e_qps = h.units_per_step() * e2e_steps / t_e2e
    if cert is not None and h.local_k is not None:
        cert["certificate_failures_e2e"] = index.certificate_failures(reset=True)
    replicated_io = index is not None and args.exchange == "allgather"
    io_ranks = world if (replicated_io or (world > 1 and index is None)) else 1
    h2d = B * k_i * 4 * io_ranks           # bytes over all ranks per step (item-sharded row-block form: each rank moves B/P rows)
    d2h = B * k * (4 + 8) * io_ranks

    # what the box's host link gives a pinned copy of this rank's share of one step's input / output on its own
    def copy_gbs(dst, src):
        best = 0.0
        for _ in range(5):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            dst.copy_(src, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            best = max(best, src.numel() * src.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9)
        return best
    # the host-facing path must give what the device-resident path gives
    last = e2e_steps - 1
    chk_v, chk_i = h.step_device(last)
    torch.cuda.synchronize()
    assert torch.equal(chk_i.cpu(), ih[last % N_SLOTS]) and torch.equal(chk_v.cpu(), vh[last % N_SLOTS]), \
        "host-buffer result differs from the device-resident result"
    h2d_gbs = copy_gbs(q_stage[0], Qh[0])
    d2h_gbs = copy_gbs(ih[0], h.out_i[:n_own])             # (the result buffers are not needed any more)

    # ---- multi-GPU: the sharded answer must equal the single-GPU answer (rank 0 builds the whole index) ----
    sharded_check = None
    if index is not None:
        got_v, got_i = index.search_rowblock_verified(batches[0], k, h.local_k) if args.exchange != "allgather" else index.search(batches[0], k)
        peer_err = [ch.error() for ch in index._channels.values()]
        if rank == 0:
            if N <= 2_000_000:
                E_full, _ = build_index(h.syn, 0, N)
                packed_full = engine.PackedItems(E_full, args.precision)
                rows = min(CHECK_ROWS, got_v.shape[0])
                ref_v, ref_i = engine.score_topk(batches[0][:rows].contiguous(), packed_full, k)
                same_i = bool(torch.equal(ref_i, got_i[:rows]))
                max_dv = float((ref_v - got_v[:rows]).abs().max().item())
                sharded_check = {"rows": rows, "reference": "single-GPU engine.score_topk on the whole index, same precision kind",
                                 "indices_equal": same_i, "max_abs_value_diff": max_dv, "peer_wait_errors": peer_err}
                del E_full, packed_full
                assert same_i and max_dv <= 1e-5 * float(ref_v.abs().max().item()), f"sharded != single-GPU: {sharded_check}"
            else:
                sharded_check = {"rows": 0, "note": "whole index not rebuilt at this size; see the n1m run and tests/test_gpu_sharded_nccl.py",
                                 "peer_wait_errors": peer_err}
            assert not any(peer_err), f"peer exchange wait timed out: {peer_err}"
        h.sync_all()

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak_src = "fallback (B200_PROFILING.md)"
    n_local = h.hi - h.lo
    fused_ms_avg = fused_ms / max(fused_n, 1)
    flops = 2.0 * B * k_i * n_local
    # bytes of E the MAIN kernel streams per launch: both fp16 planes (f32x3), or one 16-bit plane (bf16; f32r streams
    # only the high plane -- the fp32 copy is touched for the ~k re-scored candidates of a row, by the refine kernel)
    e_bytes = (4 if args.precision == "f32x3" else 2) * k_i * n_local
    alg_bytes = e_bytes + 4 * B * k_i + 12 * B * k
    passes = 3 if args.precision == "f32x3" else 1
    tf = flops / (fused_ms_avg * 1e-3) / 1e12 if fused_n else None
    peak_burst = float(peaks.get("bf16_tflops", 1640.0))
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1380.0))
    # Which measured peak the timed region is held against (both fractions are always printed): the burst figure -- a kernel
    # timed alone at full clocks -- unless the loop is long (>= 1 s) or the clocks sampled DURING the loop show the sustained
    # regime MEASURED_PEAKS.json's sustained figure was taken in (sw_power_cap active, median SM clock under 85 % of max).
    capped = bool(clocks) and "sw_power_cap" in (clocks.get("reasons") or []) and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") \
        and clocks["sm_mhz"] < 0.85 * clocks["sm_max_mhz"]
    burst = ms_total < 1000.0 and not capped
    peak_tf = peak_burst if burst else peak_sust
    step_tf = flops / (ms_total / args.steps * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:                                           # dram bytes of the MAIN kernel from an `ncu --set full` capture of this revision
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{args.workload}/{args.precision}/{world}")
        if t:
            traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
    except Exception:
        pass
    roofline = {
        "kernel": "fused_score_topk_kernel (tcgen05 score GEMM + streaming top-k), MAIN launch",
        "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (tf / peak_tf) if tf else None,
        "peak_kind": ("burst" if burst else "sustained") + f" (timed loop {ms_total:.0f} ms"
                     + (f" at {clocks['sm_mhz']:.0f} of {clocks['sm_max_mhz']:.0f} MHz under sw_power_cap" if capped else "") + ")", "peak_source": peak_src,
        "frac_of_burst": (tf / peak_burst) if tf else None, "frac_of_sustained": (tf / peak_sust) if tf else None,
        "step_achieved": step_tf, "step_frac": step_tf / peak_tf, "step_frac_of_burst": step_tf / peak_burst,
        "step_frac_of_sustained": step_tf / peak_sust,
        "traffic": traffic, "traffic_source": traffic_src,
        "launch_ms": fused_ms_avg, "launches_timed": fused_n, "share_of_step": fused_ms / ms_total if ms_total else None,
        "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes,
        "tensor_passes": passes,
        "tensor_pipe_tflops": (tf * passes) if tf else None,
        "note": "achieved / step_achieved count the algorithmic flops 2*B*k_i*N_local of THIS rank once (per-GPU figures); kind f32x3 "
                "issues 3 f16 tensor passes per product, kind f32r one pass (score upper bounds) + fp32 re-scoring of ~k "
                "candidates per row in refine_topk_kernel (inside the step, not inside launch_ms)",
        "hbm_gbs_achieved": alg_bytes / (fused_ms_avg * 1e-3) / 1e9 if fused_n else None,
        "hbm_gbs_peak": peaks.get("hbm_gbs"),
    }

    line = {
        "metric": "queries/sec (ANNCUR score+top-100)", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if (shard == "items") else "weak", "vs_baseline": None,
        "dtype": {"f32x3": "f32 (2xfp16 split operands, 3 tcgen05 passes, fp32 accumulate)", "bf16": "bf16",
                  "f32r": "f32 (one f16 tcgen05 pass of rigorous score upper bounds, fp32 FFMA re-scoring of the candidates)"}[args.precision],
        "data": "synthetic", "config": workload_config(args, world, shard),
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "anncur_search_host (C ABI, pinned host buffers)" if index is None else
                       ("ShardedIndex.search + pinned copies of the whole batch / answer on every rank" if replicated_io else
                        "ShardedIndex.search_owned: every rank uploads its B/P rows, NVLink all-gather of the rows, search, "
                        "exchange, download of the B/P merged rows it owns"),
                "passes_s": e2e_times, "h2d_gbs_alone_this_rank": h2d_gbs, "d2h_gbs_alone_this_rank": d2h_gbs,
                "streams": N_SLOTS},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "index_build_s": h.build_s,
    }
    if pinned_cores is not None:
        line["host_cores_of_rank0"] = pinned_cores
    if sharded_check is not None:
        line["sharded_equals_single_gpu"] = sharded_check
    if cert is not None:
        line["rank_budgeted_exchange"] = cert
    if index is not None and args.exchange != "allgather":
        # where the step goes on this rank: local search (pack .. refine / REDO) against exchange (scatter, wait for the
        # slowest sender, merge), CUDA events on the launching stream, 50 steps
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(50)]
        h.sync_all()
        for j, (a0, a1, a2) in enumerate(evs):
            Qj = batches[j % N_BATCHES]
            k_loc = h.local_k or k
            a0.record()
            lv, li = index.local_topk(Qj, k_loc)
            a1.record()
            if index.exchange == "p2p":
                index._channel(B, k_loc, lv.device).exchange(lv, li, k_out=k)
            else:
                index.search_rowblock(Qj, k)                      # (NCCL form: includes a second local search; see local_ms)
            a2.record()
        h.sync_all()
        loc = sum(a0.elapsed_time(a1) for a0, a1, _ in evs) / len(evs)
        exc = sum(a1.elapsed_time(a2) for _, a1, a2 in evs) / len(evs)
        if index.exchange != "p2p":
            exc -= loc
        line["step_breakdown_ms"] = {"local_search_max_over_ranks": h.max_over_ranks(loc), "local_search_min_over_ranks": -h.max_over_ranks(-loc),
                                     "exchange_max_over_ranks": h.max_over_ranks(exc), "exchange_min_over_ranks": -h.max_over_ranks(-exc),
                                     "note": "exchange = key scatter + wait for the slowest sender + merge of the owned rows"}
        if h.local_k is not None:
            index.certificate_failures(reset=True)
    if index is None:
        # rows of the last device-resident batch that needed the fallback pass (0 = the fast path served the whole batch)
        engine.score_topk(batches[0], packed, k, idx_offset=lo, out=(h.out_v, h.out_i))
        line["redo_rows_last_batch"] = engine.last_redo_rows(B, packed, k)

    # ---- side measurements ---------------------------------------------------------------------------------
    line["device_streams"] = dev_streams
    line["settle_steps_before_warmup"] = settle_steps
    if stream_calibration is not None:
        line["device_streams_calibration"] = stream_calibration
    if not args.no_extra:
        # the same steps over other stream counts (1 = strictly in order: every step pays the slowest rank of that step)
        for ns in [n for n in (1, 2, 3) if n != dev_streams]:
            ms2, _, _, _, _ = h.time_device(args.steps, 4, streams=ns)
            line[f"value_with_{ns}_stream{'s' if ns > 1 else ''}"] = {"value": h.units_per_step() * args.steps / (ms2 * 1e-3), "unit": "queries/s",
                                                                  "ms_per_step": ms2 / args.steps}
        if cert is not None and h.local_k is not None:
            cert["certificate_failures_other_stream_counts"] = index.certificate_failures(reset=True)
            # the same loop with every shard shipping all k candidates (no budget, no certificate needed)
            lk, h.local_k = h.local_k, None
            ms3, _, _, _, _ = h.time_device(args.steps, 4, streams=dev_streams)
            h.local_k = lk
            line["value_with_local_k_equal_k"] = {"value": h.units_per_step() * args.steps / (ms3 * 1e-3), "unit": "queries/s",
                                                  "ms_per_step": ms3 / args.steps}
    if world == 1 and not args.no_extra:
        line["other_precision"] = []
        v_a, i_a = engine.score_topk(batches[0], packed, k)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for other in [p for p in ("f32r", "f32x3", "bf16") if p != args.precision]:
            packed_o = engine.PackedItems(h.E, other)
            for j in range(3):
                engine.score_topk(batches[j % N_BATCHES], packed_o, k, out=(h.out_v, h.out_i))
            torch.cuda.synchronize()
            n_o = max(10, args.steps // 4)
            ev0.record()
            for j in range(n_o):
                engine.score_topk(batches[j % N_BATCHES], packed_o, k, out=(h.out_v, h.out_i))
            ev1.record()
            torch.cuda.synchronize()
            qps_o = B * n_o / (ev0.elapsed_time(ev1) * 1e-3)
            v_b, i_b = engine.score_topk(batches[0], packed_o, k)
            inter = (i_a.unsqueeze(2) == i_b.unsqueeze(1)).any(dim=2).float().sum(dim=1) / k
            line["other_precision"].append({"precision": other, "value": qps_o, "unit": "queries/s",
                                            "recall_at_k_vs_" + args.precision: float(inter.mean().item())})
            del packed_o
    if world == 1 and rank == 0 and not args.no_cpu:
        rows = cpu_sample_rows(N, B)
        fn, kind = reference_topk_fn()
        E_host = h.E.cpu()
        Q_host = batches[0][:rows].cpu()
        cpu_qps, times = cpu_topk_throughput(fn, E_host, Q_host, k, repeats=3)
        # while the CPU result is at hand: the GPU answer for the same rows must be the reference's answer
        ref = fn(Q_host, E_host, k)
        got_v, got_i = engine.score_topk(batches[0][:rows].contiguous(), packed, k)
        dense_scale = ref.values.abs().max(dim=1, keepdim=True).values
        same = (got_i.cpu() == ref.indices).float().mean().item()
        rel = ((got_v.cpu() - ref.values).abs() / dense_scale).max().item()
        line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": f"{rows} of the {B} queries of one batch x all {N} items (same E and Q as the GPU arm), median of 3 "
                                          f"after 1 warm-up ({statistics.median(times):.2f} s each); "
                                          + ("the reference's CURApprox.topk_in_row (oracle/_ref)" if kind == "reference" else "oracle port")
                                          + ": torch.matmul + torch.topk fp32",
                                "gpu_vs_cpu_check": {"index_agreement": same, "max_rel_score_err_sorted_lists": rel}}
        del E_host

    # ---- the HBM-bound regime of the same index: small batches stream E once per call (north_star: "HBM GB/s ... for small-batch
    #      streaming of E") -- device-resident, one stream, the MAIN launch timed by the library's own events ----------------
    if world == 1 and not args.no_extra and args.precision in ("f32r", "bf16", "f32x3"):
        try:
            line["small_batch"] = []
            hbm_peak = float(peaks.get("hbm_gbs", 0.0)) or None
            plane_bytes = (4.0 if args.precision == "f32x3" else 2.0) * k_i * N          # what MAIN streams (algorithmic: unpadded k_i)
            for b_small in (1, 64):
                qs = [bt[:b_small].contiguous() for bt in batches[:2]]
                ov = torch.empty((b_small, k), dtype=torch.float32, device=device)
                oi = torch.empty((b_small, k), dtype=torch.int64, device=device)
                for j in range(5):
                    engine.score_topk(qs[j % 2], packed, k, idx_offset=lo, out=(ov, oi))
                torch.cuda.synchronize()
                engine.profile_enable(True)                      # MAIN launch time from a short profiled pass (the library creates an
                engine.profile_read()                            # event pair per launch while profiling: kept out of the timed loop)
                for j in range(50):
                    engine.score_topk(qs[j % 2], packed, k, idx_offset=lo, out=(ov, oi))
                torch.cuda.synchronize()
                main_ms, main_n = engine.profile_read()
                engine.profile_enable(False)
                n_sb = 200
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for j in range(n_sb):
                    engine.score_topk(qs[j % 2], packed, k, idx_offset=lo, out=(ov, oi))
                e1.record()
                torch.cuda.synchronize()
                step_ms = e0.elapsed_time(e1) / n_sb
                main_ms = main_ms / max(main_n, 1)
                # the same call replayed as ONE CUDA graph (engine.GraphedSearch: no launch gaps between its seven kernels)
                gs = engine.GraphedSearch(packed, b_small, k, idx_offset=lo)
                for j in range(5):
                    gs(qs[j % 2])
                torch.cuda.synchronize()
                e0.record()
                for j in range(n_sb):
                    gs(qs[j % 2])
                e1.record()
                torch.cuda.synchronize()
                graph_ms = e0.elapsed_time(e1) / n_sb
                del gs
                line["small_batch"].append({
                    "batch": b_small, "value": b_small / (step_ms * 1e-3), "unit": "queries/s", "ms_per_step": step_ms,
                    "main_kernel_ms": main_ms, "bound": "hbm", "algorithmic_bytes_main": plane_bytes,
                    "hbm_gbs_main": plane_bytes / (main_ms * 1e-3) / 1e9, "hbm_gbs_step": plane_bytes / (step_ms * 1e-3) / 1e9,
                    "hbm_gbs_peak": hbm_peak,
                    "frac_main": (plane_bytes / (main_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                    "frac_step": (plane_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                    "as_one_cuda_graph": {"value": b_small / (graph_ms * 1e-3), "ms_per_step": graph_ms,
                                          "frac_step": (plane_bytes / (graph_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None}})
        except Exception as exc:
            line["small_batch"] = {"error": f"{type(exc).__name__}: {exc}"}
    h.close()

    # ---- the other BASELINE configs, device-resident figure only ---------------------------------------------
    if args.extras is None:
        extras = [] if (args.no_extra or args.workload != "n1m") else (["c2", "c3", "c4"] if world == 1 else ["c3", "c4"])
    else:
        extras = [e for e in args.extras.split(",") if e and e != "none"]
    if extras:
        line["extras"] = {}
        for name in extras:
            try:
                if name == "c3":
                    line["extras"][name] = run_c3(device, world=world, rank=rank)
                    torch.cuda.empty_cache()
                    continue
                line["extras"][name] = quick_extra(args, name, rank, local_rank, world, shard, steps=min(args.steps, 100 if name == "c2" else 20))
            except Exception as exc:                                     # an extra never takes the headline line down
                line["extras"][name] = {"error": f"{type(exc).__name__}: {exc}"}
                if world > 1:
                    raise

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
