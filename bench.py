#!/usr/bin/env python
"""Benchmark of the ANNCUR test-time search path (score + top-100) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|n1m|c4] [--precision f32r|f32x3|bf16] [--shard queries|items]

One JSON line on stdout (rank 0).  A *step* is one pass of the hot path over one batch of B queries:
``CURApprox.topk_in_row`` = ``torch.topk(Q @ E, k, dim=1)`` (eval/matrix_approx_zeshel.py:109-126 of the
reference).  Workloads (BASELINE.json ``configs``):

  c2   N = 100 000 items, k_i = 500, B = 4096, top-100            <- default, configs[1]
  n1m  N = 1 000 000 items, k_i = 500, B = 4096, top-100          (north_star headline size)
  c4   N = 10 000 000 items, k_i = 500, B = 4096, top-100         (configs[3], item-sharded)

``value``  : device-resident throughput (queries already in HBM), CUDA events, max over ranks.
``e2e``    : the same through the host-buffer C-ABI entry (anncur_search_host): pinned host Q -> H2D ->
             kernels -> D2H of (values, indices) every step.
``roofline``: dominant kernel (fused tcgen05 score + top-k), per-launch CUDA events recorded inside the
             library on the launching stream; algorithmic flops 2*B*k_i*N (counted once, also for the 3-pass
             kind), peak = MEASURED_PEAKS.json.  Default precision f32r: fp32 results from ONE f16 tensor pass
             of rigorous score upper bounds + fp32 re-scoring of the ~k candidates per row.
``cpu_baseline`` / ``--impl reference``: the oracle's CPU restatement of the reference path
             (torch.matmul + torch.topk on all host threads) -- the only place bench executes oracle/.

Multi-GPU (torchrun, one rank per GPU): ``--shard queries`` (default for c2/n1m) replicates E and gives
every rank its own query batches -- no data-path collective, weak scaling; ``--shard items`` (default
for c4) splits the items, every rank scores the same batch against its slice and the candidates are
merged after one NCCL all-gather (strong scaling in N).
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
RANK_LOW, NOISE = 64, 0.05
N_BATCHES = 4            # distinct query batches rotated through the steps


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

    def __init__(self, dev_index, period_s=0.005):
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
# synthetic workload: A = X.Y^T/sqrt(r) + noise*G; only R = A_train and Q = A_test[:, anchors]
# are ever formed.  Built with the engine's own K1/K2 (pinv + U.R) outside the timed region.
# ---------------------------------------------------------------------------------------------------
def build_workload(name, device, lo, hi, seed, n_batches, batch_seed_offset=0):
    """Item-embedding slice E[:, lo:hi] (k_i x (hi-lo) fp32 on `device`) of the N-item index + query batches."""
    import numpy as np
    import torch
    from anncur_b200 import engine

    N, k_i, n_train, B, k = WORKLOADS[name]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    r = RANK_LOW
    X_train = torch.randn((n_train, r), generator=g, device=device)
    anc = np.sort(np.random.default_rng(seed).choice(N, size=k_i, replace=False))
    anc_t = torch.as_tensor(anc, device=device)

    def item_factors(a, b):                       # Y[a:b] regenerated per chunk from a chunk-keyed seed
        gy = torch.Generator(device=device)
        gy.manual_seed(seed * 1_000_003 + a)
        return torch.randn((b - a, r), generator=gy, device=device)

    def noise(rows, a, b, salt):
        gn = torch.Generator(device=device)
        gn.manual_seed(seed * 7_000_003 + a * 31 + salt)
        return torch.randn((rows, b - a), generator=gn, device=device) * NOISE

    CH = 250_000
    chunks = [(a, min(a + CH, N)) for a in range(0, N, CH)]
    # anchor-item factors/noise: take them from the chunk they live in so that C == R[:, anc] exactly
    C = torch.empty((n_train, k_i), device=device)
    Y_anc = torch.empty((k_i, r), device=device)
    for a, b in chunks:
        sel = (anc_t >= a) & (anc_t < b)
        if not bool(sel.any()):
            continue
        Y = item_factors(a, b)
        Rc = X_train @ Y.t() / math.sqrt(r) + noise(n_train, a, b, 1)
        C[:, sel] = Rc[:, anc_t[sel] - a]
        Y_anc[sel] = Y[anc_t[sel] - a]
        del Rc, Y
    t0 = time.perf_counter()
    U = engine.pinv(C)                                                   # K1: k_i x n_train
    torch.cuda.synchronize(device)
    t_pinv = time.perf_counter() - t0
    E = torch.empty((k_i, hi - lo), device=device)
    t_gemm = 0.0
    for a, b in chunks:
        a2, b2 = max(a, lo), min(b, hi)
        if a2 >= b2:
            continue
        Y = item_factors(a, b)
        Rc = (X_train @ Y.t() / math.sqrt(r) + noise(n_train, a, b, 1))[:, a2 - a:b2 - a].contiguous()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        E[:, a2 - lo:b2 - lo] = engine.gemm(U, Rc)                        # K2: E = U . R
        torch.cuda.synchronize(device)
        t_gemm += time.perf_counter() - t0
        del Rc, Y
    batches = []
    for j in range(n_batches):
        gq = torch.Generator(device=device)
        gq.manual_seed(seed * 13 + 1000 + j + batch_seed_offset)
        Xq = torch.randn((B, r), generator=gq, device=device)
        Q = Xq @ Y_anc.t() / math.sqrt(r) + torch.randn((B, k_i), generator=gq, device=device) * NOISE
        batches.append(Q.contiguous())
    return {"E": E, "batches": batches, "N": N, "k_i": k_i, "B": B, "k": k, "n_train": n_train,
            "build_s": {"pinv": t_pinv, "gemm": t_gemm}}


def cpu_topk_throughput(E_host, Q_host, k, repeats, warmup=1):
    """The oracle's CPU restatement of CURApprox.topk_in_row on all host threads; returns (q/s, seconds list)."""
    import torch
    from oracle import cur_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        O.score_topk(Q_host, E_host, k)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return Q_host.shape[0] / med, times


def cpu_sample_rows(N, B):
    """Bounded CPU sample: rows of one batch such that one pass is ~1-3 s on a few cores."""
    rows = int(max(64, min(B, 4.0e8 // N * 1)))          # 100k -> 4000, 1M -> 400, 10M -> 64
    return min(B, rows)


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    import torch
    N, k_i, n_train, B, k = WORKLOADS[args.workload]
    torch.manual_seed(0)
    rows = cpu_sample_rows(N, B)
    # same distribution as the GPU arm; E is formed directly (its build is outside the timed region anyway)
    E = torch.randn((k_i, RANK_LOW)) @ torch.randn((RANK_LOW, N)) / math.sqrt(RANK_LOW) / math.sqrt(k_i)
    Q = torch.randn((rows, k_i))
    from oracle import cur_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(args.warmup):
        O.score_topk(Q, E, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.score_topk(Q, E, k)
    dt = time.perf_counter() - t0
    qps = rows * args.steps / dt
    sample = f"{rows} of the {B} queries of a batch per step, all {N} items, k_i={k_i}, top-{k}; torch.matmul+torch.topk fp32"
    line = {
        "impl": "reference", "metric": "queries/sec (ANNCUR score+top-100)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same config object as the GPU arm prints for these flags (the arm itself runs on the host threads of rank 0)
        "config": workload_config(args, world, args.shard or ("items" if args.workload == "c4" else "queries")),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world, shard):
    N, k_i, n_train, B, k = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: ANNCUR score+top-{k}, N={N} items, k_i={k_i}, batch {B} queries/step",
            "n_items": N, "k_i": k_i, "batch": B, "top_k": k, "precision": args.precision,
            "parallelism": {"queries": f"dp{world}: E replicated, queries sharded, no collective",
                            "items": f"items sharded over {world} ranks, NCCL all-gather + merge",
                            "cpu": "host threads"}[shard],
            "l2": "inputs larger than L2: packed E is streamed every step (>=205 MB at N=100k fp32-grade) and "
                  f"{N_BATCHES} distinct query batches rotate"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f32r", choices=["f32r", "f32x3", "bf16"])
    ap.add_argument("--shard", default=None, choices=["queries", "items"])
    ap.add_argument("--no-extra", action="store_true", help="skip the bf16 / recall side measurements")
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
    args.steps = args.steps if args.steps is not None else 200
    args.warmup = max(3, args.warmup if args.warmup is not None else 10)
    shard = args.shard or ("items" if args.workload == "c4" else "queries")

    import torch
    import torch.distributed as dist
    from anncur_b200 import engine
    from anncur_b200.sharded import ShardedIndex, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    N, k_i, n_train, B, k = WORKLOADS[args.workload]

    # ---- index + queries ---------------------------------------------------------------------------
    if shard == "items" and world > 1:
        lo, hi = shard_bounds(N, world)[rank]
        wl = build_workload(args.workload, device, lo, hi, seed=0, n_batches=N_BATCHES)          # same Q on all ranks
    else:
        lo, hi = 0, N
        wl = build_workload(args.workload, device, 0, N, seed=0, n_batches=N_BATCHES, batch_seed_offset=97 * rank)
    t0 = time.perf_counter()
    packed = engine.PackedItems(wl["E"], args.precision)
    torch.cuda.synchronize()
    t_pack = time.perf_counter() - t0
    batches = wl["batches"]
    index = ShardedIndex(wl["E"], lo, N, precision=args.precision, packed=packed) if (shard == "items" and world > 1) else None

    out_v = torch.empty((B, k), dtype=torch.float32, device=device)
    out_i = torch.empty((B, k), dtype=torch.int64, device=device)

    def step_device(j):
        if index is not None:
            return index.search(batches[j % N_BATCHES], k)
        return engine.score_topk(batches[j % N_BATCHES], packed, k, idx_offset=lo, out=(out_v, out_i))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ----------------------------------------------------------------------
    for j in range(args.warmup):
        step_device(j)
    sync_all()
    engine.profile_enable(True)
    engine.profile_read()
    engine.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        ev0.record()
        for j in range(args.steps):
            step_device(j)
        ev1.record()
        sync_all()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = engine.launch_count()
    fused_ms, fused_n = engine.profile_read()
    engine.profile_enable(False)
    units = B * args.steps * (world if (shard == "queries") else 1)
    qps = units / (ms_total * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI entry ----------------------------------------------
    # Rotating streams, each with its own workspace and pinned result buffers, so that the H2D copy of one batch and the
    # D2H copy of another overlap the kernels of a third (the calls are asynchronous; every step still moves its
    # own query batch in and its own result out inside the timed region).
    N_SLOTS = 3          # measured with the current kernels (tools/e2e_probe.py, 200 steps at C2): 1 stream 0.712 ms/step, 2: 0.525, 3: 0.478, 4: 0.492
    Qh = [b.cpu().pin_memory() for b in batches]
    vh = [torch.empty((B, k), dtype=torch.float32).pin_memory() for _ in range(N_SLOTS)]
    ih = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(N_SLOTS)]
    streams = [torch.cuda.Stream(device=device) for _ in range(N_SLOTS)]
    q_stage = [torch.empty_like(batches[0]) for _ in range(N_SLOTS)]

    def step_e2e(j):
        s = j % N_SLOTS
        with torch.cuda.stream(streams[s]):
            if index is not None:
                # item-sharded: H2D of the (replicated) batch, collective search, D2H of the merged result on every rank
                q_stage[s].copy_(Qh[j % N_BATCHES], non_blocking=True)
                v, i = index.search(q_stage[s], k)
                vh[s].copy_(v, non_blocking=True)
                ih[s].copy_(i, non_blocking=True)
            else:
                engine.search_host(Qh[j % N_BATCHES], packed, k, vh[s], ih[s], idx_offset=lo, ws_key=f"search_host{s}")

    e2e_steps = args.steps
    for j in range(2 * N_SLOTS):
        step_e2e(j)
    sync_all()
    e2e_times = []
    for _ in range(3):                     # host-timed (the copies are part of it): median of three passes of `steps` steps
        t0 = time.perf_counter()
        for j in range(e2e_steps):
            step_e2e(j)
        torch.cuda.synchronize()
        e2e_times.append(max_over_ranks(time.perf_counter() - t0))
        sync_all()
    t_e2e = statistics.median(e2e_times)
    e2e_qps = B * e2e_steps * (world if shard == "queries" else 1) / t_e2e
    h2d = B * k_i * 4
    d2h = B * k * (4 + 8)
    # what the box's host link gives a pinned copy of one step's input / output on its own (CUDA events, best of 5): the e2e
    # figure cannot exceed B / (h2d / h2d_gbs) when the copy engine is the slowest stage of the pipeline
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
    h2d_gbs = copy_gbs(q_stage[0], Qh[0])
    d2h_gbs = copy_gbs(ih[0], out_i)

    # the host-buffer path must give what the device-resident path gives
    chk_v, chk_i = engine.score_topk(batches[(e2e_steps - 1) % N_BATCHES], packed, k, idx_offset=lo) if index is None \
        else index.search(batches[(e2e_steps - 1) % N_BATCHES], k)
    torch.cuda.synchronize()
    assert torch.equal(chk_i.cpu(), ih[(e2e_steps - 1) % N_SLOTS]) and torch.equal(chk_v.cpu(), vh[(e2e_steps - 1) % N_SLOTS]), \
        "host-buffer result differs from the device-resident result"

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json, sustained: the kernel is timed inside a long step loop)"
    except Exception:
        peak_src = "fallback (B200_PROFILING.md)"
    n_local = hi - lo
    fused_ms_avg = fused_ms / max(fused_n, 1)
    flops = 2.0 * B * k_i * n_local
    # bytes of E the MAIN kernel streams per launch: both fp16 planes (f32x3), or one 16-bit plane (bf16; f32r streams
    # only the high plane -- the fp32 copy is touched for the ~k re-scored candidates of a row, by the refine kernel)
    e_bytes = (4 if args.precision == "f32x3" else 2) * k_i * n_local
    alg_bytes = e_bytes + 4 * B * k_i + 12 * B * k
    passes = 3 if args.precision == "f32x3" else 1
    tf = flops / (fused_ms_avg * 1e-3) / 1e12 if fused_n else None
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # dram__bytes_read.sum + dram__bytes_write.sum of the MAIN kernel, one `ncu --set full` capture per (workload, kind)
    # (profiles/r1_ncu_fused_c2_f32x3_v8_details.txt); null where no capture was taken
    # and profiles/r1_ncu_fused_c2_f32r_v16_details.txt
    NCU_TRAFFIC = {("c2", "f32x3"): 238.13e6 + 16.71e6, ("c2", "f32r"): 110.78e6 + 12.60e6}
    roofline = {
        "kernel": "fused_score_topk_kernel (tcgen05 score GEMM + streaming top-k)",
        "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (tf / peak_tf) if tf else None,
        "traffic": NCU_TRAFFIC.get((args.workload, args.precision)) if world == 1 else None, "peak_source": peak_src,
        "launch_ms": fused_ms_avg, "launches_timed": fused_n, "share_of_step": fused_ms / ms_total if ms_total else None,
        "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes,
        "tensor_passes": passes,
        "tensor_pipe_tflops": (tf * passes) if tf else None,
        "tensor_pipe_frac": (tf * passes / peak_tf) if tf else None,
        "note": "achieved/frac count the algorithmic flops 2*B*k_i*N once; kind f32x3 issues 3 f16 tensor passes per product "
                "(h.h + h.l + l.h), so its tensor pipe runs at tensor_pipe_tflops; kind f32r issues one pass (upper bounds of "
                "the scores) and re-scores ~k candidates per row in fp32 in a separate kernel (not part of launch_ms)",
        "hbm_gbs_achieved": alg_bytes / (fused_ms_avg * 1e-3) / 1e9 if fused_n else None,
        "hbm_gbs_peak": peaks.get("hbm_gbs"),
    }

    line = {
        "metric": "queries/sec (ANNCUR score+top-100)", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if shard == "queries" else "strong", "vs_baseline": None,
        "dtype": {"f32x3": "f32 (2xfp16 split operands, 3 tcgen05 passes, fp32 accumulate)", "bf16": "bf16",
                  "f32r": "f32 (one f16 tcgen05 pass of rigorous score upper bounds, fp32 FFMA re-scoring of the candidates)"}[args.precision],
        "data": "synthetic", "config": workload_config(args, world, shard),
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "anncur_search_host (C ABI, pinned host buffers)" if index is None else "ShardedIndex.search + pinned copies",
                "passes_s": e2e_times, "h2d_gbs_alone": h2d_gbs, "d2h_gbs_alone": d2h_gbs,
                "copy_bound_queries_per_s": B / max(h2d / (h2d_gbs * 1e9), d2h / (d2h_gbs * 1e9)) * (world if shard == "queries" else 1)},
        "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline,
        "index_build_s": {"pinv": wl["build_s"]["pinv"], "U@R": wl["build_s"]["gemm"], "pack": t_pack},
    }
    if index is None:
        # rows of the last device-resident batch that needed the fallback pass (0 = the fast path served the whole batch)
        engine.score_topk(batches[0], packed, k, idx_offset=lo, out=(out_v, out_i))
        line["redo_rows_last_batch"] = engine.last_redo_rows(B, packed, k)

    # ---- side measurements (rank 0, single GPU): other precision + recall, CPU baseline -----------------
    if world == 1 and not args.no_extra:
        line["other_precision"] = []
        v_a, i_a = engine.score_topk(batches[0], packed, k)
        for other in [p for p in ("f32r", "f32x3", "bf16") if p != args.precision]:
            packed_o = engine.PackedItems(wl["E"], other)
            for j in range(3):
                engine.score_topk(batches[j % N_BATCHES], packed_o, k, out=(out_v, out_i))
            torch.cuda.synchronize()
            n_o = max(10, args.steps // 4)
            ev0.record()
            for j in range(n_o):
                engine.score_topk(batches[j % N_BATCHES], packed_o, k, out=(out_v, out_i))
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
        E_host = wl["E"].cpu()
        Q_host = batches[0][:rows].cpu()
        cpu_qps, times = cpu_topk_throughput(E_host, Q_host, k, repeats=3)
        # while the CPU result is at hand: the GPU answer for the same rows must be the reference's answer
        from oracle import cur_oracle as O
        ref = O.score_topk(Q_host, E_host, k)
        got_v, got_i = engine.score_topk(batches[0][:rows].contiguous(), packed, k)
        dense_scale = ref.values.abs().max(dim=1, keepdim=True).values
        same = (got_i.cpu() == ref.indices).float().mean().item()
        rel = ((got_v.cpu() - ref.values).abs() / dense_scale).max().item()
        line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{rows} of the {B} queries of one batch x all {N} items, median of 3 after 1 warm-up "
                                          f"({statistics.median(times):.2f} s each); torch.matmul + torch.topk fp32",
                                "gpu_vs_cpu_check": {"index_agreement": same, "max_rel_score_err_sorted_lists": rel}}

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
