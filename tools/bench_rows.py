#!/usr/bin/env python
"""Per-row measurement of the hot-path scope table at the C2 size
(N = 100 000 items, k_i = 500, 2000 anchor queries, B = 4096 test queries): every kernel family timed on the
GPU (CUDA events, 3 warm-ups, median of 5) next to the oracle's CPU restatement of the same reference lines on
the box's host cores (bounded samples, stated per row), with the algorithmic work and the bound that applies.

    python tools/bench_rows.py > profiles/r1_rows_c2.jsonl

Not the headline bench (that is bench.py); this is the evidence that every row of the table runs on the GPU,
and what each one is bounded by."""
import json
import math
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from anncur_b200 import engine
from oracle import cur_oracle as O

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, K_I, N_TRAIN, B, K, K_R = 100_000, 500, 2000, 4096, 100, 500
PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


def gpu_ms(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def cpu_s(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def emit(row, what, gpu_ms_, cpu_s_full, cpu_note, **kw):
    d = {"row": row, "what": what, "gpu_ms": round(gpu_ms_, 4), "cpu_s_scaled_to_full": round(cpu_s_full, 4),
         "gpu_vs_cpu": round(cpu_s_full * 1e3 / gpu_ms_, 1), "cpu_cores": os.cpu_count(), "cpu_sample": cpu_note}
    d.update(kw)
    print(json.dumps(d), flush=True)


torch.set_num_threads(os.cpu_count() or 1)
r = 64
Y = torch.randn(N, r, device=dev)
Xtr = torch.randn(N_TRAIN, r, device=dev)
R = Xtr @ Y.t() / math.sqrt(r) + 0.05 * torch.randn(N_TRAIN, N, device=dev)          # anchor-query rows (k_q x N)
anc = torch.from_numpy(np.sort(np.random.default_rng(0).choice(N, K_I, replace=False))).to(dev)
C = R[:, anc].contiguous()                                                            # k_q x k_i
Xte = torch.randn(B, r, device=dev)
A_test = Xte @ Y.t() / math.sqrt(r) + 0.05 * torch.randn(B, N, device=dev)           # exact scores of the test queries
Q = A_test[:, anc].contiguous()

# A1 pinv
U = engine.pinv(C)
t = gpu_ms(lambda: engine.pinv(C), warm=1, reps=3)
Ch = C.cpu()
emit("A1", f"U = pinv(C[{N_TRAIN}x{K_I}])  (fp64 normal equations: Gram + cooperative Cholesky + solves; Jacobi SVD fallback)", t, cpu_s(lambda: O.pinv_f32(Ch)), "full size, np.linalg.pinv",
     bound="fp64 Gram + latency of the cooperative Cholesky panels (the Jacobi route, taken for rank-deficient / ill-conditioned inputs, is a tournament of column-pair rotations: 39 ms)")
# A2 E = U R
E = engine.gemm(U, R)
t = gpu_ms(lambda: engine.gemm(U, R))
Uh, Rh = U.cpu(), R.cpu()
fl = 2.0 * K_I * N_TRAIN * N
emit("A2", f"E = U[{K_I}x{N_TRAIN}] @ R[{N_TRAIN}x{N}]  (fp32 FFMA)", t, cpu_s(lambda: Uh @ Rh), "full size, torch.matmul",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="fp32 FFMA pipe (no fp32 tensor path that keeps fp32 products)")
t = gpu_ms(lambda: engine.gemm_tc(U, R))
emit("A2", f"E = U @ R on the tcgen05 pipeline (fp32-grade 3-pass; packs R first)", t, cpu_s(lambda: Uh @ Rh), "full size, torch.matmul",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="tensor pipe at 1/3 rate + one packing pass over R (0.8 GB read, 0.8 GB written)")
# A3 dense scores
t = gpu_ms(lambda: engine.gemm(Q, E))
Qh, Eh = Q.cpu(), E.cpu()
fl = 2.0 * B * K_I * N
s_cpu = cpu_s(lambda: Qh[:1024] @ Eh) * (B / 1024)
emit("A3", f"scores = Q[{B}x{K_I}] @ E[{K_I}x{N}] dense (get_complete_row)", t, s_cpu, "1024 of 4096 rows, torch.matmul",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="fp32 FFMA pipe; writes the 1.6 GB score matrix")
packed_r = engine.PackedItems(E, "f32r")
dense_out = torch.empty((B, N), dtype=torch.float32, device=dev)
t = gpu_ms(lambda: engine.score_dense(Q, packed_r, out=dense_out))
emit("A3", f"scores = Q @ E dense on the tcgen05 pipeline (fp32-grade 3-pass, F32R index)", t, s_cpu, "1024 of 4096 rows, torch.matmul",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="tensor pipe at 1/3 rate; writes the 1.6 GB score matrix")
del dense_out
# A4 fused
for kind in ("f32r", "f32x3", "bf16"):
    packed = engine.PackedItems(E, kind)
    ov = torch.empty((B, K), dtype=torch.float32, device=dev)
    oi = torch.empty((B, K), dtype=torch.int64, device=dev)
    t = gpu_ms(lambda: engine.score_topk(Q, packed, K, out=(ov, oi)))
    s_cpu = cpu_s(lambda: O.score_topk(Qh[:1024], Eh, K)) * (B / 1024)
    emit("A4", f"topk_in_row: top-{K} of Q @ E fused, kind {kind}", t, s_cpu, "1024 of 4096 rows, torch.matmul + torch.topk",
         flops=fl, tflops=round(fl / t / 1e9, 1), bound="tensor pipe (see bench.py roofline)")
# A5 exact top-k over the dense exact scores + approx top-k_r + rerank/overlap
t1 = gpu_ms(lambda: engine.topk_rows(A_test, K))
byts = 4.0 * B * N
ex_v, ex_i = engine.topk_rows(A_test, K)
ap_v, ap_i = engine.score_topk(Q, engine.PackedItems(E, "f32r"), K_R)
t2 = gpu_ms(lambda: engine.rerank_overlap(A_test, ap_i, ex_i, [1, 10, 50, 100]))
Ah = A_test[:64].cpu()
approx_h = (Qh[:64] @ Eh)
s_cpu = cpu_s(lambda: O.retrieve_and_rerank(Ah, approx_h, K, K_R), reps=1) * (B / 64)
emit("A5", f"exact top-{K} of {B}x{N} scores (torch.topk per row)", t1, s_cpu, "64 of 4096 rows, the reference's per-query Python loop (3 topk + N-long temp)",
     bytes=byts, gbs=round(byts / t1 / 1e6, 1), hbm_frac=round(byts / t1 / 1e6 / HBM, 3), bound="HBM (one read of the score matrix)")
emit("A5+A6", f"rerank top-{K_R} by exact score + overlap for k in 1,10,50,100", t2, s_cpu, "same loop (rerank + compute_overlap are inside it)",
     bound="latency (gather of k_r scattered scores per row, smem sort)")
# A7 reconstruction error
t = gpu_ms(lambda: engine.recon_error_rows(Q, E, A_test))
s_cpu = cpu_s(lambda: (torch.norm(Qh[:1024] @ Eh - A_test[:1024].cpu()), torch.norm(A_test[:1024].cpu()))) * (B / 1024)
emit("A7", f"|Q E - A|_F and |A|_F over {B}x{N} (never materialises Q E)", t, s_cpu, "1024 of 4096 rows, torch",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="fp32 FFMA pipe + one HBM read of A")
t = gpu_ms(lambda: engine.recon_error_packed(Q, packed_r, A_test))
emit("A7", f"|Q E - A|_F and |A|_F on the tcgen05 pipeline (fp32-grade 3-pass)", t, s_cpu, "1024 of 4096 rows, torch",
     flops=fl, tflops=round(fl / t / 1e9, 1), bound="tensor pipe at 1/3 rate + one HBM read of A (1.6 GB)")
# A8 adaptive round (m = 250 anchors per query, k_q = 500 anchor queries, next 125)
Bq, m, k_q = 256, 250, 500
R_anc = R[:k_q].contiguous()
anchors = torch.stack([torch.randperm(N, device=dev)[:m] for _ in range(Bq)])
c = torch.gather(A_test[:Bq], 1, anchors)
t = gpu_ms(lambda: engine.adaptive_round(R_anc, anchors, c, 125), warm=1, reps=3)
Rn, An = R_anc.cpu().numpy(), A_test[:4].cpu().numpy()
anc_n = anchors[:4].cpu().numpy()


def cpu_adaptive():
    for q in range(4):
        M = Rn[:, anc_n[q]].astype(np.float64)
        e = An[q, anc_n[q]] @ np.linalg.pinv(M, rcond=1e-15)
        s = (e @ Rn.astype(np.float64)).astype(np.float32)
        s[anc_n[q]] = -np.inf
        np.argpartition(-s, 125)[:125]


emit("A8", f"adaptive round: {Bq} queries x (pinv of {k_q}x{m} + re-score over {N} + top-125)  [not in the reference]", t,
     cpu_s(cpu_adaptive, reps=1) * (Bq / 4), "4 of 256 queries, numpy pinv per query (oracle restatement of SURVEY 8a-A8)",
     bound="fp64 Gram on the fp64 tensor cores + blocked fp64 Cholesky per query")

# A8 whole procedure at BASELINE configs[2]: 4 rounds x 125 anchors, incremental solver + fused re-score
from anncur_b200 import adaptive_anncur
from anncur_b200.adaptive import AdaptiveIndex
index = AdaptiveIndex(R_anc)
first = torch.randperm(N, device=dev)[:125].sort().values
t = gpu_ms(lambda: adaptive_anncur(R_anc, A_test, first, 4, 125, 100, index=index), warm=2, reps=3)
An2 = A_test[:2].cpu().numpy()
t_cpu = cpu_s(lambda: O.adaptive_anncur(Rn, An2, first.cpu().numpy(), 4, 125, 100), reps=1) * (B / 2)
emit("A8", f"adaptive ANNCUR, whole procedure: {B} queries x 4 rounds x 125 anchors over {N} items (incremental fp64 solver + fused re-score)  [not in the reference]",
     t, t_cpu, "2 of 4096 queries, oracle.adaptive_anncur (numpy fp64 pinv per query and round)",
     queries_per_s=round(B / t * 1e3), bound="fp64 DMMA (Gram rows 28 %, block GEMMs 25 %), L2 round trips of the substitution kernel 17 %, re-score 18 %")
