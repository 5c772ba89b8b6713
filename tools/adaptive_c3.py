#!/usr/bin/env python
"""BASELINE config 3 timed end to end: adaptive multi-round ANNCUR, N = 100 000 items, k_q = 500 anchor queries, 4 rounds of
125 anchors (k_i = 500 in total), B = 4096 queries, top-100 by exact score.  Synthetic low-rank-plus-noise scores.
    python tools/adaptive_c3.py [--b 4096]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anncur_b200 import adaptive_anncur

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--kq", type=int, default=500)
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--per_round", type=int, default=125)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--rescore", default="fused", choices=["fused", "ffma"])
ap.add_argument("--solver", default="incremental", choices=["incremental", "full"])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
r = 64
Y = torch.randn(a.n, r, device=dev)
R = torch.randn(a.kq, r, device=dev) @ Y.t() / r ** 0.5 + 0.05 * torch.randn(a.kq, a.n, device=dev)
X = torch.randn(a.b, r, device=dev) @ Y.t() / r ** 0.5 + 0.05 * torch.randn(a.b, a.n, device=dev)
first = torch.randperm(a.n, device=dev)[:a.per_round].sort().values
from anncur_b200.adaptive import AdaptiveIndex
index = AdaptiveIndex(R) if a.rescore == "fused" else None
t0 = time.perf_counter()
adaptive_anncur(R, X[:256], first, a.rounds, a.per_round, a.k, rescore=a.rescore, index=index, solver=a.solver)
torch.cuda.synchronize()
t_first = time.perf_counter() - t0                     # includes anncur_adaptive_prepare (once per index + first anchors)
adaptive_anncur(R, X, first, a.rounds, a.per_round, a.k, rescore=a.rescore, index=index, solver=a.solver)
torch.cuda.synchronize()
dts = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    anc, idx, val = adaptive_anncur(R, X, first, a.rounds, a.per_round, a.k, rescore=a.rescore, index=index, solver=a.solver)
    torch.cuda.synchronize()
    dts.append(time.perf_counter() - t0)
dt = sorted(dts)[len(dts) // 2]
exact = torch.topk(X, a.k, dim=1).indices
recall = (idx.unsqueeze(2) == exact.unsqueeze(1)).any(2).float().mean().item()
print(f"adaptive ANNCUR ({a.rescore} re-score, {a.solver} solver; first call on 256 queries incl. one-off set-up {t_first * 1e3:.1f} ms): {a.b} queries x {a.rounds} rounds x {a.per_round} anchors over {a.n} items: {dt * 1e3:.1f} ms "
      f"({a.b / dt:.0f} q/s), recall@{a.k} vs exact = {recall:.3f}")
