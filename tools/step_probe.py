#!/usr/bin/env python
"""Small driver for profiling: random low-rank E and Q of a given shape, a few score+top-k steps.
    python tools/step_probe.py --n 100000 --ki 500 --b 4096 --k 100 --precision f32x3 --steps 5
Prints per-step time (CUDA events).  Used under ncu for launch lists / kernel captures; not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anncur_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--ki", type=int, default=500)
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--precision", default="f32r")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (engine.GraphedSearch)")
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda", 0)
r = 64
Y = torch.randn(a.n, r, device=dev)
W = torch.randn(a.ki, r, device=dev)
E = (W @ Y.t()) / r ** 0.5 + 0.05 * torch.randn(a.ki, a.n, device=dev)
Q = [torch.randn(a.b, r, device=dev) @ W.t() / r ** 0.5 + 0.05 * torch.randn(a.b, a.ki, device=dev) for _ in range(2)]
packed = engine.PackedItems(E, a.precision)
ov = [torch.empty((a.b, a.k), dtype=torch.float32, device=dev) for _ in range(a.streams)]
oi = [torch.empty((a.b, a.k), dtype=torch.int64, device=dev) for _ in range(a.streams)]
streams = [torch.cuda.Stream() for _ in range(a.streams)] if a.streams > 1 else [torch.cuda.current_stream()]
graphed = engine.GraphedSearch(packed, a.b, a.k) if a.graph else None
def step(j):
    if graphed is not None:
        return graphed(Q[j % 2])
    s = j % a.streams
    with torch.cuda.stream(streams[s]):
        engine.score_topk(Q[j % 2], packed, a.k, out=(ov[s], oi[s]))
for j in range(a.warmup * a.streams):
    step(j)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
engine.profile_enable(True)
e0.record()
for s_ in streams[1:] if a.streams > 1 else []:
    s_.wait_stream(torch.cuda.current_stream())
for j in range(a.steps):
    step(j)
for s_ in streams if a.streams > 1 else []:
    torch.cuda.current_stream().wait_stream(s_)
e1.record()
torch.cuda.synchronize()
ms, n = engine.profile_read()
print(f"step {e0.elapsed_time(e1) / a.steps:.4f} ms  main kernel {ms / max(n, 1):.4f} ms  ({a.b / (e0.elapsed_time(e1) / a.steps) * 1e3:.0f} q/s)")
