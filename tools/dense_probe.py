#!/usr/bin/env python
"""Timing of the dense products (A2 item-embedding build, A3 get_complete_row, A7 reconstruction error): FFMA kernels against
the tcgen05 pipeline (anncur_score_dense / anncur_recon_error_packed).  C2 sizes by default.  Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anncur_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--ki", type=int, default=500)
ap.add_argument("--kq", type=int, default=2000)
ap.add_argument("--b", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
U = torch.randn(a.ki, a.kq, device=dev) / a.kq ** 0.5
R = torch.randn(a.kq, a.n, device=dev)
Q = torch.randn(a.b, a.ki, device=dev)


def timed(name, fn, flops, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:58s} {ms:9.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
    return out


f_build = 2.0 * a.ki * a.kq * a.n
E1 = timed("A2 E = U @ R            FFMA (anncur_gemm_f32)", lambda: engine.gemm(U, R), f_build)
E2 = timed("A2 E = U @ R            tcgen05 3-pass incl. packing R", lambda: engine.gemm_tc(U, R), f_build)
pr = engine.PackedItems(R, "f32x3")
timed("A2 E = U @ R            tcgen05 3-pass, R already packed", lambda: engine.score_dense(U, pr), f_build)
print("   max |E_tc - E_ffma| / max|E| =", ((E1 - E2).abs().max() / E1.abs().max()).item())
del pr
f_row = 2.0 * a.b * a.ki * a.n
packed = engine.PackedItems(E1, "f32r")
out = torch.empty((a.b, a.n), device=dev)
timed("A3 scores = Q @ E dense FFMA", lambda: engine.gemm(Q, E1), f_row, reps=3)
timed("A3 scores = Q @ E dense tcgen05 3-pass (F32R index)", lambda: engine.score_dense(Q, packed, out=out), f_row, reps=3)
A = out + 0.01 * torch.randn_like(out)
timed("A7 recon error          FFMA", lambda: engine.recon_error_rows(Q, E1, A), f_row, reps=3)
timed("A7 recon error          tcgen05 3-pass", lambda: engine.recon_error_packed(Q, packed, A), f_row, reps=3)
