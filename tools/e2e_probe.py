#!/usr/bin/env python
"""Experiment: end-to-end (host buffers in, host buffers out) step time of the C2 workload under different
stream arrangements.  Not a benchmark -- decides how anncur_search_host should pipeline its copies.
    python tools/e2e_probe.py [--workload c2] [--steps 100]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from anncur_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--precision", default="f32r")
a = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.build_workload(a.workload, dev, 0, bench.WORKLOADS[a.workload][0], seed=0, n_batches=4)
B, k = wl["B"], wl["k"]
packed = engine.PackedItems(wl["E"], a.precision)
Qh = [b.cpu().pin_memory() for b in wl["batches"]]


def run(name, step, n_slots, finish):
    for j in range(2 * n_slots):
        step(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(a.steps):
        step(j)
    finish()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:50s} {dt / a.steps * 1e3:.4f} ms/step  {B * a.steps / dt / 1e6:.2f} M q/s")


# (a) what bench.py does: n_slots streams, each call = H2D + kernels + D2H on its own stream
for n_slots in (1, 2, 3, 4):
    vh = [torch.empty((B, k), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
    ih = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(n_slots)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_slots)]

    def step(j, n_slots=n_slots, vh=vh, ih=ih, streams=streams):
        s = j % n_slots
        with torch.cuda.stream(streams[s]):
            engine.search_host(Qh[j % 4], packed, k, vh[s], ih[s], ws_key=f"sh{s}")
    run(f"search_host on {n_slots} rotating streams", step, n_slots, lambda: None)

# (b) kernels serialised on one compute stream, copies on dedicated streams, events in between
for n_slots in (2, 3):
    s_in, s_c, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    q_dev = [torch.empty_like(wl["batches"][0]) for _ in range(n_slots)]
    v_dev = [torch.empty((B, k), dtype=torch.float32, device=dev) for _ in range(n_slots)]
    i_dev = [torch.empty((B, k), dtype=torch.int64, device=dev) for _ in range(n_slots)]
    vh = [torch.empty((B, k), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
    ih = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(n_slots)]
    e_in = [torch.cuda.Event() for _ in range(n_slots)]
    e_done = [torch.cuda.Event() for _ in range(n_slots)]
    e_out = [torch.cuda.Event() for _ in range(n_slots)]
    started = [False] * n_slots

    def step(j, n_slots=n_slots):
        s = j % n_slots
        with torch.cuda.stream(s_in):
            if started[s]:
                s_in.wait_event(e_done[s])          # the slot's previous kernels have consumed q_dev[s]
            q_dev[s].copy_(Qh[j % 4], non_blocking=True)
            e_in[s].record(s_in)
        with torch.cuda.stream(s_c):
            s_c.wait_event(e_in[s])
            if started[s]:
                s_c.wait_event(e_out[s])            # the slot's previous results have left v_dev / i_dev
            engine.score_topk(q_dev[s], packed, k, out=(v_dev[s], i_dev[s]))
            e_done[s].record(s_c)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_done[s])
            vh[s].copy_(v_dev[s], non_blocking=True)
            ih[s].copy_(i_dev[s], non_blocking=True)
            e_out[s].record(s_out)
        started[s] = True
    run(f"copy-in / compute / copy-out streams, {n_slots} slots", step, n_slots, lambda: None)
