#!/usr/bin/env python
"""Experiment: BASELINE configs[2] as two half batches on two streams (latency-bound solver kernels of one half under the
DMMA-bound Gram launches of the other) against one batch on one stream.
    python tools/adaptive_2stream_probe.py"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anncur_b200 import adaptive_anncur
from anncur_b200.adaptive import AdaptiveIndex

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, kq, B, r = 100000, 500, 4096, 64
Y = torch.randn(N, r, device=dev)
R = torch.randn(kq, r, device=dev) @ Y.t() / r ** 0.5 + 0.05 * torch.randn(kq, N, device=dev)
X = torch.randn(B, r, device=dev) @ Y.t() / r ** 0.5 + 0.05 * torch.randn(B, N, device=dev)
first = torch.randperm(N, device=dev)[:125].sort().values
index = AdaptiveIndex(R)


def run(parts, threads):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = X.chunk(parts)
    out = [None] * parts

    def work(i):
        with torch.cuda.stream(streams[i]):
            out[i] = adaptive_anncur(R, chunks[i], first, 4, 125, 100, index=index)

    def once():
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        if threads:
            ts = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        else:
            for i in range(parts):
                work(i)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)

    once(); once()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        once()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[1] * 1e3, out


base, o1 = run(1, False)
print(f"1 stream, one batch of {B}: {base:.2f} ms ({B / base * 1e3:.0f} q/s)")
for parts, threads in ((2, False), (2, True), (4, True)):
    ms, o = run(parts, threads)
    same = torch.equal(torch.cat([x[0] for x in o]), o1[0][0])
    print(f"{parts} streams ({'one host thread each' if threads else 'issued from one thread'}): {ms:.2f} ms ({B / ms * 1e3:.0f} q/s), anchors equal: {same}")
