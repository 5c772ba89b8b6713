import sys, os
sys.path.insert(0, "/root/repo")
import torch, math
from anncur_b200 import engine
dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, k_q, m, B = 100000, 500, 250, 1024
R = torch.randn(k_q, 64, device=dev) @ torch.randn(64, N, device=dev) / 8 + 0.05 * torch.randn(k_q, N, device=dev)
A = torch.randn(B, 64, device=dev) @ torch.randn(64, N, device=dev) / 8
anchors = torch.stack([torch.randperm(N, device=dev)[:m] for _ in range(B)])
c = torch.gather(A, 1, anchors)
for _ in range(2):
    engine.adaptive_round(R, anchors, c, 125)
torch.cuda.synchronize()
import time
for mm in (125, 250, 375, 500):
    anchors = torch.stack([torch.randperm(N, device=dev)[:mm] for _ in range(B)])
    c = torch.gather(A, 1, anchors)
    engine.adaptive_round(R, anchors, c, 125)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nxt, val = engine.adaptive_round(R, anchors, c, 125)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # check against fp64 pinv on 3 queries
    import numpy as np
    Rn = R.double().cpu().numpy()
    ok = True
    for q in range(3):
        an = anchors[q].cpu().numpy()
        e = c[q].double().cpu().numpy() @ np.linalg.pinv(Rn[:, an])
        sc = e @ Rn
        sc[an] = -np.inf
        want = np.argsort(-sc)[:125]
        got = nxt[q].cpu().numpy()
        tau = 1e-4 * np.abs(sc[np.isfinite(sc)]).max()
        ok = ok and all(sc[j] >= sc[want[-1]] - tau for j in got)
    print(f"m = {mm}: {dt * 1e3:.2f} ms per round of {B} queries, parity ok = {ok}")
