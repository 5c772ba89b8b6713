import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anncur_b200 import engine
dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, k_q, m, B = 100000, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 500, 768
R = torch.randn(k_q, 64, device=dev) @ torch.randn(64, N, device=dev) / 8 + 0.05 * torch.randn(k_q, N, device=dev)
A = torch.randn(B, 64, device=dev) @ torch.randn(64, N, device=dev) / 8
anchors = torch.randint(0, N, (B, m), device=dev)
anchors = torch.stack([torch.arange(m, device=dev) * (N // m) + (i % (N // m)) for i in range(B)])
c = torch.gather(A, 1, anchors)
engine.adaptive_round(R, anchors, c, 125)
torch.cuda.synchronize()
