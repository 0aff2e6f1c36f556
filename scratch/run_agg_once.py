import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from gnan_b200 import ops
R, N, C, nb = 16384, 169343, 1, 10
hop = torch.randint(0, nb - 1, (R, ops.hop_ld(N)), dtype=torch.uint8, device="cuda")
T = torch.randn(nb, C, device="cuda", requires_grad=True)
S = torch.randn(N, C, device="cuda", requires_grad=True)
for _ in range(2):
    out = ops.aggregate_rows(hop, T, S)
    out.sum().backward()
torch.cuda.synchronize()
