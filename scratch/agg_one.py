import sys, torch
sys.path.insert(0, '/root/repo')
from gnan_b200 import ops
dev='cuda'; R,N,C,nbins=8192,169343,1,12
g = torch.Generator(device=dev).manual_seed(0)
hop = ops.alloc_hop(R, N, dev); hop[:, :N] = torch.randint(0, nbins-1, (R,N), device=dev, dtype=torch.uint8, generator=g)
T = torch.randn(R, nbins, C, device=dev, requires_grad=True); S = torch.randn(N, C, device=dev, requires_grad=True)
for _ in range(3):
    out = ops.aggregate_rows(hop, T, S, per_row=True); out.sum().backward()
torch.cuda.synchronize()
