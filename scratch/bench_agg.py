import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from gnan_b200 import ops
dev = 'cuda'
def run(R, N, C, nbins, per_row=True, iters=5):
    g = torch.Generator(device=dev).manual_seed(0)
    hop = ops.alloc_hop(R, N, dev)
    h = torch.randint(0, nbins - 1, (R, N), device=dev, dtype=torch.uint8, generator=g)
    hop[:, :N] = h
    T = torch.randn((R, nbins, C) if per_row else (nbins, C), device=dev, requires_grad=True)
    S = torch.randn(N, C, device=dev, requires_grad=True)
    gO = torch.randn(R, C, device=dev)
    for _ in range(2):
        out = ops.aggregate_rows(hop, T, S, per_row=per_row); (out * gO).sum().backward()
    tf = tb = 0.0
    for _ in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); out = ops.aggregate_rows(hop, T, S, per_row=per_row); e[1].record()
        (out * gO).sum().backward(); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    tf /= iters; tb /= iters
    gb = R * N / 1e9
    print(f"R={R} N={N} C={C} nbins={nbins}: fwd {tf:.3f} ms ({gb/tf*1e3:.0f} GB/s)  bwd {tb:.3f} ms ({gb/tb*1e3:.0f} GB/s)")
run(16384, 169343, 1, 12)
run(19717, 19717, 3, 12)
run(4096, 169343, 4, 12)
run(2708, 2708, 7, 12)
