"""Times the plain GEMMs around gnan_mlp_bwd_ext at the ogbn-arxiv shape (dev tool; prints ms per variant)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gnan_b200 import ops

R, G, H, C = 169343, 129, 64, 40
dev = "cuda"
dS = torch.randn(R, C, device=dev)
wo = torch.randn(G, C, H, device=dev)
a1 = torch.randn(R, G * H, device=dev)
B = wo.permute(1, 0, 2).reshape(C, G * H)


def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


torch.backends.cuda.matmul.allow_tf32 = False
print("dh fp32 simt      ", t(lambda: torch.matmul(dS, B)))
print("dh 3xtf32 concat  ", t(lambda: ops._matmul_3xtf32(dS, B)))
print("dWo fp32 simt     ", t(lambda: torch.matmul(dS.t(), a1)))
dSt = dS.t().contiguous()
print("dWo fp32 (dS^T contiguous)", t(lambda: torch.matmul(dSt, a1)))
torch.backends.cuda.matmul.allow_tf32 = True
print("dWo tf32 single   ", t(lambda: torch.matmul(dS.t(), a1)))
ah, al = ops._split_tf32(dS)
print("split a1          ", t(lambda: ops._split_tf32(a1)))
h, l = ops._split_tf32(a1)
print("dWo 3 tf32 gemms  ", t(lambda: torch.matmul(ah.t(), h) + torch.matmul(al.t(), h) + torch.matmul(ah.t(), l)))
# chunked along R with fp32 accumulate through addmm
out = torch.zeros(C, G * H, device=dev)
print("dWo baddbmm split-K 64", t(lambda: torch.bmm(dS[: R // 64 * 64].view(64, -1, C).transpose(1, 2), a1[: R // 64 * 64].view(64, -1, G * H)).sum(0)))
torch.backends.cuda.matmul.allow_tf32 = False
print("dWo fp32 bmm split-K 64", t(lambda: torch.bmm(dS[: R // 64 * 64].view(64, -1, C).transpose(1, 2), a1[: R // 64 * 64].view(64, -1, G * H)).sum(0)))
print("copy 5.6GB        ", t(lambda: a1.clone()))
