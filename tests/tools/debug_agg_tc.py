"""Diagnosis of the tensor-core aggregation (csrc/agg_tc.cu) on structured inputs: which column / row / bin does a single
marked pair land in? Prints mismatches compactly. Run on a GPU box: python tests/tools/debug_agg_tc.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gnan_b200 import ops  # noqa: E402

DEV = torch.device("cuda")


def run(R, N, C, nbins, marks, base=3, seed=0, algo="tc"):
    """hop = base everywhere, marks = [(i, j, h)]; S[j, c] = (j + 1) * (c + 1); global table T = 1 -> returns Bsum [R, nbins, C]"""
    hop = ops.alloc_hop(R, N, DEV)
    h = np.full((R, N), base, dtype=np.uint8)
    for i, j, v in marks:
        h[i, j] = v
    hop[:, :N] = torch.tensor(h, device=DEV)
    S = (torch.arange(1, N + 1, dtype=torch.float32)[:, None] * torch.arange(1, C + 1, dtype=torch.float32)[None, :]).to(DEV)
    T = torch.ones(nbins, C, device=DEV)
    out, bsum = ops.agg_rows_fwd(hop, T, None, S, False, True, {"tc": 2, "cuda": 1}[algo])
    torch.cuda.synchronize()
    return out.cpu().numpy(), bsum.cpu().numpy(), h, S.cpu().numpy()


def expect(h, S, nbins):
    R, N = h.shape
    b = np.where(h == 255, nbins - 1, h)
    out = np.zeros((R, nbins, S.shape[1]))
    for d in range(nbins):
        out[:, d, :] = (b == d).astype(np.float64) @ S.astype(np.float64)
    return out


def main():
    for nbins in (6, 12, 20):
        R, N, C = 40, 300, 2
        print(f"=== nbins={nbins} R={R} N={N} C={C}")
        # 1) total: everything in bin `base`
        out, bs, h, S = run(R, N, C, nbins, [])
        want = expect(h, S, nbins)
        print(" base-only max abs err", np.abs(bs - want).max(), " bin3 row0 got", bs[0, 3], "want", want[0, 3])
        if np.abs(bs - want).max() > 1e-3:
            nz = np.argwhere(np.abs(bs) > 0)
            print("  nonzero (row,bin,c) sample:", nz[:12].tolist(), "values", [float(bs[tuple(t)]) for t in nz[:6]])
        # 2) one marked pair at a time: which (row, bin, column) receives it
        for (i, j, v) in [(0, 0, 2), (0, 1, 2), (0, 5, 2), (0, 17, 2), (0, 130, 2), (1, 2, 2), (9, 3, 2), (33, 4, 2), (0, 7, 0),
                          (0, 9, nbins - 2), (2, 11, 255), (39, 299, 1)]:
            out, bs, h, S = run(R, N, C, nbins, [(i, j, v)])
            want = expect(h, S, nbins)
            err = np.abs(bs - want).max()
            d = nbins - 1 if v == 255 else v
            msg = f" mark (row {i}, col {j}, hop {v}): max err {err:.3g}"
            if err > 1e-3:
                diff = bs - want
                nz = np.argwhere(np.abs(diff) > 1e-3)
                msg += "  wrong at " + str([(tuple(t), float(diff[tuple(t)])) for t in nz[:6]]) + f"  (S[j]={S[j, 0]})"
            print(msg)
    # 3) random case against float64
    rng = np.random.default_rng(0)
    for (R, N, C, nbins) in [(70, 1000, 3, 10), (70, 1000, 7, 16), (33, 515, 40, 9), (50, 700, 1, 27)]:
        h = rng.integers(0, nbins, size=(R, N))
        hb = h.copy(); hb[hb == nbins - 1] = 255
        hop = ops.alloc_hop(R, N, DEV)
        hop[:, :N] = torch.tensor(hb.astype(np.uint8), device=DEV)
        S = torch.tensor(rng.normal(size=(N, C))).float()
        T = torch.tensor(rng.normal(size=(nbins, C))).float()
        out, bsum = ops.agg_rows_fwd(hop, T.to(DEV), None, S.to(DEV), False, True, 2)
        out2, bsum2 = ops.agg_rows_fwd(hop, T.to(DEV), None, S.to(DEV), False, True, 1)
        want = expect(hb, S.numpy(), nbins)
        e1 = np.abs(bsum.cpu().numpy() - want).max() / np.abs(want).max()
        e2 = np.abs(bsum2.cpu().numpy() - want).max() / np.abs(want).max()
        wo = (want * T.numpy()[None].astype(np.float64)).sum(1)
        print(f"random R={R} N={N} C={C} nbins={nbins}: Bsum rel err tc {e1:.3g} cuda {e2:.3g}; out rel err tc "
              f"{np.abs(out.cpu().numpy() - wo).max() / np.abs(wo).max():.3g} cuda {np.abs(out2.cpu().numpy() - wo).max() / np.abs(wo).max():.3g}")


if __name__ == "__main__":
    main()
