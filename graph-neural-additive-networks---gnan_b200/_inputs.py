"""Resolving the `inputs` object the reference passes to forward (a PyG `Data`: GNAN.py:56, trainer.py:46-52) into
device tensors the kernels take. Accepts either the compact form (`inputs.hop_data`, a preprocess.HopData) or the
reference's fp32 `[N,N]` pair `node_distances` / `normalization_matrix`, which is converted once and cached on the
object."""
import torch

from .preprocess import HopData, from_reference_format


def resolve(inputs, device):
    x = inputs.x
    if not torch.is_tensor(x):
        raise TypeError("inputs.x must be a tensor")
    x = x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    hd = getattr(inputs, "hop_data", None)
    if hd is None:
        cache = getattr(inputs, "_gnan_b200_hop_cache", None)
        nd = getattr(inputs, "node_distances", None)
        if nd is None:
            raise AttributeError("inputs needs .hop_data (gnan_b200.preprocess) or .node_distances/.normalization_matrix")
        key = (nd.data_ptr(), tuple(nd.shape))
        if cache is not None and cache[0] == key:
            hd = cache[1]
        else:
            nm = getattr(inputs, "normalization_matrix", None)
            hd = from_reference_format(nd.to(device, non_blocking=True).float(),
                                       None if nm is None else nm.to(device, non_blocking=True))
            try:
                inputs._gnan_b200_hop_cache = (key, hd)
            except Exception:
                pass
    elif not isinstance(hd, HopData):
        raise TypeError("inputs.hop_data must be a gnan_b200.preprocess.HopData")
    if hd.hop.device != x.device:
        hd = hd.to(x.device)
    return x, hd
