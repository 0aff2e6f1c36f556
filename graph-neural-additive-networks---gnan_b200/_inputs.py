"""Resolving the `inputs` object the reference passes to forward (a PyG `Data`: GNAN.py:56, trainer.py:46-52) into
device tensors the kernels take. Accepts either the compact form (`inputs.hop_data`, a preprocess.HopData) or the
reference's fp32 `[N,N]` pair `node_distances` / `normalization_matrix`, which is converted once and cached on the
object."""
import torch

from .preprocess import HopData, from_reference_format


def compressed_of(holder, x, device, enabled=True):
    """CompressedFeatures of `holder` (its `.x_compressed`, else built from `x` once and cached on the object), or None when
    the dense kernels should be used (disabled, or too many distinct values: sparse.compress_features returns None)."""
    if not enabled:
        return None
    cx = getattr(holder, "x_compressed", None)
    if cx is None and getattr(holder, "_gnan_b200_no_dedup", False):
        return None                      # a throw-away batch object: building (sorts, syncs) would cost more than it saves
    if cx is None and x is not None:
        key = (x.data_ptr(), x._version, tuple(x.shape))
        cache = getattr(holder, "_gnan_b200_cx_cache", None)
        if cache is not None and cache[0] == key:
            cx = cache[1]
        else:
            from .sparse import compress_features
            cx = compress_features(x.to(device, non_blocking=True))
            try:
                holder._gnan_b200_cx_cache = (key, cx)
            except Exception:
                pass
    if cx is not None and cx.device != torch.device(device):
        cx = cx.to(device)
    return cx


def resolve(inputs, device, need_x=True):
    x = inputs.x
    if x is None and getattr(inputs, "x_compressed", None) is not None:
        need_x = False
    elif not torch.is_tensor(x):
        raise TypeError("inputs.x must be a tensor")
    if need_x:
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
    if hd.hop.device != torch.device(device):
        hd = hd.to(device)
    return (x if need_x else None), hd
