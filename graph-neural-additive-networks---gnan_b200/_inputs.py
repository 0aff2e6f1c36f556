"""Resolving the `inputs` object the reference passes to forward (a PyG `Data`: GNAN.py:56, trainer.py:46-52) into
device tensors the kernels take. Accepts either the compact form (`inputs.hop_data`, a preprocess.HopData) or the
reference's fp32 `[N,N]` pair `node_distances` / `normalization_matrix`, which is converted once and cached on the
object."""
import torch

from .preprocess import HopData, from_reference_format


AUTO_DEDUP_MIN_EVALUATIONS = 1 << 16     # below this many (row, feature) pairs the dense kernels are a single small launch


def _version(t):
    """in-place modification counter of a tensor (inference tensors have none: they cannot be modified in place either)"""
    try:
        return t._version
    except RuntimeError:
        return 0


def capturing():
    """True while the current CUDA stream is being captured into a graph (False where there is no CUDA device at all)"""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def compressed_of(holder, x, device, enabled=True):
    """CompressedFeatures to use for `holder`, or None for the dense kernels.

    `holder.x_compressed` (built by the caller with sparse.compress_features) is always honoured. Otherwise the compressed
    form is built from `x` on first use and cached on the object, but only where that pays: objects that persist across steps
    (the reference's Data objects, PackedDataset.as_batch()) and carry at least AUTO_DEDUP_MIN_EVALUATIONS (row, feature)
    pairs. A PackedBatch made per step (apsp_batched, PackedDataset.batch) is never analysed implicitly: the column modes and
    sorts (plus a few host syncs) would cost more than they save."""
    if not enabled:
        return None
    cx = getattr(holder, "x_compressed", None)
    if cx is None and x is not None and capturing():
        return None        # an implicit (cached) compressed form would be baked into the CUDA graph and go stale when x is
                           # refreshed in place between replays; pass x_compressed explicitly to share evaluations there
    if cx is None and x is not None:
        from .preprocess import PackedBatch
        if isinstance(holder, PackedBatch) and not getattr(holder, "_gnan_b200_persistent", False):
            return None
        if x.shape[0] * x.shape[1] < AUTO_DEDUP_MIN_EVALUATIONS:
            return None
        key = (x.data_ptr(), _version(x), tuple(x.shape))
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
        nm = getattr(inputs, "normalization_matrix", None)
        # identity AND version of both matrices: an in-place refresh (copy_) must not hit the converted copy of the old values
        key = (nd.data_ptr(), _version(nd), tuple(nd.shape)) + (() if nm is None else (nm.data_ptr(), _version(nm)))
        if cache is not None and cache[0] == key:
            hd = cache[1]
        else:
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
