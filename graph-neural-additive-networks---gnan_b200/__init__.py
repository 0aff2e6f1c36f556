"""gnan_b200 — B200 (sm_100a) implementation of the GNAN hot path behind the reference's nn.Module API.

    from gnan_b200.GNAN import GNAN, TensorGNAN          # reference: GNAN.py
    from gnan_b200.models import GNAN, TensorGNAN        # reference: models.py (what main.py imports)
    from gnan_b200.batched import TensorGNAN             # reference: batched_pyg_main.py
    from gnan_b200.preprocess import pre_process, apsp   # reference: pre_process_datasets.py

The compute path is libgnan_b200.so (csrc/, C ABI in include/gnan_b200.h), loaded lazily on first use; there is no CPU
or eager fallback. Build it with `python __graft_entry__.py` or `python <package>/build.py`.
"""
__version__ = "0.1.0"
