"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes wrapper over oracle/apsp_oracle.c (see its header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_apsp.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def apsp(edge_index, num_nodes):
    """int32 [n,n] hop matrix, -1 unreachable; edge_index int [2,E] followed src->dst."""
    ei = np.ascontiguousarray(np.asarray(edge_index), dtype=np.int64)
    src, dst = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
    hop = np.empty((num_nodes, num_nodes), dtype=np.int32)
    rc = _lib().gnan_oracle_apsp(ctypes.c_int32(num_nodes), ctypes.c_int64(src.shape[0]),
                                 src.ctypes.data_as(ctypes.c_void_p), dst.ctypes.data_as(ctypes.c_void_p),
                                 hop.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RuntimeError(f"gnan_oracle_apsp rc={rc}")
    return hop


def level_counts(hop, nbins=None):
    hop = np.ascontiguousarray(hop, dtype=np.int32)
    if nbins is None:
        nbins = int(hop.max()) + 2
    cnt = np.empty((hop.shape[0], nbins), dtype=np.int32)
    rc = _lib().gnan_oracle_level_counts(ctypes.c_int32(hop.shape[0]), ctypes.c_int32(hop.shape[1]),
                                         hop.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(nbins),
                                         cnt.ctypes.data_as(ctypes.c_void_p))
    if rc < 0:
        raise RuntimeError(f"gnan_oracle_level_counts rc={rc}")
    return cnt


def reference_format(hop, cnt):
    hop = np.ascontiguousarray(hop, dtype=np.int32)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    nd = np.empty(hop.shape, dtype=np.float32)
    nm = np.empty(hop.shape, dtype=np.float32)
    _lib().gnan_oracle_reference_format(ctypes.c_int32(hop.shape[0]), ctypes.c_int32(hop.shape[1]),
                                        hop.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(cnt.shape[1]),
                                        cnt.ctypes.data_as(ctypes.c_void_p),
                                        nd.ctypes.data_as(ctypes.c_void_p), nm.ctypes.data_as(ctypes.c_void_p))
    return nd, nm


def apsp_rows(edge_index, num_nodes, rows):
    """hop rows of sources 0..rows-1 only (int32 [rows,n]): full BFS is run on a graph restricted to what those sources
    reach by calling the C routine on the whole graph once per block is unnecessary — it computes all rows; for large n use
    the row-limited C entry point."""
    ei = np.ascontiguousarray(np.asarray(edge_index), dtype=np.int64)
    src, dst = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
    hop = np.empty((rows, num_nodes), dtype=np.int32)
    rc = _lib().gnan_oracle_apsp_rows(ctypes.c_int32(num_nodes), ctypes.c_int64(src.shape[0]),
                                      src.ctypes.data_as(ctypes.c_void_p), dst.ctypes.data_as(ctypes.c_void_p),
                                      ctypes.c_int32(rows), hop.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RuntimeError(f"gnan_oracle_apsp_rows rc={rc}")
    return hop
