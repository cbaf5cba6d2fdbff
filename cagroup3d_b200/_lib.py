"""ctypes binding of libcagroup3d_b200.so (the C ABI in include/cagroup3d_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcagroup3d_b200.so")

# signature spec: p = device/host pointer, i = int, f = float, l = long long
SIGNATURES = {
    "cg3d_hash_capacity": "i",
    "cg3d_scan_workspace_ints": "i",
    "cg3d_quantize": "piifffippp",
    "cg3d_stride_coords": "piipp",
    "cg3d_exclusive_scan_i32": "pipppp",
    "cg3d_unique_first": "pippippppp" + "p",
    "cg3d_hash_build": "pippip",
    "cg3d_hash_lookup": "pippipp",
    "cg3d_neighbor_table": "pippiiipp",
    "cg3d_transpose_table": "pippiiipp",
    "cg3d_count_rules": "plpp",
    "cg3d_spconv_simt": "ppppiiiipppipppip",
    "cg3d_spconv_tc": "ppppiiiipppipppip",
    "cg3d_affine_act": "ppppplii" + "p",
    "cg3d_interp_trilinear": "pippiipippp",
    "cg3d_avgpool_window": "pipiipipp",
    "cg3d_segment_mean": "pipippiiippp",
    "cg3d_gather_rows": "piipiifpp",
    "cg3d_coord_bounds": "pipp",
    "cg3d_vote_points": "ppiifippp",
    "cg3d_semantic_flags": "piifpp",
    "cg3d_compact_rows": "ppiipp",
    "cg3d_class_points": "pppppppp" + "iiiii" + "f" + "pppp",
    "cg3d_head_decode": "pipiiiippppp" + "ip",
    "cg3d_boxes_pairwise_bev": "pipiipp",
    "cg3d_nms_segments": "ppiifippp",
    "cg3d_roi_grid_coords": "piiiifiipp",
    "cg3d_roi_pool_table": "piipp",
    "cg3d_roi_decode": "ppiiipp",
}
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "l": ctypes.c_longlong}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cagroup3d_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, sig in SIGNATURES.items():
            if not hasattr(lib, name):
                continue                      # optional symbol (checked by exported_symbols test)
            fn = getattr(lib, name)
            fn.argtypes = [_CT[c] for c in sig] if name not in ("cg3d_hash_capacity", "cg3d_scan_workspace_ints") \
                else [ctypes.c_int]
            fn.restype = ctypes.c_int
        _lib = lib
    return _lib


def _arg(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on the current CUDA stream; raise on a non-zero status."""
    fn = getattr(load(), name)
    rc = fn(*[_arg(a) for a in args], stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}")


def hash_capacity(n: int) -> int:
    return load().cg3d_hash_capacity(int(n))


def scan_workspace_ints(n: int) -> int:
    return load().cg3d_scan_workspace_ints(int(n))
