"""ctypes binding of libcagroup3d_b200.so (the C ABI in include/cagroup3d_b200.h).

The argument types of every entry point are parsed from the header itself, so the binding cannot
drift from the declared ABI.  There is no fallback: if the library is missing, a declared symbol is
not exported, or a call returns a non-zero status, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcagroup3d_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "cagroup3d_b200.h")

_lib = None
_host_only = set()          # entry points without a trailing `void* stream`


# pointee type of a pointer parameter -> tensor dtypes it may point at (None: untyped, e.g. `void*`)
_POINTEE = {
    "float": (torch.float32,),
    "int": (torch.int32,),
    "unsigned": (torch.int32,),
    "unsigned int": (torch.int32,),
    "unsigned short": (torch.int16, torch.bfloat16, torch.float16, torch.uint16),
    "unsigned char": (torch.uint8, torch.int8, torch.bool),
    "long long": (torch.int64,),
    "unsigned long long": (torch.int64,),
    "void": None,
}
_pointees = {}              # name -> [dtype tuple | None | "scalar"] per parameter


def parse_header(path: str = HEADER_PATH):
    """-> {name: [ctypes types]} for every `int cg3d_*(...)` prototype.  The pointee type of every pointer parameter is
    kept as well (`_pointees`): call() checks device, dtype and unit inner stride of each tensor against it."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(cg3d_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(1), m.group(2)
        types, pts = [], []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if a == "void":
                continue
            if "*" in a:
                types.append(ctypes.c_void_p)
                base = re.sub(r"\bconst\b", "", a.split("*")[0]).strip()
                if base not in _POINTEE:
                    raise RuntimeError(f"unparsed pointee '{base}' of {name}")
                pts.append(_POINTEE[base])
            elif a.startswith("float"):
                types.append(ctypes.c_float)
                pts.append("scalar")
            elif a.startswith("long long"):
                types.append(ctypes.c_longlong)
                pts.append("scalar")
            elif a.startswith("int"):
                types.append(ctypes.c_int)
                pts.append("scalar")
            else:
                raise RuntimeError(f"unparsed parameter '{a}' of {name}")
        protos[name] = types
        _pointees[name] = pts
        if not args.strip().endswith("stream"):
            _host_only.add(name)
    return protos


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cagroup3d_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, types in parse_header().items():
            fn = getattr(lib, name)           # AttributeError if the library lacks a declared symbol
            fn.argtypes = types
            fn.restype = ctypes.c_int
        _lib = lib
    return _lib


def _arg(a, name="", i=0, want=None):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        # a wrong device / dtype / stride would be a silent out-of-bounds access on the GPU, not an error
        if not a.is_cuda:
            raise TypeError(f"{name}: argument {i} is a {a.device} tensor; the C ABI takes device pointers")
        if want == "scalar":
            raise TypeError(f"{name}: argument {i} is a tensor where the header declares a scalar")
        if want is not None and a.dtype not in want:
            raise TypeError(f"{name}: argument {i} has dtype {a.dtype}, the header declares {want}")
        if a.dim() >= 1 and a.stride(-1) != 1 and a.shape[-1] != 1 and a.numel() > 1:
            raise TypeError(f"{name}: argument {i} has inner stride {a.stride(-1)}; rows must be contiguous")
        return a.data_ptr()
    return a


# torch.cuda.current_stream() builds a Stream object through three layers of device-index helpers (~5 us, 12 % of the host
# time of a training step); the raw handle of the current stream is one C call
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)


def stream_ptr() -> int:
    if _raw_stream is not None and _get_device is not None:
        return _raw_stream(_get_device())
    return torch.cuda.current_stream().cuda_stream


_bound = {}                 # name -> (ctypes function, pointee table)


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on the current CUDA stream; raise on a non-zero status."""
    ent = _bound.get(name)
    if ent is None:
        ent = _bound[name] = (getattr(load(), name), tuple(_pointees.get(name, ())))
    fn, pts = ent
    npts = len(pts)
    rc = fn(*[_arg(a, name, i, pts[i] if i < npts else None) for i, a in enumerate(args)], stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}")


def host(name: str, *args) -> int:
    """Host-only helper entry points (sizes of workspaces); returns the int result."""
    return getattr(load(), name)(*args)


def hash_capacity(n: int) -> int:
    return host("cg3d_hash_capacity", int(n))


def scan_workspace_ints(n: int) -> int:
    return host("cg3d_scan_workspace_ints", int(n))


def sort_workspace_ints(n: int) -> int:
    return host("cg3d_sort_workspace_ints", int(n))
