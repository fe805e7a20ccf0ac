"""ctypes binding of ``libpnslam.so`` (the C ABI declared in ``include/pnslam.h``).

There is deliberately no fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised (the reference's error convention is Python
exceptions only).  PyTorch is used by the callers of this module
only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpnslam.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

P = C.c_void_p


class PnGrid(C.Structure):
    _fields_ = [("data", P), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int)]


class PnGridMlp(C.Structure):
    _fields_ = [("B", P), ("W", P * 5), ("b", P * 5), ("Wc", P * 5), ("bc", P * 5), ("Wo", P), ("bo", P),
                ("c_dim", C.c_int), ("n_out", C.c_int)]


class PnGridMlpGrad(C.Structure):
    _fields_ = [("B", P), ("W", P * 5), ("b", P * 5), ("Wc", P * 5), ("bc", P * 5), ("Wo", P), ("bo", P)]


class PnCoarseMlp(C.Structure):
    _fields_ = [("W", P * 5), ("b", P * 5), ("Wo", P), ("bo", P)]


class PnCoarseMlpGrad(C.Structure):
    _fields_ = [("W", P * 5), ("b", P * 5), ("Wo", P), ("bo", P)]


class PnImapMlp(C.Structure):
    _fields_ = [("B", P), ("W", P * 8), ("b", P * 8), ("Wo", P), ("bo", P), ("hidden", C.c_int), ("n_blocks", C.c_int)]


class PnPoints(C.Structure):
    _fields_ = [("pts64", P), ("pts32", P), ("rays_o", P), ("rays_d", P), ("z", P), ("S", C.c_int), ("N", C.c_int64)]


class PnStash(C.Structure):
    _fields_ = [("relu_bits", P), ("H", P), ("C", P), ("E", P)]


class PnWscratch(C.Structure):
    _fields_ = [("GA", P), ("GH", P), ("GARG", P), ("P32", P), ("GO", P)]


OUT_SET_ALL, OUT_SET_W, OUT_ADD_W, OUT_SET_RGB = 0, 1, 2, 3

_lib: Optional[C.CDLL] = None
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pnslam.h")


def _declare_prototypes(lib: C.CDLL) -> None:
    """Set ctypes argtypes for every ``int pn_*(...)`` declared in include/pnslam.h so
    that a call with the wrong number or kind of arguments raises instead of
    corrupting the stack."""
    import re
    if not os.path.exists(HEADER_PATH):
        return
    src = re.sub(r"/\*.*?\*/", "", open(HEADER_PATH).read(), flags=re.S)
    for m in re.finditer(r"\bint\s+(pn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        if not hasattr(lib, name):
            continue
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(C.c_void_p)
                elif "int64_t" in a:
                    types.append(C.c_int64)
                elif a.startswith("float") or " float " in " " + a:
                    types.append(C.c_float)
                elif a.startswith("double"):
                    types.append(C.c_double)
                else:
                    types.append(C.c_int)
        fn = getattr(lib, name)
        fn.argtypes = types
        fn.restype = C.c_int


def lib() -> C.CDLL:
    """Load the CUDA library once per process; raise loudly if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                f"Build it with `bash {BUILD_SCRIPT}` or `python -c 'import __graft_entry__ as g; g.build()'`.")
        _lib = C.CDLL(LIB_PATH)
        _declare_prototypes(_lib)
        _lib.pn_last_error.restype = C.c_char_p
        _lib.pn_version.restype = C.c_int
        _lib.pn_launch_count.restype = C.c_longlong
    return _lib


# Optional per-kernel timing for bench.py: when PROFILE is a dict, every checked
# C-ABI call is bracketed by CUDA events on the current stream and the pairs are
# appended under the call's name.  None (default) = no events.
PROFILE = None


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed: {lib().pn_last_error().decode()}")


class timed:
    """with timed('name', device): ... records an event pair when PROFILE is on."""

    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream(self.device))
            PROFILE.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def _dev_index(device) -> int:
    if isinstance(device, int):
        return device
    idx = getattr(device, "index", None)
    if idx is None:
        idx = torch.device(device).index
    return torch.cuda.current_device() if idx is None else idx


def stream_ptr(device) -> int:
    """Raw cudaStream_t of torch's CURRENT stream on the device (queried every call: callers
    may switch streams between calls)."""
    return torch._C._cuda_getCurrentRawStream(_dev_index(device))


class device_guard:
    """``with device_guard(dev):`` makes dev the current CUDA device only when it is not already
    (the common case costs one integer comparison instead of two cudaSetDevice calls)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = _dev_index(device)
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False



_bound_cache = {}


def f64x6(bound) -> C.Array:
    """(3,2) bound -> host double[6] = [lo_x hi_x lo_y hi_y lo_z hi_z].  Host-side values are
    memoised per tensor object and in-place version (a bound is set once by the orchestrator)."""
    if isinstance(bound, torch.Tensor):
        key = (id(bound), bound._version, bound.data_ptr())
        hit = _bound_cache.get(key)
        if hit is not None:
            return hit
        vals = bound.detach().to("cpu", torch.float64).reshape(-1).tolist()
        if len(_bound_cache) > 64:
            _bound_cache.clear()
        arr = (C.c_double * 6)(*vals)
        _bound_cache[key] = arr
        return arr
    vals = [float(v) for row in bound for v in row]
    return (C.c_double * 6)(*vals)


def fill_ptr_array(arr, tensors: Sequence[Optional[torch.Tensor]]) -> None:
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
