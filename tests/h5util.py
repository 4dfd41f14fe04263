"""ctypes wrapper of host/libswkhost.so's HDF5 hooks (host/capi.cpp) for the tests."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "host", "libswkhost.so")
DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float32, np.float64]
_lib = None


def lib():
    global _lib
    if _lib is None:
        import subprocess

        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "host")], check=True)
        _lib = C.CDLL(LIB)
        _lib.swkh_h5_error.restype = C.c_char_p
        _lib.swkh_config_json.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    return _lib


def _ck(rc):
    if rc != 0:
        raise RuntimeError(lib().swkh_h5_error().decode())


def names(path):
    buf = C.create_string_buffer(1 << 16)
    _ck(lib().swkh_h5_names(path.encode(), buf, C.c_size_t(len(buf))))
    return [n for n in buf.value.decode().split("\n") if n]


def info(path, name):
    rank, dt, lay = C.c_int(0), C.c_int(0), C.c_int(0)
    dims = (C.c_uint64 * 8)()
    _ck(lib().swkh_h5_info(path.encode(), name.encode(), C.byref(rank), dims, C.byref(dt), C.byref(lay)))
    return tuple(dims[i] for i in range(rank.value)), DTYPES[dt.value], lay.value


def read(path, name, as_dtype=None):
    shape, dt, _ = info(path, name)
    dt = np.dtype(as_dtype or dt)
    out = np.empty(shape, dt)
    code = [np.dtype(d) for d in DTYPES].index(dt)
    _ck(lib().swkh_h5_read(path.encode(), name.encode(), code, out.ctypes.data_as(C.c_void_p), C.c_uint64(out.size)))
    return out


def write(path, datasets):
    """datasets: dict name -> numpy array (C-contiguous)"""
    arrs = [np.ascontiguousarray(a) for a in datasets.values()]
    n = len(arrs)
    nm = (C.c_char_p * n)(*[k.encode() for k in datasets])
    ranks = (C.c_int * n)(*[a.ndim for a in arrs])
    flat = [d for a in arrs for d in a.shape]
    dims = (C.c_uint64 * max(1, len(flat)))(*flat)
    dts = (C.c_int * n)(*[[np.dtype(d) for d in DTYPES].index(a.dtype) for a in arrs])
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    _ck(lib().swkh_h5_write(path.encode(), n, nm, ranks, dims, dts, ptrs))
