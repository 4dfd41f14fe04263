"""oracle/pyphantom.py — TEST INFRASTRUCTURE, not product code.

ctypes front-end to the phantom / dwi / config generator oracles:
  * oracle/libphantom_oracle.so     plain-C restatement of `spinwalk phantom` (oracle/phantom_oracle.c)
  * oracle/_ref/libswref_gen.so     the reference's own generators, compiled unmodified (oracle/ref_gen_harness.cpp), serial
  * oracle/_ref/libswref_gen_omp.so the same with OpenMP (the reference's CMake links it when found): bench.py's CPU baseline

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "libphantom_oracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libswref_gen.so")
LIB_REF_OMP = os.path.join(HERE, "_ref", "libswref_gen_omp.so")

CYLINDER, SPHERE, TWOPOOLS = 0, 1, 2


class _Spec(C.Structure):
    _fields_ = [("shape", C.c_int32), ("fov_um", C.c_float), ("resolution", C.c_uint64), ("dchi", C.c_float), ("Y", C.c_float),
                ("radius_um", C.c_float), ("volume_fraction", C.c_float), ("orientation_deg", C.c_float), ("seed", C.c_int32)]


@dataclass
class Phantom:
    mask: np.ndarray            # uint8 [n,n,nzw]
    fieldmap: np.ndarray | None  # float32 [n,n,nzw]
    bvf: float | None
    shapes: np.ndarray          # float32 [n_shapes,4]


def have_ref(omp: bool = False) -> bool:
    return os.path.exists(LIB_REF_OMP if omp else LIB_REF)


def _args(kw):
    d = dict(shape=CYLINDER, fov_um=100.0, resolution=64, dchi=0.11e-6, Y=0.78, radius_um=8.0, volume_fraction=4.0, orientation_deg=90.0, seed=0)
    d.update(kw)
    return d


def oracle(zwin=None, **kw) -> Phantom:
    """C restatement.  zwin=(zlo, zhi) computes those z slices only (outputs [n,n,zhi-zlo], bvf None)."""
    a = _args(kw)
    lib = C.CDLL(LIB_ORACLE)
    n = int(a["resolution"])
    zlo, zhi = (0, n) if zwin is None else zwin
    sp = _Spec(a["shape"], a["fov_um"], n, a["dchi"], a["Y"], a["radius_um"], a["volume_fraction"], a["orientation_deg"], a["seed"])
    calc = a["Y"] >= 0 and a["shape"] != TWOPOOLS
    mask = np.zeros((n, n, zhi - zlo), np.uint8)
    fm = np.zeros((n, n, zhi - zlo), np.float32) if calc else None
    cap = 1 << 18
    shapes = np.zeros((cap, 4), np.float32)
    ns = C.c_uint32(0)
    bvf = C.c_float(0)
    fp = None if fm is None else fm.ctypes.data_as(C.c_void_p)
    if zwin is None:
        rc = lib.swo_phantom_generate(C.byref(sp), mask.ctypes.data_as(C.c_void_p), fp, C.byref(bvf), shapes.ctypes.data_as(C.c_void_p), cap, C.byref(ns))
    else:
        rc = lib.swo_phantom_generate_window(C.byref(sp), int(zlo), int(zhi), mask.ctypes.data_as(C.c_void_p), fp, shapes.ctypes.data_as(C.c_void_p), cap, C.byref(ns))
    if rc != 0:
        raise RuntimeError(f"phantom oracle refused the spec (rc={rc})")
    return Phantom(mask, fm, bvf.value if zwin is None else None, shapes[: min(ns.value, cap)].copy())


def oracle_shapes(**kw) -> np.ndarray:
    a = _args(kw)
    lib = C.CDLL(LIB_ORACLE)
    sp = _Spec(a["shape"], a["fov_um"], int(a["resolution"]), a["dchi"], a["Y"], a["radius_um"], a["volume_fraction"], a["orientation_deg"], a["seed"])
    cap = 1 << 18
    shapes = np.zeros((cap, 4), np.float32)
    ns = C.c_uint32(0)
    rc = lib.swo_phantom_shapes(C.byref(sp), shapes.ctypes.data_as(C.c_void_p), cap, C.byref(ns))
    if rc != 0:
        raise RuntimeError(f"phantom oracle refused the spec (rc={rc})")
    return shapes[: min(ns.value, cap)].copy()


def reference(omp: bool = False, **kw) -> Phantom:
    """The reference's own generator classes (phantom::cylinder / sphere / twopools ::run(false))."""
    a = _args(kw)
    lib = C.CDLL(LIB_REF_OMP if omp else LIB_REF)
    n = int(a["resolution"])
    calc = a["Y"] >= 0 and a["shape"] != TWOPOOLS
    mask = np.zeros((n, n, n), np.uint8)
    fm = np.zeros((n, n, n), np.float32) if calc else None
    cap = 1 << 18
    shapes = np.zeros((cap, 4), np.float32)
    ns = C.c_uint32(0)
    bvf = C.c_float(0)
    rc = lib.swref_phantom(int(a["shape"]), C.c_float(a["fov_um"]), C.c_uint64(n), C.c_float(a["dchi"]), C.c_float(a["Y"]), C.c_float(a["radius_um"]),
                           C.c_float(a["volume_fraction"]), C.c_float(a["orientation_deg"]), C.c_int32(a["seed"]), mask.ctypes.data_as(C.c_void_p),
                           None if fm is None else fm.ctypes.data_as(C.c_void_p), C.byref(bvf), shapes.ctypes.data_as(C.c_void_p), cap, C.byref(ns))
    if rc != 0:
        raise RuntimeError(f"reference phantom generator failed (rc={rc})")
    return Phantom(mask, fm, bvf.value, shapes[: min(ns.value, cap)].copy())


def reference_dwi(config_path: str, b_values, direction, start_ms: int, delta_ms: int, DELTA_ms: int) -> bool:
    """`spinwalk dwi` of the reference: edits config_path in place."""
    lib = C.CDLL(LIB_REF)
    b = (C.c_double * len(b_values))(*[float(v) for v in b_values])
    d = (C.c_float * 3)(*[float(v) for v in direction])
    return lib.swref_dwi(b, len(b_values), d, int(start_ms), int(delta_ms), int(DELTA_ms), config_path.encode()) == 0


def reference_config(seq_name: str, TE_us: int, timestep_us: int, phantoms, output: str) -> bool:
    """`spinwalk config` of the reference: writes output and default_config.ini next to it."""
    lib = C.CDLL(LIB_REF)
    arr = (C.c_char_p * len(phantoms))(*[p.encode() for p in phantoms])
    return lib.swref_config(seq_name.encode(), int(TE_us), int(timestep_us), arr, len(phantoms), output.encode()) == 0


def oracle_mesh(fov_um: float, resolution: int, vertices, faces) -> np.ndarray:
    """C restatement of `spinwalk phantom -p`: vertices float64 [nv,3] in the PLY file's unit (mm), faces [nf,3]."""
    lib = C.CDLL(LIB_ORACLE)
    v = np.ascontiguousarray(vertices, np.float64)
    f = np.ascontiguousarray(faces, np.uint64)
    n = int(resolution)
    mask = np.zeros((n, n, n), np.uint8)
    rc = lib.swo_phantom_mesh(C.c_float(fov_um), C.c_uint64(n), v.ctypes.data_as(C.c_void_p), C.c_uint64(len(v)), f.ctypes.data_as(C.c_void_p), C.c_uint64(len(f)),
                              mask.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"mesh oracle refused the input (rc={rc})")
    return mask


def reference_mesh(fov_um: float, resolution: int, ply_path: str):
    """phantom::ply(fov, resolution, ..., ply_path).run(false) of the reference: (mask, bvf)."""
    lib = C.CDLL(LIB_REF)
    n = int(resolution)
    mask = np.zeros((n, n, n), np.uint8)
    bvf = C.c_float(-1)
    rc = lib.swref_phantom_ply(C.c_float(fov_um), C.c_uint64(n), ply_path.encode(), mask.ctypes.data_as(C.c_void_p), C.byref(bvf))
    if rc != 0:
        raise RuntimeError(f"reference ply phantom failed (rc={rc})")
    return mask, bvf.value
