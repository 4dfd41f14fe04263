"""Host mirror of the reference's `phantom` subcommand (src/phantom/handler.cpp:10-35) above include/spinwalk_phantom.h.

`PhantomSpec` carries the options of `spinwalk phantom` (src/spinwalk.cpp:58-72) with the same defaults (src/spinwalk.cpp:33-36);
`generate()` returns the arrays the reference writes to the phantom file (/mask, /fieldmap, /fov, /bvf; phantom_base.cpp:84-100).
The shape placement runs on the host (sequential RNG, as in the reference), the voxel fill on the GPU; there is no CPU path
for the fill — without a CUDA device `generate()` raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import SHAPE_CYLINDER, SHAPE_SPHERE, SHAPE_TWOPOOLS  # noqa: F401


class PhantomError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"spinwalk phantom error {code}: {msg}")
        self.code = code


@dataclass
class PhantomSpec:
    shape: int = SHAPE_CYLINDER
    fov_um: float = 1000.0
    resolution: int = 500
    dchi: float = 0.11e-6
    oxy_level: float = 0.75
    radius_um: float = 50.0
    volume_fraction: float = 4.0
    orientation_deg: float = 90.0
    seed: int = -1

    def c(self) -> L.PhantomSpec:
        return L.PhantomSpec(int(self.shape), float(self.fov_um), int(self.resolution), float(self.dchi), float(self.oxy_level),
                             float(self.radius_um), float(self.volume_fraction), float(self.orientation_deg), int(self.seed))

    @property
    def has_fieldmap(self) -> bool:
        return self.shape != SHAPE_TWOPOOLS and self.oxy_level >= 0


def _ck(lib, rc):
    if rc != L.SWK_OK:
        raise PhantomError(rc, (lib.swk_phantom_last_error() or b"").decode())


def shapes(spec: PhantomSpec) -> np.ndarray:
    """Placement only (host): float32 [n][4] = centre x, y, z and radius in µm, in acceptance order."""
    lib = L.load()
    cs = spec.c()
    n = C.c_uint32(0)
    _ck(lib, lib.swk_phantom_shapes(C.byref(cs), None, 0, C.byref(n)))
    out = np.zeros((n.value, 4), np.float32)
    if n.value:
        if spec.seed < 0:
            raise ValueError("shapes() needs a fixed seed (a random seed would place a different set on the second call)")
        _ck(lib, lib.swk_phantom_shapes(C.byref(cs), out.ctypes.data, n.value, C.byref(n)))
    return out


def generate(spec: PhantomSpec, device: int = 0, out=None, out_host=None):
    """Returns (mask uint8 [n,n,n], fieldmap float32 [n,n,n] or None, fov_m float32[3], stats dict).

    out_host = (mask, fieldmap) C-contiguous numpy arrays of those shapes / dtypes to fill instead of allocating new ones.

    out = (mask, fieldmap) torch CUDA tensors (uint8 / float32, contiguous, [n,n,n]) makes the generator write into them on the
    device instead of allocating numpy arrays (fieldmap may be None when the spec has no field map)."""
    lib = L.load()
    n = int(spec.resolution)
    cs = spec.c()
    st = L.PhantomStats()
    fov = np.full(3, np.float32(spec.fov_um) * np.float32(1e-6), np.float32)  # phantom_base.cpp:63
    if out is not None:
        import torch

        mask, fm = out
        assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and tuple(mask.shape) == (n, n, n)
        fp = None
        if spec.has_fieldmap:
            assert fm is not None and fm.is_cuda and fm.dtype == torch.float32 and fm.is_contiguous() and tuple(fm.shape) == (n, n, n)
            fp = fm.data_ptr()
        torch.cuda.synchronize(mask.device)
        _ck(lib, lib.swk_phantom_generate(mask.device.index or 0, C.byref(cs), mask.data_ptr(), fp, 1, C.byref(st)))
        return mask, (fm if spec.has_fieldmap else None), fov, st.asdict()
    if out_host is not None:
        mask, fm = out_host
        assert mask.dtype == np.uint8 and mask.shape == (n, n, n) and mask.flags.c_contiguous
        if spec.has_fieldmap:
            assert fm is not None and fm.dtype == np.float32 and fm.shape == (n, n, n) and fm.flags.c_contiguous
        else:
            fm = None
    else:
        mask = np.empty((n, n, n), np.uint8)
        fm = np.empty((n, n, n), np.float32) if spec.has_fieldmap else None
    _ck(lib, lib.swk_phantom_generate(int(device), C.byref(cs), mask.ctypes.data, None if fm is None else fm.ctypes.data, 0, C.byref(st)))
    return mask, fm, fov, st.asdict()


def generate_mesh(fov_um: float, resolution: int, vertices, faces, device: int = 0):
    """`spinwalk phantom -p`: mask of a closed triangle mesh centred in the FoV (≙ phantom::ply::run, src/phantom/phantom_ply.cpp:141-227).
    vertices float64 [nv,3] in the PLY file's unit (mm), faces [nf,3].  Returns (mask uint8 [n,n,n], fov_m float32[3], stats dict)."""
    lib = L.load()
    v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    f = np.ascontiguousarray(faces, np.uint64).reshape(-1, 3)
    n = int(resolution)
    mask = np.empty((n, n, n), np.uint8)
    st = L.PhantomStats()
    _ck(lib, lib.swk_phantom_mesh(int(device), float(fov_um), n, v.ctypes.data, len(v), f.ctypes.data, len(f), mask.ctypes.data, 0, C.byref(st)))
    fov = np.full(3, np.float32(fov_um) * np.float32(1e-6), np.float32)
    return mask, fov, st.asdict()
