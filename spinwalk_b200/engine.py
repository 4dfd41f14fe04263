"""Host-side mirror of the reference's `sim` driver on raw arrays, above the C-ABI.

The reference's seam is sim::monte_carlo (src/sim/monte_carlo.cuh:10-25, monte_carlo.cu:199-355):
read config -> parameters::prepare -> per phantom {upload, per-scale kernel launches, download}.
`SimConfig` holds what config_reader leaves in `parameters` / `parameters_hvec`
(src/sim/config_reader.cpp:78-189, INI units), `Engine` is a thin object wrapper over
include/spinwalk_engine.h, and `simulate()` is the equivalent of one iteration of the phantom loop
(monte_carlo.cu:227-349) with arrays in place of HDF5 files.

Everything numeric happens in libspinwalk_b200.so (CUDA); this module only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from ._lib import MODE_COMPAT, MODE_FAST, OUT_ALL, OUT_M1, OUT_T, OUT_XYZ1, RUN_NO_PACK, RUN_NO_REBIN, RUN_NO_SORT, RUN_STATS, RUN_ZSLAB, RUN_NO_ZSLAB, RUN_NO_SHARE, RUN_NO_ONEWALK, SCALE_FOV, SCALE_GRADIENT, SCALE_PHASE_CYCLING  # noqa: F401


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"spinwalk engine error {code}: {msg}")
        self.code = code


@dataclass
class SimConfig:
    """INI-level description of one `sim` run (units as in config/config_default.ini: microseconds,
    degrees, mT/m, m^2/s, ms).  Defaults are config_default.ini's."""

    TR_us: int = 10000
    timestep_us: int = 50
    TE_us: list = field(default_factory=lambda: [5000, 6000, 7000])
    RF_FA_deg: list = field(default_factory=lambda: [15.0, 0.0, 0.0])
    RF_PH_deg: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    RF_T_us: list = field(default_factory=lambda: [0, 100000, 200000])
    dephasing_deg: list = field(default_factory=list)
    dephasing_T_us: list = field(default_factory=list)
    gradient_X_mTm: list = field(default_factory=list)
    gradient_Y_mTm: list = field(default_factory=list)
    gradient_Z_mTm: list = field(default_factory=list)
    gradient_T_us: list = field(default_factory=list)
    n_dummy_scan: int = 0
    linear_phase_cycling: float = 0.0
    quadratic_phase_cycling: float = 0.0
    B0: float = 9.4
    seed: int = 0
    n_spins: int = 100000
    cross_fov: int = 0
    record_trajectory: int = 0
    max_iterations: int = 10000
    scales: list = field(default_factory=lambda: [1.0])
    scale_type: int = SCALE_FOV
    diffusivity: list = field(default_factory=lambda: [1e-9, 1e-9])
    T1_ms: list = field(default_factory=lambda: [2200.0, 2200.0])
    T2_ms: list = field(default_factory=lambda: [41.0, 41.0])
    pXY: list = field(default_factory=lambda: [1.0, 0.0, 0.0, 1.0])

    # --- derived, as the reference derives them ---
    def timepoints(self, us):
        """config_reader::timing_scale (config_reader.cpp:39-46): integer division by TIME_STEP."""
        return [int(v) // int(self.timestep_us) for v in us]

    @property
    def n_substrate(self):
        return len(self.diffusivity)

    @property
    def n_timepoints(self):
        return int(self.TR_us) // int(self.timestep_us)

    @property
    def n_TE(self):
        return len(self.TE_us)


def _arr(x, dt):
    return np.ascontiguousarray(np.asarray(x, dtype=dt).ravel())


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class Engine:
    """One engine = one CUDA device (≙ one sim::monte_carlo object, monte_carlo.cu:33-55)."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.swk_create(int(device), C.byref(h))
        if rc != L.SWK_OK:
            raise EngineError(rc, (self._lib.swk_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.cfg = None
        self.n_local = 0
        self.n_scales = 0
        self.trj = 1
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.swk_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != L.SWK_OK:
            raise EngineError(rc, (self._lib.swk_last_error(self._h) or b"").decode())

    # ---- phantom generated on the device (include/spinwalk_phantom.h) ------------------------
    def generate_phantom(self, spec):
        """spec: spinwalk_b200.phantom_gen.PhantomSpec.  The phantom is generated in the engine's own device memory and
        becomes its current phantom; returns the generator's stats dict."""
        cs = spec.c()
        st = L.PhantomStats()
        self._ck(self._lib.swk_generate_phantom(self._h, C.byref(cs), C.byref(st)))
        n = int(spec.resolution)
        self.dims = (n, n, n)
        self.fov = (float(np.float32(spec.fov_um) * np.float32(1e-6)),) * 3
        self.has_fieldmap = spec.has_fieldmap
        return st.asdict()

    def get_phantom(self):
        """(mask uint8 [nx,ny,nz], fieldmap float32 or None) copied from the device."""
        mask = np.empty(self.dims, np.uint8)
        fm = np.empty(self.dims, np.float32) if self.has_fieldmap else None
        self._ck(self._lib.swk_get_phantom(self._h, mask.ctypes.data, None if fm is None else fm.ctypes.data))
        return mask, fm

    # ---- phantom -------------------------------------------------------------------------
    def set_phantom(self, mask, fieldmap_T, fov_m):
        """mask uint8 [nx,ny,nz], fieldmap float32 (Tesla at 1 T) or None, fov in metres.
        numpy arrays are uploaded; torch CUDA tensors are copied device-to-device."""
        dims = (C.c_uint64 * 3)(*[int(d) for d in mask.shape])
        fov = (C.c_float * 3)(*[float(np.float32(f)) for f in fov_m])
        if _is_torch_cuda(mask):
            import torch

            assert mask.dtype == torch.uint8 and mask.is_contiguous()
            fp = None
            if fieldmap_T is not None:
                assert fieldmap_T.dtype == torch.float32 and fieldmap_T.is_contiguous() and fieldmap_T.shape == mask.shape
                fp = fieldmap_T.data_ptr()
            torch.cuda.synchronize(mask.device)
            self._ck(self._lib.swk_set_phantom(self._h, mask.data_ptr(), fp, dims, fov, 1))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            f = None if fieldmap_T is None else np.ascontiguousarray(fieldmap_T, dtype=np.float32)
            if f is not None and f.shape != m.shape:
                raise ValueError("fieldmap and mask shapes differ")
            self._ck(self._lib.swk_set_phantom(self._h, m.ctypes.data, None if f is None else f.ctypes.data, dims, fov, 0))
        self.fov = tuple(float(np.float32(f)) for f in fov_m)
        self.dims = tuple(int(d) for d in mask.shape)
        self.has_fieldmap = fieldmap_T is not None

    # ---- sequence ------------------------------------------------------------------------
    def set_sequence(self, cfg: SimConfig):
        """config values -> parameters::prepare (swk_prepare) -> swk_set_sequence."""
        if cfg.seed == 0:
            raise ValueError("SEED = 0 (random seed) must be resolved by the caller")
        p = L.Params()
        p.B0 = cfg.B0
        p.linear_phase_cycling = cfg.linear_phase_cycling
        p.quadratic_phase_cycling = cfg.quadratic_phase_cycling
        p.timestep_us, p.TR_us, p.n_dummy_scan = int(cfg.timestep_us), int(cfg.TR_us), int(cfg.n_dummy_scan)
        p.n_spins, p.n_substrate = int(cfg.n_spins), cfg.n_substrate
        p.seed, p.max_iterations = int(cfg.seed), int(cfg.max_iterations)
        p.cross_fov, p.record_trajectory = int(cfg.cross_fov), int(cfg.record_trajectory)
        D = _arr(cfg.diffusivity, np.float64)
        sigma = np.zeros_like(D)
        self._ck(self._lib.swk_prepare(C.byref(p), float(cfg.RF_FA_deg[0]), float(cfg.T1_ms[0]), D.ctypes.data, len(D), sigma.ctypes.data))
        arrays = {
            "step_sigma_m": sigma, "T1_ms": _arr(cfg.T1_ms, np.float32), "T2_ms": _arr(cfg.T2_ms, np.float32),
            "pXY": _arr(cfg.pXY, np.float32), "RF_FA_deg": _arr(cfg.RF_FA_deg, np.float32),
            "RF_PH_deg": _arr(cfg.RF_PH_deg, np.float32), "RF_tp": _arr(cfg.timepoints(cfg.RF_T_us), np.int32),
            "TE_tp": _arr(cfg.timepoints(cfg.TE_us), np.int32), "dephasing_deg": _arr(cfg.dephasing_deg, np.float32),
            "dephasing_tp": _arr(cfg.timepoints(cfg.dephasing_T_us), np.int32),
            "gradX_mTm": _arr(cfg.gradient_X_mTm, np.float32), "gradY_mTm": _arr(cfg.gradient_Y_mTm, np.float32),
            "gradZ_mTm": _arr(cfg.gradient_Z_mTm, np.float32), "gradient_tp": _arr(cfg.timepoints(cfg.gradient_T_us), np.int32),
        }
        t = L.Tables()
        for (pn, nn) in L._TABLE_FIELDS:
            a = arrays[pn]
            setattr(t, pn, a.ctypes.data if a.size else None)
            setattr(t, nn, a.size)
        self._ck(self._lib.swk_set_sequence(self._h, C.byref(p), C.byref(t)))
        self.cfg = cfg
        self.params = p
        self.n_dummy_scan = p.n_dummy_scan
        self.trj = p.n_timepoints * (p.n_dummy_scan + 1) if cfg.record_trajectory else 1

    # ---- spins ---------------------------------------------------------------------------
    def set_spins(self, xyz0=None, m0=None, spin_first=0, n_local=None):
        x = None if xyz0 is None else _arr(xyz0, np.float32)
        if x is not None:
            n_local = x.size // 3
        if n_local is None:
            n_local = self.cfg.n_spins
        m = None if m0 is None else _arr(m0, np.float32)
        self._ck(self._lib.swk_set_spins(self._h, None if x is None else x.ctypes.data, None if m is None else m.ctypes.data,
                                         int(spin_first), int(n_local)))
        self.n_local = int(n_local)

    # ---- run -----------------------------------------------------------------------------
    def run_device(self, scales=None, scale_type=None, mode=MODE_FAST, flags=OUT_ALL, d_sums_ptr=None):
        """All scales in one launch; inputs already resident.  Returns the stats dict."""
        sc = _arr(self.cfg.scales if scales is None else scales, np.float32)
        st = self.cfg.scale_type if scale_type is None else scale_type
        self._ck(self._lib.swk_run_device(self._h, sc.ctypes.data, sc.size, int(st), int(mode), int(flags), d_sums_ptr))
        self.n_scales = sc.size
        self.flags = flags
        return self.stats()

    def stats(self):
        s = L.Stats()
        self._ck(self._lib.swk_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def download(self, M1=True, XYZ1=True, T=True):
        K, S, E = self.n_scales, self.n_local, self.cfg.n_TE
        m1 = np.empty((K, S, E, 3), np.float32) if (M1 and self.flags & OUT_M1) else None
        x1 = np.empty((K, S, self.trj, 3), np.float32) if (XYZ1 and self.flags & OUT_XYZ1) else None
        t = np.empty((K, S, E), np.uint8) if (T and self.flags & OUT_T) else None
        self._ck(self._lib.swk_download(self._h, *(None if a is None else a.ctypes.data for a in (m1, x1, t))))
        return m1, x1, t

    def sums(self):
        s = np.zeros((self.n_scales, self.cfg.n_TE, self.cfg.n_substrate, 4), np.float64)
        if s.size:
            self._ck(self._lib.swk_get_sums(self._h, s.ctypes.data))
        return s

    def run(self, xyz0, m0=None, spin_first=0, scales=None, scale_type=None, mode=MODE_FAST, outputs=True, stats=True, out=None):
        """swk_run: HOST buffers in, HOST buffers out (upload + kernel + download in one C call).
        `out` may carry preallocated (M1, XYZ1, T) numpy arrays (e.g. pinned)."""
        x = _arr(xyz0, np.float32)
        n_local = x.size // 3
        m = None if m0 is None else _arr(m0, np.float32)
        sc = _arr(self.cfg.scales if scales is None else scales, np.float32)
        st_ = self.cfg.scale_type if scale_type is None else scale_type
        K, E = sc.size, self.cfg.n_TE
        if out is not None:
            m1, x1, t = out
        elif outputs:
            m1 = np.empty((K, n_local, E, 3), np.float32)
            x1 = np.empty((K, n_local, self.trj, 3), np.float32)
            t = np.empty((K, n_local, E), np.uint8)
        else:
            m1 = x1 = t = None
        sums = np.zeros((K, E, self.cfg.n_substrate, 4), np.float64)
        s = L.Stats()
        self._ck(self._lib.swk_run(self._h, x.ctypes.data, None if m is None else m.ctypes.data, int(spin_first), int(n_local),
                                   sc.ctypes.data, K, int(st_), int(mode),
                                   *(None if a is None else a.ctypes.data for a in (m1, x1, t)),
                                   sums.ctypes.data if sums.size else None, C.byref(s) if stats else None))
        self.n_local, self.n_scales = n_local, K
        self.flags = (OUT_ALL if m1 is not None else 0)
        return dict(M1=m1, XYZ1=x1, T=t, sums=sums, stats=s.asdict() if stats else self.stats())

    # ---- diagnostics ---------------------------------------------------------------------
    def probe_gather(self, threads_per_sm=2048, iters=2048):
        """swk_probe_gather: random dependent 4-byte gathers per second over this engine's voxel table."""
        r, b = C.c_double(0.0), C.c_uint64(0)
        self._ck(self._lib.swk_probe_gather(self._h, int(threads_per_sm), int(iters), C.byref(r), C.byref(b)))
        return {"gathers_per_s": r.value, "table_bytes": b.value}

    # ---- plumbing ------------------------------------------------------------------------
    @property
    def stream_ptr(self):
        return self._lib.swk_stream(self._h)

    @property
    def device_bytes(self):
        return int(self._lib.swk_device_bytes(self._h))


def simulate(cfg: SimConfig, mask, fieldmap_T, fov_m, xyz0=None, m0=None, mode=MODE_FAST, device=0, spin_first=0):
    """One phantom through the engine (≙ one iteration of monte_carlo.cu:227-349).
    Returns dict(M1, XYZ1, T, sums, stats) with the reference's array layouts."""
    with Engine(device) as e:
        e.set_phantom(mask, fieldmap_T, fov_m)
        e.set_sequence(cfg)
        if xyz0 is None:
            e.set_spins(None, m0, spin_first, cfg.n_spins)
            st = e.run_device(mode=mode, flags=OUT_ALL | RUN_STATS)
            m1, x1, t = e.download()
            return dict(M1=m1, XYZ1=x1, T=t, sums=e.sums(), stats=st)
        return e.run(xyz0, m0, spin_first, mode=mode)
