"""oracle/pyoracle.py — TEST INFRASTRUCTURE, not product code.

ctypes front-end to
  * oracle/liboracle.so            (our plain-C restatement, oracle/sim_oracle.c)
  * oracle/_ref/libswref_cpu.so    (reference kernels.cu built by g++,  mt19937 flavour)
  * oracle/_ref/libswref_cuda.so   (reference kernels.cu built by nvcc, minstd flavour; host
                                    instantiation of sim::sim + launcher of the reference cu_sim)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (spinwalk_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF_CPU = os.path.join(HERE, "_ref", "libswref_cpu.so")
LIB_REF_CUDA = os.path.join(HERE, "_ref", "libswref_cuda.so")

RNG_MT19937, RNG_MINSTD = 0, 1
SCALE_FOV, SCALE_GRADIENT, SCALE_PHASE = 0, 1, 2


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference exists, oracle/_ref (see oracle/Makefile)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class _CCase(C.Structure):
    _fields_ = [
        ("fov", C.c_double * 3),
        ("phantom_size", C.c_uint64 * 3),
        ("seed", C.c_uint64),
        ("max_iterations", C.c_uint64),
        ("B0", C.c_float),
        ("linear_phase_cycling", C.c_float),
        ("quadratic_phase_cycling", C.c_float),
        ("timestep_us", C.c_int32),
        ("TR_us", C.c_int32),
        ("n_dummy_scan", C.c_int32),
        ("n_spins", C.c_uint32),
        ("n_substrate", C.c_uint32),
        ("cross_fov", C.c_int32),
        ("record_trajectory", C.c_int32),
        ("diffusivity", C.c_void_p),
        ("T1_ms", C.c_void_p),
        ("T2_ms", C.c_void_p),
        ("pXY", C.c_void_p),
        ("RF_FA_deg", C.c_void_p),
        ("RF_PH_deg", C.c_void_p),
        ("RF_tp", C.c_void_p),
        ("n_RF", C.c_uint32),
        ("TE_tp", C.c_void_p),
        ("n_TE", C.c_uint32),
        ("dephasing_deg", C.c_void_p),
        ("dephasing_tp", C.c_void_p),
        ("n_dephasing", C.c_uint32),
        ("gradX_mTm", C.c_void_p),
        ("gradY_mTm", C.c_void_p),
        ("gradZ_mTm", C.c_void_p),
        ("gradient_tp", C.c_void_p),
        ("n_gradient", C.c_uint32),
        ("scales", C.c_void_p),
        ("n_scales", C.c_uint32),
        ("scale_type", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("steps", "mask_gathers", "field_gathers", "rejects", "lost")]

    def asdict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).ravel())


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).ravel())


@dataclass
class Case:
    """A `sim` case in the units the reference holds after config_reader::prepare
    (times already in TIMEPOINTS, diffusivity in m^2/s) — see oracle/sim_case.h."""

    fov: tuple
    phantom_size: tuple
    n_spins: int
    TR_us: int
    timestep_us: int = 50
    seed: int = 10
    max_iterations: int = 10000
    B0: float = 9.4
    linear_phase_cycling: float = 0.0
    quadratic_phase_cycling: float = 0.0
    n_dummy_scan: int = 0
    cross_fov: int = 0
    record_trajectory: int = 0
    diffusivity: list = field(default_factory=lambda: [1e-9, 1e-9])
    T1_ms: list = field(default_factory=lambda: [2200.0, 2200.0])
    T2_ms: list = field(default_factory=lambda: [41.0, 41.0])
    pXY: list = field(default_factory=lambda: [1.0, 0.0, 0.0, 1.0])
    RF_FA_deg: list = field(default_factory=lambda: [90.0])
    RF_PH_deg: list = field(default_factory=lambda: [0.0])
    RF_tp: list = field(default_factory=lambda: [0])
    TE_tp: list = field(default_factory=lambda: [400])
    dephasing_deg: list = field(default_factory=list)
    dephasing_tp: list = field(default_factory=list)
    gradX_mTm: list = field(default_factory=list)
    gradY_mTm: list = field(default_factory=list)
    gradZ_mTm: list = field(default_factory=list)
    gradient_tp: list = field(default_factory=list)
    scales: list = field(default_factory=lambda: [1.0])
    scale_type: int = SCALE_FOV

    # ---- derived (parameters::prepare, simulation_parameters.cuh:227-245) ----
    @property
    def n_substrate(self):
        return len(self.diffusivity)

    @property
    def n_timepoints(self):
        return int(self.TR_us) // int(self.timestep_us)

    @property
    def n_dummy(self):
        if self.n_dummy_scan >= 0:
            return self.n_dummy_scan
        return int(5.0 * float(np.float32(self.T1_ms[0])) / float(np.float32(self.TR_us * 1e-3)))

    @property
    def trj(self):
        return self.n_timepoints * (self.n_dummy + 1) if self.record_trajectory else 1

    @property
    def n_TE(self):
        return len(self.TE_tp)

    @property
    def n_scales(self):
        return len(self.scales)

    def total_steps(self):
        """accepted spin-steps of the whole run: S*K*R*Nt."""
        return self.n_spins * self.n_scales * (self.n_dummy + 1) * self.n_timepoints

    def to_c(self):
        """Returns (ctypes struct, keepalive list of numpy arrays)."""
        keep = {
            "diffusivity": np.ascontiguousarray(np.asarray(self.diffusivity, dtype=np.float64)),
            "T1_ms": _f32(self.T1_ms),
            "T2_ms": _f32(self.T2_ms),
            "pXY": _f32(self.pXY),
            "RF_FA_deg": _f32(self.RF_FA_deg),
            "RF_PH_deg": _f32(self.RF_PH_deg),
            "RF_tp": _i32(self.RF_tp),
            "TE_tp": _i32(self.TE_tp),
            "dephasing_deg": _f32(self.dephasing_deg),
            "dephasing_tp": _i32(self.dephasing_tp),
            "gradX_mTm": _f32(self.gradX_mTm),
            "gradY_mTm": _f32(self.gradY_mTm),
            "gradZ_mTm": _f32(self.gradZ_mTm),
            "gradient_tp": _i32(self.gradient_tp),
            "scales": _f32(self.scales),
        }
        assert len(keep["pXY"]) == self.n_substrate**2
        assert len(keep["RF_FA_deg"]) == len(keep["RF_PH_deg"]) == len(keep["RF_tp"]) >= 1
        assert len(keep["gradX_mTm"]) == len(keep["gradY_mTm"]) == len(keep["gradZ_mTm"]) == len(keep["gradient_tp"])
        assert len(keep["dephasing_deg"]) == len(keep["dephasing_tp"])
        assert self.seed != 0
        c = _CCase()
        for i in range(3):
            c.fov[i] = float(np.float32(self.fov[i]))
            c.phantom_size[i] = int(self.phantom_size[i])
        for k in (
            "seed max_iterations B0 linear_phase_cycling quadratic_phase_cycling timestep_us TR_us "
            "n_dummy_scan n_spins cross_fov record_trajectory scale_type"
        ).split():
            setattr(c, k, getattr(self, k))
        c.n_substrate = self.n_substrate
        for k, a in keep.items():
            setattr(c, k, a.ctypes.data if a.size else None)
        c.n_RF = len(keep["RF_tp"])
        c.n_TE = len(keep["TE_tp"])
        c.n_dephasing = len(keep["dephasing_tp"])
        c.n_gradient = len(keep["gradient_tp"])
        c.n_scales = len(keep["scales"])
        return c, keep

    def alloc_outputs(self):
        K, S, E = self.n_scales, self.n_spins, self.n_TE
        return (
            np.zeros((K, S, E, 3), np.float32),
            np.zeros((K, S, self.trj, 3), np.float32),
            np.zeros((K, S, E), np.uint8),
        )


def default_m0(n):
    m = np.zeros((n, 3), np.float32)
    m[:, 2] = 1.0  # monte_carlo.cu:162-164
    return m


_libs = {}


def _lib(path):
    if path not in _libs:
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (oracle.pyoracle.build())")
        _libs[path] = C.CDLL(path)
    return _libs[path]


def have_ref_cpu():
    return os.path.exists(LIB_REF_CPU)


def have_ref_cuda():
    return os.path.exists(LIB_REF_CUDA)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _prep_inputs(case, fieldmap_T, mask, xyz0, m0):
    V = int(np.prod(case.phantom_size))
    mask = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
    assert mask.size == V
    if fieldmap_T is not None:
        fieldmap_T = _f32(fieldmap_T)
        assert fieldmap_T.size == V
    xyz0 = _f32(xyz0)
    assert xyz0.size == 3 * case.n_spins
    m0 = _f32(default_m0(case.n_spins) if m0 is None else m0)
    return fieldmap_T, mask, xyz0, m0


def run_oracle(case: Case, fieldmap_T, mask, xyz0, m0=None, flavour=RNG_MINSTD, threads=None, spins=None):
    """Run the C restatement.  Returns dict(M1, XYZ1, T, stats, seconds)."""
    lib = _lib(LIB_ORACLE)
    fieldmap_T, mask, xyz0, m0 = _prep_inputs(case, fieldmap_T, mask, xyz0, m0)
    cc, keep = case.to_c()
    M1, XYZ1, T = case.alloc_outputs()
    st, sec = Stats(), C.c_double(0)
    b, e = spins if spins else (0, case.n_spins)
    threads = threads or os.cpu_count() or 1
    lib.swo_run.restype = C.c_int
    rc = lib.swo_run(C.byref(cc), _ptr(fieldmap_T), _ptr(mask), _ptr(xyz0), _ptr(m0), _ptr(M1), _ptr(XYZ1), _ptr(T),
                     C.c_uint32(b), C.c_uint32(e), C.c_int(flavour), C.c_int(threads), C.byref(st), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"swo_run failed rc={rc}")
    del keep
    return dict(M1=M1, XYZ1=XYZ1, T=T, stats=st.asdict(), seconds=sec.value)


def run_ref(case: Case, fieldmap_T, mask, xyz0, m0=None, flavour=RNG_MINSTD, threads=None, spins=None):
    """Run the UNMODIFIED reference sim::sim on the host (oracle/_ref)."""
    lib = _lib(LIB_REF_CPU if flavour == RNG_MT19937 else LIB_REF_CUDA)
    assert lib.swref_flavour() == flavour
    fieldmap_T, mask, xyz0, m0 = _prep_inputs(case, fieldmap_T, mask, xyz0, m0)
    cc, keep = case.to_c()
    M1, XYZ1, T = case.alloc_outputs()
    sec = C.c_double(0)
    b, e = spins if spins else (0, case.n_spins)
    threads = threads or os.cpu_count() or 1
    lib.swref_run.restype = C.c_int
    rc = lib.swref_run(C.byref(cc), _ptr(fieldmap_T), _ptr(mask), _ptr(xyz0), _ptr(m0), _ptr(M1), _ptr(XYZ1), _ptr(T),
                       C.c_uint32(b), C.c_uint32(e), C.c_int(threads), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"swref_run failed rc={rc}")
    del keep
    return dict(M1=M1, XYZ1=XYZ1, T=T, seconds=sec.value)


def run_ref_cuda(case: Case, fieldmap_T, mask, xyz0, m0=None, device=0):
    """Launch the reference's own __global__ cu_sim (compiled for sm_100a) on a GPU."""
    lib = _lib(LIB_REF_CUDA)
    fieldmap_T, mask, xyz0, m0 = _prep_inputs(case, fieldmap_T, mask, xyz0, m0)
    cc, keep = case.to_c()
    M1, XYZ1, T = case.alloc_outputs()
    ms = C.c_float(0)
    lib.swref_cuda_run.restype = C.c_int
    rc = lib.swref_cuda_run(C.byref(cc), _ptr(fieldmap_T), _ptr(mask), _ptr(xyz0), _ptr(m0), _ptr(M1), _ptr(XYZ1),
                            _ptr(T), C.c_int(device), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"swref_cuda_run failed: cudaError {rc}")
    del keep
    return dict(M1=M1, XYZ1=XYZ1, T=T, kernel_ms=ms.value)


def init_positions(seed, fov, n_spins, impl="oracle"):
    """monte_carlo.cu:142-151 default XYZ0.  impl: 'oracle' (C restatement) or 'ref' (libstdc++)."""
    f = np.asarray(fov, dtype=np.float32)
    out = np.zeros((n_spins, 3), np.float32)
    if impl == "oracle":
        _lib(LIB_ORACLE).swo_init_positions(C.c_uint64(seed), _ptr(f), C.c_uint32(n_spins), _ptr(out))
    else:
        _lib(LIB_REF_CPU).swref_init_positions(C.c_uint64(seed), _ptr(f), C.c_uint32(n_spins), _ptr(out))
    return out
