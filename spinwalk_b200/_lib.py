"""ctypes declarations of include/spinwalk_engine.h and include/spinwalk_phantom.h (the C-ABI).  No torch, no numpy types cross it."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SPINWALK_B200_LIB selects another build of the SAME library (tuning experiments: launch bounds etc.)
LIB_PATH = os.environ.get("SPINWALK_B200_LIB") or os.path.join(HERE, "libspinwalk_b200.so")

SWK_OK, SWK_ERR_INVALID, SWK_ERR_CUDA, SWK_ERR_MEMORY, SWK_ERR_STATE, SWK_ERR_SUBSTRATE = range(6)
SCALE_FOV, SCALE_GRADIENT, SCALE_PHASE_CYCLING = 0, 1, 2
MODE_COMPAT, MODE_FAST = 0, 1
OUT_M1, OUT_XYZ1, OUT_T, OUT_ALL, RUN_STATS, RUN_NO_SORT, RUN_NO_PACK, RUN_NO_REBIN = 1, 2, 4, 7, 16, 32, 64, 128
RUN_ZSLAB, RUN_NO_ZSLAB, RUN_NO_SHARE, RUN_NO_ONEWALK = 256, 512, 1024, 2048


class Params(C.Structure):  # struct swk_params
    _fields_ = [
        ("B0", C.c_float), ("c", C.c_float), ("s", C.c_float),
        ("linear_phase_cycling", C.c_float), ("quadratic_phase_cycling", C.c_float),
        ("timestep_us", C.c_int32), ("TR_us", C.c_int32), ("n_dummy_scan", C.c_int32),
        ("n_spins", C.c_uint32), ("n_timepoints", C.c_uint32), ("n_substrate", C.c_uint32),
        ("seed", C.c_uint64), ("max_iterations", C.c_uint64),
        ("cross_fov", C.c_int32), ("record_trajectory", C.c_int32),
    ]


_TABLE_FIELDS = [
    ("step_sigma_m", "n_step_sigma"), ("T1_ms", "n_T1"), ("T2_ms", "n_T2"), ("pXY", "n_pXY"),
    ("RF_FA_deg", "n_RF_FA"), ("RF_PH_deg", "n_RF_PH"), ("RF_tp", "n_RF"), ("TE_tp", "n_TE"),
    ("dephasing_deg", "n_dephasing_deg"), ("dephasing_tp", "n_dephasing"),
    ("gradX_mTm", "n_gradX"), ("gradY_mTm", "n_gradY"), ("gradZ_mTm", "n_gradZ"), ("gradient_tp", "n_gradient"),
]


class Tables(C.Structure):  # struct swk_tables
    _fields_ = [f for p, n in _TABLE_FIELDS for f in ((p, C.c_void_p), (n, C.c_uint32))]


class Stats(C.Structure):  # struct swk_stats
    _fields_ = [
        ("steps", C.c_uint64), ("mask_gathers", C.c_uint64), ("field_gathers", C.c_uint64),
        ("rejects", C.c_uint64), ("lost", C.c_uint64), ("kernel_ms", C.c_float), ("device_ms", C.c_float), ("n_launches", C.c_uint32),
    ]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


SHAPE_CYLINDER, SHAPE_SPHERE, SHAPE_TWOPOOLS = 0, 1, 2


class PhantomSpec(C.Structure):  # struct swk_phantom_spec
    _fields_ = [
        ("shape", C.c_int32), ("fov_um", C.c_float), ("resolution", C.c_uint64), ("dchi", C.c_float), ("oxy_level", C.c_float),
        ("radius_um", C.c_float), ("volume_fraction", C.c_float), ("orientation_deg", C.c_float), ("seed", C.c_int32),
    ]


class PhantomStats(C.Structure):  # struct swk_phantom_stats
    _fields_ = [
        ("n_shapes", C.c_uint32), ("volume_fraction", C.c_float), ("place_ms", C.c_float), ("kernel_ms", C.c_float),
        ("n_launches", C.c_uint32), ("exact_columns", C.c_uint64),
    ]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/spinwalk_engine.h and include/spinwalk_phantom.h declare: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "swk_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "swk_destroy": (None, [_P]),
    "swk_last_error": (C.c_char_p, [_P]),
    "swk_version": (C.c_int, []),
    "swk_device_count": (C.c_int, []),
    "swk_device_info": (C.c_int, [C.c_char_p, C.c_size_t]),
    "swk_prepare": (C.c_int, [C.POINTER(Params), C.c_float, C.c_float, _P, C.c_uint32, _P]),
    "swk_set_phantom": (C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    "swk_set_sequence": (C.c_int, [_P, C.POINTER(Params), C.POINTER(Tables)]),
    "swk_set_spins": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32]),
    "swk_run_device": (C.c_int, [_P, _P, C.c_uint32, C.c_int, C.c_int, C.c_int, _P]),
    "swk_download": (C.c_int, [_P, _P, _P, _P]),
    "swk_get_sums": (C.c_int, [_P, _P]),
    "swk_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "swk_run": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.c_int, C.c_int, _P, _P, _P, _P, C.POINTER(Stats)]),
    "swk_set_host_rows": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "swk_alloc_pinned": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "swk_free_pinned": (None, [_P]),
    "swk_probe_gather": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "swk_stream": (_P, [_P]),
    "swk_device_sums": (_P, [_P]),
    "swk_device_bytes": (C.c_uint64, [_P]),
    "swk_debug_rng": (C.c_int, [_P, C.c_int, _P, C.c_uint32, _P]),
    # include/spinwalk_phantom.h
    "swk_phantom_shapes": (C.c_int, [C.POINTER(PhantomSpec), _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "swk_phantom_generate": (C.c_int, [C.c_int, C.POINTER(PhantomSpec), _P, _P, C.c_int, C.POINTER(PhantomStats)]),
    "swk_generate_phantom": (C.c_int, [_P, C.POINTER(PhantomSpec), C.POINTER(PhantomStats)]),
    "swk_get_phantom": (C.c_int, [_P, _P, _P]),
    "swk_phantom_mesh": (C.c_int, [C.c_int, C.c_float, C.c_uint64, _P, C.c_uint64, _P, C.c_uint64, _P, C.c_int, C.POINTER(PhantomStats)]),
    "swk_phantom_last_error": (C.c_char_p, []),
}

_lib = None


def load():
    """Load libspinwalk_b200.so.  Fails loudly when the CUDA extension has not been built —
    there is deliberately no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m spinwalk_b200.build` "
                "(spinwalk_b200 has no CPU or pure-Python fallback)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
