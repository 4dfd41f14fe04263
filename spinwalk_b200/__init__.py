"""spinwalk_b200 — B200-native engine for the hot path of SpinWalk's `sim` (per-spin Monte-Carlo time loop).

The product is spinwalk_b200/libspinwalk_b200.so (hand-written sm_100a CUDA behind the C-ABI of
include/spinwalk_engine.h).  This package is the thin host mirror used by tests, bench.py and Python
callers; importing it does not need a GPU, creating an Engine does.  There is no CPU fallback.
"""
from .engine import (MODE_COMPAT, MODE_FAST, OUT_ALL, OUT_M1, OUT_T, OUT_XYZ1, RUN_NO_PACK, RUN_NO_REBIN, RUN_NO_SORT, RUN_STATS, RUN_ZSLAB, RUN_NO_ZSLAB, RUN_NO_SHARE, RUN_NO_ONEWALK, SCALE_FOV, SCALE_GRADIENT,  # noqa: F401
                     SCALE_PHASE_CYCLING, Engine, EngineError, SimConfig, simulate)

__version__ = "0.1.0"
