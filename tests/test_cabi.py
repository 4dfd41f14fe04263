"""CPU checks of the drop-in boundary: libspinwalk_b200.so loads, exports every function that
include/spinwalk_engine.h declares (and the Python declarations cover exactly those), fails loudly without a GPU,
and its parameters::prepare restatement (swk_prepare) agrees with the oracle's.  No compute is launched here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("spinwalk_engine.h", "spinwalk_phantom.h"))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swk_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = header_functions()
    for must in ("swk_create", "swk_destroy", "swk_set_phantom", "swk_set_sequence", "swk_set_spins", "swk_run_device",
                 "swk_download", "swk_get_sums", "swk_run", "swk_last_error"):
        assert must in names


def test_library_exports_every_header_symbol(engine_lib):
    from spinwalk_b200 import _lib

    names = header_functions()
    assert sorted(_lib.SYMBOLS) == names, "spinwalk_b200/_lib.py and include/spinwalk_engine.h disagree"
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert getattr(raw, n) is not None
    assert engine_lib.swk_version() == 1


def test_enum_values_match_the_header():
    """modes, scale types, output and run flags, error codes: the ctypes mirror holds the values a C program sees in the header."""
    import subprocess
    import tempfile

    from spinwalk_b200 import _lib

    names = ["SWK_OK", "SWK_ERR_INVALID", "SWK_ERR_CUDA", "SWK_ERR_MEMORY", "SWK_ERR_STATE", "SWK_ERR_SUBSTRATE", "SWK_SCALE_FOV", "SWK_SCALE_GRADIENT",
             "SWK_SCALE_PHASE_CYCLING", "SWK_MODE_COMPAT", "SWK_MODE_FAST", "SWK_OUT_M1", "SWK_OUT_XYZ1", "SWK_OUT_T", "SWK_OUT_ALL", "SWK_RUN_STATS", "SWK_RUN_NO_SORT",
             "SWK_RUN_NO_PACK", "SWK_RUN_NO_REBIN", "SWK_RUN_ZSLAB", "SWK_RUN_NO_ZSLAB", "SWK_RUN_NO_SHARE", "SWK_RUN_NO_ONEWALK"]
    prog = '#include <stdio.h>\n#include "spinwalk_engine.h"\nint main(){' + "".join(f'printf("%d\\n", (int){n});' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "e.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "e.c"), "-o", os.path.join(d, "e")], check=True)
        out = [int(v) for v in subprocess.run([os.path.join(d, "e")], capture_output=True, text=True, check=True).stdout.split()]
    mirror = [getattr(_lib, n[4:] if not n.startswith("SWK_OK") and not n.startswith("SWK_ERR") else n) for n in names]
    assert out == mirror, dict(zip(names, zip(out, mirror)))


def test_struct_sizes_match_the_header(engine_lib):
    """ctypes mirrors of swk_params / swk_tables / swk_stats have the C layout (checked via a tiny C program)."""
    import subprocess
    import tempfile

    from spinwalk_b200 import _lib

    prog = ('#include <stdio.h>\n#include "spinwalk_phantom.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(swk_params), sizeof(swk_tables), '
            'sizeof(swk_stats), sizeof(swk_phantom_spec), sizeof(swk_phantom_stats));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [C.sizeof(_lib.Params), C.sizeof(_lib.Tables), C.sizeof(_lib.Stats), C.sizeof(_lib.PhantomSpec), C.sizeof(_lib.PhantomStats)]


def test_no_cpu_fallback(engine_lib):
    """Without a CUDA device the engine refuses to exist (it must never silently compute on the CPU)."""
    import spinwalk_b200 as sw

    if engine_lib.swk_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(sw.EngineError, match="no CUDA device"):
        sw.Engine(0)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under spinwalk_b200/, host/ or include/ may reference it."""
    for pkg in ("spinwalk_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt and "import oracle" not in txt, f


def test_swk_prepare_matches_oracle(engine_lib, oracle):
    import numpy as np

    from spinwalk_b200 import _lib

    case = oracle.Case(fov=(1e-4,) * 3, phantom_size=(8, 8, 8), n_spins=10, TR_us=10000, timestep_us=50, n_dummy_scan=-1,
                       T1_ms=[1000.0, 2200.0], RF_FA_deg=[16.0], diffusivity=[1e-9, 2.5e-9])
    p = _lib.Params()
    p.timestep_us, p.TR_us, p.n_dummy_scan = case.timestep_us, case.TR_us, case.n_dummy_scan
    D = np.asarray(case.diffusivity, np.float64)
    sig = np.zeros(2)
    assert engine_lib.swk_prepare(C.byref(p), 16.0, 1000.0, D.ctypes.data, 2, sig.ctypes.data) == 0
    lo = oracle._lib(oracle.LIB_ORACLE)
    cc, keep = case.to_c()
    lo.swo_n_dummy_scan.restype = C.c_int32
    lo.swo_n_timepoints.restype = C.c_uint32
    lo.swo_step_sigma.restype = C.c_double
    lo.swo_step_sigma.argtypes = [C.c_double, C.c_int32]
    assert p.n_dummy_scan == lo.swo_n_dummy_scan(C.byref(cc)) == 500  # 5*T1/TR (simulation_parameters.cuh:239-242)
    assert p.n_timepoints == lo.swo_n_timepoints(C.byref(cc)) == 200
    assert sig[0] == lo.swo_step_sigma(1e-9, 50) and sig[1] == lo.swo_step_sigma(2.5e-9, 50)
    assert p.c == np.float32(np.cos(np.float32(16.0 * 0.0174532925199433))) and p.s == np.float32(np.sin(np.float32(16.0 * 0.0174532925199433)))
