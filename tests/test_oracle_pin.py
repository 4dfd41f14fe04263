"""Pins the CPU oracle (oracle/sim_oracle.c):
  1. against the UNMODIFIED reference sim::sim compiled here from /root/reference into oracle/_ref (both RNG
     flavours), bit for bit, on every case of tests/cases.py  — runs wherever oracle/_ref exists;
  2. against the committed golden vectors tests/golden/*.npz, which were produced by that same reference build
     (tests/golden/make_golden.py) — runs everywhere, including boxes without /root/reference."""
import glob
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("flavour", [0, 1], ids=["mt19937", "minstd"])
@pytest.mark.parametrize("name", list(cases.ALL))
def test_oracle_equals_reference_build(oracle, name, flavour, capfd):
    if not (oracle.have_ref_cpu() if flavour == 0 else oracle.have_ref_cuda()):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    o = oracle.run_oracle(case, fm, mask, xyz0, flavour=flavour)
    r = oracle.run_ref(case, fm, mask, xyz0, flavour=flavour)
    capfd.readouterr()  # the reference printf()s its lost-spin warnings
    assert np.array_equal(o["T"], r["T"])
    assert np.array_equal(o["XYZ1"].view(np.uint32), r["XYZ1"].view(np.uint32))
    assert np.array_equal(o["M1"].view(np.uint32), r["M1"].view(np.uint32))


def test_default_positions_equal_reference(oracle):
    if not oracle.have_ref_cpu():
        pytest.skip("oracle/_ref not built")
    fov = np.array([600e-6, 300e-6, 123e-6], np.float32)
    assert np.array_equal(oracle.init_positions(10, fov, 5000, "oracle"), oracle.init_positions(10, fov, 5000, "ref"))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_equals_golden(oracle, path):
    g = np.load(path)
    name, flavour = str(g["case"]), int(g["flavour"])
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    assert np.array_equal(xyz0, g["xyz0"]), "seeded inputs changed: regenerate tests/golden with make_golden.py"
    o = oracle.run_oracle(case, fm, mask, xyz0, flavour=flavour)
    assert np.array_equal(o["T"], g["T"])
    assert np.array_equal(o["XYZ1"].view(np.uint32), g["XYZ1"].view(np.uint32))
    assert np.array_equal(o["M1"].view(np.uint32), g["M1"].view(np.uint32))


def test_golden_present():
    assert len(glob.glob(os.path.join(GOLDEN, "*.npz"))) >= 2 * len(cases.ALL) - 2
