"""The CPU oracle equals the UNMODIFIED reference sim::sim (oracle/_ref, both RNG flavours) bit for bit on seeded RANDOM cases drawn over
the whole parameter space of the hot path (tests/random_cases.py) — the hand-written cases of tests/cases.py pin the paths somebody
thought of, these pin the combinations nobody did.  Needs /root/reference to have been compiled here (skipped on the GPU box)."""
import numpy as np
import pytest

import random_cases

N_CASES = 200


@pytest.mark.parametrize("flavour", [0, 1], ids=["mt19937", "minstd"])
def test_oracle_equals_reference_build_on_random_cases(oracle, flavour, capfd):
    if not (oracle.have_ref_cpu() if flavour == 0 else oracle.have_ref_cuda()):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    lost_cases = 0
    for seed in range(N_CASES):
        case, mask, fm, fov, xyz0 = random_cases.make(seed)
        o = oracle.run_oracle(case, fm, mask, xyz0, flavour=flavour, threads=2)
        r = oracle.run_ref(case, fm, mask, xyz0, flavour=flavour, threads=2)
        capfd.readouterr()  # the reference printf()s its lost-spin warnings
        assert np.array_equal(o["T"], r["T"]), seed
        assert np.array_equal(o["XYZ1"].view(np.uint32), r["XYZ1"].view(np.uint32)), seed
        assert np.array_equal(o["M1"].view(np.uint32), r["M1"].view(np.uint32)), seed
        lost_cases += int(o["stats"]["lost"] > 0)
    assert 0 < lost_cases < N_CASES  # the early-exit paths (kernels.cu:141-159) are exercised, but not by every case
