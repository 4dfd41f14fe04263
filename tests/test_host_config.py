"""Host config reader (host/sim_config.cpp) == the reference's sim::config_reader (src/sim/config_reader.cpp:25-324).

Goldens in tests/golden/config/*.json were produced by the reference's own reader compiled unmodified
(oracle/_ref/ref_config, see tests/golden/make_config_golden.py); where that binary exists the comparison is also made
live, including on the reference's shipped config/*.ini files."""
import ctypes as C
import json
import os
import subprocess
import sys

import pytest

import h5util

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_config_golden as mk  # noqa: E402

REF_BIN = os.path.join(h5util.ROOT, "oracle", "_ref", "ref_config")
CASES = sorted(f[:-4] for f in os.listdir(os.path.join(HERE, "golden", "config")) if f.endswith(".ini"))
PATH_KEYS = ("error", "seq_name", "output_dir")  # not exposed by the reference reader's getters / ours only


def ours(path, check_files=1):
    buf = C.create_string_buffer(1 << 20)
    h5util.lib().swkh_config_json(path.encode(), check_files, buf, len(buf))
    return json.loads(buf.value.decode())


def strip(d):
    return {k: v for k, v in d.items() if k not in PATH_KEYS}


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("cfg"))
    mk.stage(root)
    return root


@pytest.mark.parametrize("name", CASES)
def test_matches_reference_golden(staged, name):
    got = mk.normalise(ours(os.path.join(staged, "cfg", name + ".ini")), staged)
    want = json.load(open(os.path.join(HERE, "golden", "config", name + ".json")))
    assert got["ok"] == want["ok"]
    if want["ok"]:
        assert strip(got) == strip(want)
    else:
        assert got["error"]  # rejected with a message, like the reference's log line + `return false`


@pytest.mark.parametrize("name", CASES)
def test_matches_reference_live(staged, name):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/ref_config not built (needs /root/reference)")
    p = os.path.join(staged, "cfg", name + ".ini")
    ref = json.loads(subprocess.run([REF_BIN, p], capture_output=True, text=True).stdout.strip().splitlines()[-1])
    got = ours(p)
    assert got["ok"] == ref["ok"]
    if ref["ok"]:
        assert strip(got) == strip(ref)


def test_reference_shipped_configs(tmp_path):
    """config/*.ini of the reference tree parse to the same values (gradient.ini is rejected by both: it overrides T1/T2 for
    one substrate while inheriting a 2x2 P_XY)."""
    src = "/root/reference/config"
    if not os.path.exists(REF_BIN) or not os.path.isdir(src):
        pytest.skip("reference tree not present")
    import shutil

    cfg = tmp_path / "a" / "b"
    cfg.mkdir(parents=True)
    (tmp_path / "phantoms").mkdir()
    for i in range(2):
        (tmp_path / "phantoms" / f"phantom_{i}.h5").write_bytes(b"")
    n_ok = 0
    for f in sorted(os.listdir(src)):
        shutil.copy(os.path.join(src, f), cfg / f)
    for f in sorted(os.listdir(src)):
        ref = json.loads(subprocess.run([REF_BIN, str(cfg / f)], capture_output=True, text=True).stdout.strip().splitlines()[-1])
        got = ours(str(cfg / f))
        assert got["ok"] == ref["ok"], f
        if ref["ok"]:
            assert strip(got) == strip(ref), f
            n_ok += 1
    assert n_ok >= 5


def test_missing_files_and_output_dir(staged, tmp_path):
    """check(): every listed file must exist (config_reader.cpp:245-251); OUTPUT_DIR is created; output names are
    {OUTPUT_DIR}/{SEQ_NAME}_{phantom stem}.h5 (config_reader.cpp:266-271)."""
    p = tmp_path / "c.ini"
    p.write_text(open(os.path.join(staged, "cfg", "no_scales.ini")).read().replace("../ph/a.h5", "nowhere.h5"))
    r = ours(str(p))
    assert not r["ok"] and "does not exist" in r["error"]
    r = ours(os.path.join(staged, "cfg", "child.ini"))
    assert r["ok"] and os.path.isdir(os.path.join(staged, "cfg", "child_out"))
    assert r["output_files"] == [os.path.join(os.path.realpath(staged), "cfg", "child_out", "child   ; trailing text stays in the value_b.h5")]
    assert r["scales"] == [0.125] and r["n_substrate"] == 3 and r["scale_type"] == 1
    r = ours(os.path.join(staged, "cfg", "nonexistent.ini"))
    assert not r["ok"] and "does not exist" in r["error"]


# ---- randomised configs against the reference reader (hypothesis) ---------------------------------------------------
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

_num = st.one_of(st.integers(0, 200000).map(str), st.floats(0, 1e5, allow_nan=False).map(lambda v: f"{v:.4g}"),
                 st.sampled_from(["1e3", "2.5e4", "100e3", "7", "0", "0.0", "50", "1e-9", "-1", "-3.5", "abc", "", "12 ; c", " 9 "]))
_vec = st.lists(_num, min_size=0, max_size=4).map(" ".join)
_scan_keys = ["TR", "TE", "RF_FA", "RF_PH", "RF_T", "DEPHASING", "DEPHASING_T", "GRADIENT_X", "GRADIENT_Y", "GRADIENT_Z", "GRADIENT_T",
              "TIME_STEP", "DUMMY_SCAN", "LINEAR_PHASE_CYCLING", "QUADRATIC_PHASE_CYCLING"]
_sim_keys = ["B0", "SEED", "NUMBER_OF_SPINS", "CROSS_FOV", "RECORD_TRAJECTORY", "MAX_ITERATIONS", "WHAT_TO_SCALE", "SCALE[0]", "SCALE[1]", "SCALE[3]"]
_tis_keys = ["DIFFUSIVITY[0]", "DIFFUSIVITY[1]", "T1[0]", "T1[1]", "T2[0]", "T2[1]", "P_XY[0]", "P_XY[1]"]


@st.composite
def _child_ini(draw):
    lines = ["[GENERAL]", "PARENT_CONFIG = base.ini"]
    if draw(st.booleans()):
        lines.append("SEQ_NAME = " + draw(st.sampled_from(["x", "a b", "q;r", ""])))
    for sec, keys, strat in (("SCAN_PARAMETERS", _scan_keys, _vec), ("SIMULATION_PARAMETERS", _sim_keys, _num), ("TISSUE_PARAMETERS", _tis_keys, _vec)):
        chosen = draw(st.lists(st.sampled_from(keys), max_size=5, unique=True))
        if chosen:
            lines.append(f"[{sec}]" + draw(st.sampled_from(["", " ; note", "   "])))
            for k in chosen:
                lines.append(draw(st.sampled_from(["", "  ", "\t"])) + k + draw(st.sampled_from(["=", " = ", " =", "= "])) + draw(strat))
    return "\n".join(lines) + "\n"


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=_child_ini())
def test_random_children_match_reference(staged, text):
    """random child configs over base.ini: accepted or rejected alike, and when accepted, identical values.  (A malformed number
    makes the reference throw from std::stof / stoi — it would terminate; both report failure here.)"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/ref_config not built (needs /root/reference)")
    p = os.path.join(staged, "cfg", "_random_child.ini")
    with open(p, "w") as f:
        f.write(text)
    r = subprocess.run([REF_BIN, p], capture_output=True, text=True)
    ref = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"ok": False}  # e.g. SIGFPE on TIME_STEP = 0
    got = ours(p)
    assert got["ok"] == ref["ok"], (text, got.get("error"))
    if ref["ok"]:
        assert strip(got) == strip(ref), text
