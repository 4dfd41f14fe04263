"""Host-side generators (host/generators.cpp, host/ini_edit.cpp) == the reference's `spinwalk config` / `spinwalk dwi`
(src/config/*.cpp, src/dwi/*.cpp + the vendored mINI writer), byte for byte.

Goldens in tests/golden/generators/out/ were written by the reference's own code compiled unmodified
(oracle/_ref/libswref_gen.so; tests/golden/make_generator_golden.py).  Where that library exists the INI writer is also compared
live with mINI on randomised files and edits."""
import ctypes as C
import os
import shutil

import pytest

import generator_cases as gc
import h5util

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "golden", "generators", "out")
REF_LIB = os.path.join(h5util.ROOT, "oracle", "_ref", "libswref_gen.so")


def our_config(seq, te, dt, phantoms, output):
    arr = (C.c_char_p * len(phantoms))(*[p.encode() for p in phantoms])
    buf = C.create_string_buffer(4096)
    return h5util.lib().swkh_config(seq.encode(), te, dt, arr, len(phantoms), output.encode(), buf, len(buf)) == 0


def our_dwi(cfg, b, v, d):
    bb = (C.c_double * len(b))(*b)
    vv = (C.c_float * 3)(*v)
    buf = C.create_string_buffer(4096)
    return h5util.lib().swkh_dwi(bb, len(b), vv, d[0], d[1], d[2], cfg.encode(), buf, len(buf)) == 0


def test_config_and_dwi_outputs_equal_the_reference_goldens(tmp_path):
    root = os.path.realpath(str(tmp_path))
    got = gc.run_all(root, our_config, our_dwi)
    assert sorted(got) == sorted(os.listdir(OUT))
    for name, text in got.items():
        with open(os.path.join(OUT, name), newline="") as f:
            want = f.read()
        assert text == want, name


def test_dwi_demo_recipe_numbers(tmp_path):
    """Known answer of the demo notebook (demo/spinwalk_dwi.ipynb): b = 100 s/mm^2, δ = 10 ms, Δ = 20 ms => G = sqrt(b 1e6 / (γ² δ² (Δ-δ/3)))."""
    import math

    cfg = str(tmp_path / "d.ini")
    assert our_config("gre", 60000, 50, ["p.h5"], cfg)
    assert our_dwi(cfg, gc.B_DEMO, (1.0, 0.0, 0.0), (15, 10, 20))
    kv = {}
    for line in open(cfg):
        if "=" in line:
            k, v = line.split("=", 1)
            kv[k.strip()] = v.strip()
    gx = [float(x) for x in kv["GRADIENT_X"].split()]
    gt = [int(x) for x in kv["GRADIENT_T"].split()]
    assert len(gx) == len(gt) == 404 and gt == sorted(gt) and len(set(gt)) == 404
    G = math.sqrt(100e6 / (267515315.0 ** 2 * 0.01 ** 2 * (0.02 - 0.01 / 3))) * 1000
    assert abs(max(gx) - G) < 1e-6 and gx[0] == 0 and gx[201] == 0 and gx[202] == 0
    assert kv["RF_T"] == "0 30000" and kv["WHAT_TO_SCALE"] == "1"
    assert float(kv["SCALE[49]"]) == pytest.approx(math.sqrt(50.0), abs=1e-6) and float(kv["SCALE[50]"]) == 0.0


def test_dwi_error_messages(tmp_path):
    cfg = str(tmp_path / "x.ini")
    open(cfg, "w").write("[GENERAL]\nSEQ_NAME = x\n")
    buf = C.create_string_buffer(4096)
    bb, vv = (C.c_double * 1)(100.0), (C.c_float * 3)(1, 0, 0)
    assert h5util.lib().swkh_dwi(bb, 1, vv, 1, 1, 2, cfg.encode(), buf, len(buf)) == 1
    assert b"TIME_STEP is not set" in buf.value  # pgse.cpp:61-65
    assert h5util.lib().swkh_dwi(bb, 1, vv, 1, 1, 2, str(tmp_path / "missing.ini").encode(), buf, len(buf)) == 1
    assert b"does not exist" in buf.value  # pgse.cpp:27-28


# ---- the INI writer against mINI itself, on randomised documents ----

SECTIONS = ["A", "SCAN_PARAMETERS", "b c", "Z[0]"]
KEYS = ["K", "key two", "E\\=Q", "X[1]", "k"]
VALUES = ["", "1", " 2 3 ", "v ; not a comment", "a=b", "0 0.5  1e-3"]


def random_ini(rnd):
    lines = []
    for _ in range(rnd.randint(0, 14)):
        r = rnd.random()
        if r < 0.22:
            lines.append(rnd.choice(["[%s]", " [%s] ", "[ %s ] ; tail", "[%s"]) % rnd.choice(SECTIONS))
        elif r < 0.62:
            lines.append(rnd.choice(["%s = %s", "%s=%s", "  %s =%s", "%s= %s", "%s\t=\t%s  "]) % (rnd.choice(KEYS), rnd.choice(VALUES)))
        elif r < 0.75:
            lines.append("")
        elif r < 0.85:
            lines.append(rnd.choice(["; comment", "  ; indented", ";"]))
        elif r < 0.93:
            lines.append(rnd.choice(["junk", "   ", "# not a comment"]))
        else:
            lines.append("\r")
    text = "\n".join(lines)
    if rnd.random() < 0.3:
        text += "\n"
    if rnd.random() < 0.1:
        text = "﻿" + text
    return text


def random_edits(rnd):
    ops = []
    for _ in range(rnd.randint(0, 6)):
        op = rnd.choice([0, 0, 0, 1, 2, 3])
        ops.append((op, rnd.choice(SECTIONS + ["NEW", " padded "]), rnd.choice(KEYS + ["fresh", "a=b"]), rnd.choice(VALUES)))
    return ops


def apply(fn, path, ops, pretty):
    n = len(ops)
    arr = lambda i: (C.c_char_p * n)(*[o[i].encode() for o in ops])  # noqa: E731
    return fn(path.encode(), n, (C.c_int * n)(*[o[0] for o in ops]), arr(1), arr(2), arr(3), pretty)


def test_ini_writer_equals_mini_on_random_documents(tmp_path):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libswref_gen.so not built (no reference tree here)")
    import random

    ref = C.CDLL(REF_LIB)
    ours = h5util.lib()
    rnd = random.Random(20240517)
    for trial in range(1500):
        text, ops, pretty = random_ini(rnd), random_edits(rnd), rnd.randint(0, 1)
        pa, pb = str(tmp_path / "a.ini"), str(tmp_path / "b.ini")
        if rnd.random() < 0.05:  # a file that does not exist yet is created
            for p in (pa, pb):
                if os.path.exists(p):
                    os.remove(p)
        else:
            for p in (pa, pb):
                with open(p, "w", encoding="utf-8", newline="") as f:
                    f.write(text)
        ra, rb = apply(ref.swref_ini_edit, pa, ops, pretty), apply(ours.swkh_ini_edit, pb, ops, pretty)
        assert ra == rb
        with open(pa, "rb") as fa, open(pb, "rb") as fb:
            a, b = fa.read(), fb.read()
        assert a == b, f"trial {trial}: pretty={pretty}\nINPUT:\n{text!r}\nEDITS: {ops}\nmINI:\n{a!r}\nours:\n{b!r}"


def test_reference_config_creation(tmp_path):
    """The reference's own test (tests/test_config.cpp:36-62): generate_gre(TE = 12345, timestep = 25, two phantoms) writes gre.ini and
    default_config.ini next to it; TE and TIME_STEP read back."""
    out = str(tmp_path / "spinwalk_test" / "gre.ini")
    assert our_config("gre", 12345, 25, ["phantom1.h5", "phantom2.h5"], out)
    parent = str(tmp_path / "spinwalk_test" / "default_config.ini")
    assert os.path.exists(out) and os.path.exists(parent)
    kv = dict(line.split(" = ", 1) for line in open(out).read().splitlines() if " = " in line)
    assert int(kv["TE"]) == 12345 and int(kv["TIME_STEP"]) == 25 and kv["PARENT_CONFIG"] == parent
    assert kv["PHANTOM[0]"] == "phantom1.h5" and kv["PHANTOM[1]"] == "phantom2.h5"
