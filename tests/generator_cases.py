"""The `spinwalk config` / `spinwalk dwi` invocations replayed on the reference (goldens) and on host/generators.cpp (tests)."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
IN = os.path.join(HERE, "golden", "generators", "in")

B_DEMO = [100.0 * k for k in range(1, 51)] + [0.0]  # demo/spinwalk_dwi.ipynb: -b 100 ... 5000, 0


def run_all(root, config, dwi):
    """config(seq, TE_us, dt_us, phantoms, output) -> bool and dwi(config_path, b_values, direction, (start, delta, DELTA)) -> bool run in `root`;
    returns {golden file name: text with root rewritten to $ROOT}."""
    results = {}

    def grab(name, path):
        with open(path, newline="") as f:
            results[name] = f.read().replace(root, "$ROOT")

    # config: the three sequences; the output directory does not exist yet; a relative phantom path is stored as given
    for seq, te, dt, phantoms in (("GRE", 20000, 50, ["./phantoms/r8_Y78.h5", "./phantoms/r8_Y85.h5"]), ("se", 30000, 25, ["/data/ph.h5"]), ("bSSFP", 5000, 50, ["p.h5"])):
        d = os.path.join(root, "cfg_" + seq.lower())
        assert config(seq, te, dt, phantoms, os.path.join(d, "sub", seq.lower() + ".ini"))
        grab(f"config_{seq.lower()}.ini", os.path.join(d, "sub", seq.lower() + ".ini"))
        grab(f"config_{seq.lower()}_default.ini", os.path.join(d, "sub", "default_config.ini"))
    assert not config("flash", 1000, 10, ["p.h5"], os.path.join(root, "bad", "x.ini"))  # "Invalid sequence name!"

    # dwi: (a) demo recipe on a generated GRE config, (b) child config inheriting TIME_STEP, no SCAN_PARAMETERS section of its own,
    # (c) in-place edits of existing entries with odd spacing, oblique direction, n_points not dividing evenly
    d = os.path.join(root, "dwi")
    os.makedirs(d)
    assert config("gre", 60000, 50, ["./phantoms/spheres.h5"], os.path.join(d, "dwi_demo.ini"))
    assert dwi(os.path.join(d, "dwi_demo.ini"), B_DEMO, (1.0, 0.0, 0.0), (15, 10, 20))
    grab("dwi_demo.ini", os.path.join(d, "dwi_demo.ini"))
    for f in ("dwi_parent.ini", "dwi_child.ini", "dwi_inplace.ini"):
        shutil.copy(os.path.join(IN, f), os.path.join(d, f))
    assert dwi(os.path.join(d, "dwi_child.ini"), [1000.0, 250.0, 4000.0], (0.267, 0.534, 0.801), (10, 3, 5))
    grab("dwi_child.ini", os.path.join(d, "dwi_child.ini"))
    assert dwi(os.path.join(d, "dwi_inplace.ini"), [700.0], (0.0, -2.0, 1.0), (2, 1, 7))
    grab("dwi_inplace.ini", os.path.join(d, "dwi_inplace.ini"))
    # refusals: Δ < δ, zero direction; the file must stay untouched
    shutil.copy(os.path.join(IN, "dwi_inplace.ini"), os.path.join(d, "untouched.ini"))
    assert not dwi(os.path.join(d, "untouched.ini"), [700.0], (0.0, 0.0, 1.0), (2, 5, 3))
    assert not dwi(os.path.join(d, "untouched.ini"), [700.0], (0.0, 0.0, 0.0), (2, 1, 7))
    grab("untouched.ini", os.path.join(d, "untouched.ini"))
    return results
