"""The `spinwalk` binary itself (host/main.cpp) on the subcommands that need no GPU: `config` and `dwi` replay the invocations of
tests/generator_cases.py through the real command line and must write the reference's files byte for byte (goldens written by the
reference's own code, tests/golden/generators/out); plus the top-level conventions of src/spinwalk.cpp:45-90 (help, version, required
options, unknown options)."""
import os
import subprocess

import pytest

import generator_cases as gc
import h5util

BIN = os.path.join(h5util.ROOT, "bin", "spinwalk")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generators", "out")


@pytest.fixture(scope="module")
def cli():
    h5util.lib()  # runs `make -C host`, which also links bin/spinwalk when the engine library is there
    if not os.path.exists(BIN):
        pytest.skip("bin/spinwalk is not built (needs spinwalk_b200/libspinwalk_b200.so)")
    return BIN


def _run(cli, *args):
    return subprocess.run([cli, *map(str, args)], capture_output=True, text=True)


def test_config_and_dwi_through_the_command_line_equal_the_reference_goldens(cli, tmp_path):
    def config(seq, te, dt, phantoms, output):
        return _run(cli, "config", "-s", seq, "-p", *phantoms, "-e", te, "-t", dt, "-o", output).returncode == 0

    def dwi(cfg, b, v, d):
        return _run(cli, "dwi", "-b", *[repr(float(x)) for x in b], "-v", *v, "-d", *d, "-c", cfg).returncode == 0

    root = os.path.realpath(str(tmp_path))
    got = gc.run_all(root, config, dwi)
    assert sorted(got) == sorted(os.listdir(OUT))
    for name, text in got.items():
        with open(os.path.join(OUT, name), newline="") as f:
            assert text == f.read(), name


def test_top_level_conventions(cli, tmp_path):
    r = _run(cli)  # no subcommand: help, exit 0 (src/spinwalk.cpp:87-90)
    assert r.returncode == 0 and "sim" in (r.stdout + r.stderr) and "phantom" in (r.stdout + r.stderr)
    for args in (("--help",), ("-h",), ("sim", "--help"), ("phantom", "-h"), ("config", "--help"), ("dwi", "-h")):
        assert _run(cli, *args).returncode == 0, args
    r = _run(cli, "--version")
    assert r.returncode == 0 and "spinwalk" in r.stdout
    r = _run(cli, "--nonsense")
    assert r.returncode != 0 and "not expected" in r.stderr
    r = _run(cli, "sim")
    assert r.returncode != 0 and "--configs is required" in r.stderr
    r = _run(cli, "sim", "-c", str(tmp_path / "missing.ini"))
    assert r.returncode != 0 and "File does not exist" in r.stderr
    ini = tmp_path / "a.ini"
    ini.write_text("[GENERAL]\nSEQ_NAME = x\n")
    r = _run(cli, "sim", "-p", "-c", str(ini))  # the reference's CPU switch: refused, there is no CPU path
    assert r.returncode == 1 and "no CPU path" in r.stderr
    r = _run(cli, "config", "-s", "GRE", "-e", "1000", "-t", "10", "-o", str(tmp_path / "c.ini"))
    assert r.returncode != 0 and "required" in r.stderr
    r = _run(cli, "phantom", "-c", "-z", "10", "-o", str(tmp_path / "p.h5"))
    assert r.returncode != 0 and "--fov is required" in r.stderr
