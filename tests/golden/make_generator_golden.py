"""Golden outputs of the reference's `spinwalk config` and `spinwalk dwi` (UNMODIFIED src/config/*.cpp, src/dwi/*.cpp + the vendored
mINI writer, compiled by oracle/Makefile into oracle/_ref/libswref_gen.so).  Run in the build container:
    python tests/golden/make_generator_golden.py
Inputs: tests/golden/generators/in/*.ini.  Outputs: tests/golden/generators/out/*, absolute paths rewritten to $ROOT.
tests/generator_cases.py lists the invocations; tests/test_host_generators.py replays them on host/generators.cpp."""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from generator_cases import run_all  # noqa: E402
from oracle import pyphantom as pp  # noqa: E402


def main():
    out = os.path.join(HERE, "generators", "out")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    root = os.path.realpath(tempfile.mkdtemp())
    files = run_all(root, lambda seq, te, dt, phantoms, output: pp.reference_config(seq, te, dt, phantoms, output),
                    lambda cfg, b, v, d: pp.reference_dwi(cfg, b, v, *d))
    for name, text in files.items():
        with open(os.path.join(out, name), "w", newline="") as f:
            f.write(text)
        print(name, len(text))


if __name__ == "__main__":
    main()
