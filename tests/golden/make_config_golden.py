"""Regenerates tests/golden/config/*.json: what the REFERENCE's own sim::config_reader (oracle/_ref/ref_config, built by
oracle/Makefile from /root/reference/src/sim/config_reader.cpp) makes of tests/golden/config/*.ini.  Paths are stored
relative to the temporary root the configs were parsed in ($ROOT)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def stage(root):
    os.makedirs(os.path.join(root, "cfg"))
    os.makedirs(os.path.join(root, "ph"))
    for f in os.listdir(os.path.join(HERE, "config")):
        if f.endswith(".ini"):
            shutil.copy(os.path.join(HERE, "config", f), os.path.join(root, "cfg", f))
    for f in ("a.h5", "b.h5"):
        open(os.path.join(root, "ph", f), "w").close()


def normalise(obj, root):
    s = json.dumps(obj).replace(os.path.realpath(root), "$ROOT").replace(root, "$ROOT")
    return json.loads(s)


if __name__ == "__main__":
    ref = os.path.join(REPO, "oracle", "_ref", "ref_config")
    if not os.path.exists(ref):
        sys.exit("oracle/_ref/ref_config is missing: make -C oracle ref")
    root = tempfile.mkdtemp()
    stage(root)
    for f in sorted(os.listdir(os.path.join(root, "cfg"))):
        out = subprocess.run([ref, os.path.join(root, "cfg", f)], capture_output=True, text=True).stdout.strip().splitlines()[-1]
        with open(os.path.join(HERE, "config", f.replace(".ini", ".json")), "w") as g:
            json.dump(normalise(json.loads(out), root), g, indent=1, sort_keys=True)
            g.write("\n")
        print(f, json.loads(out).get("ok"))
