"""The drop-in host: `bin/spinwalk sim -c x.ini` (host/main.cpp -> sim_driver.cpp -> C-ABI) on real files.

A phantom HDF5 file (written by host/h5lite.cpp) + an INI in the reference's syntax go in; the output HDF5 file must hold
the reference's datasets (monte_carlo.cu:168-197) with, in --compat mode, exactly the arrays the reference's own cu_sim
kernel produces for the same inputs (T and XYZ bitwise, M to 2e-6) — positions included, because the host seeds
std::mt19937 the way monte_carlo.cu:142-151 does."""
import os
import subprocess

import numpy as np
import pytest

import cases
import h5util

pytestmark = pytest.mark.gpu
BIN = os.path.join(h5util.ROOT, "bin", "spinwalk")

INI = """[GENERAL]
SEQ_NAME = se_test
[FILES]
OUTPUT_DIR = ./out
PHANTOM[0] = ./phantom.h5
[TISSUE_PARAMETERS]
DIFFUSIVITY[0] = 1.0e-9
DIFFUSIVITY[1] = 1.0e-9
P_XY[0] = 1.0 0.0
P_XY[1] = 0.0 1.0
T1[0] = 2200
T1[1] = 2200
T2[0] = 41
T2[1] = 41
[SCAN_PARAMETERS]
TR = 40000
TE = 20000
RF_FA = 90.0 180.0
RF_PH = 0.0 90
RF_T = 0 10000
TIME_STEP = 50
DUMMY_SCAN = 0
[SIMULATION_PARAMETERS]
B0 = 9.4
SEED = 10
NUMBER_OF_SPINS = {S}
CROSS_FOV = 0
RECORD_TRAJECTORY = 0
MAX_ITERATIONS = 1e4
WHAT_TO_SCALE = 0
SCALE[0] = 0.1
SCALE[1] = 1.0
SCALE[2] = 8.0
"""


@pytest.fixture(scope="module")
def cli():
    subprocess.run(["make", "-s", "-C", os.path.join(h5util.ROOT, "host")], check=True)
    assert os.path.exists(BIN), "bin/spinwalk was not built (needs spinwalk_b200/libspinwalk_b200.so)"
    return BIN


def _stage(tmp_path, S, int8_mask=False):
    case, mask, fm, fov, xyz0 = cases.se(n_spins=S)
    h5util.write(str(tmp_path / "phantom.h5"), {"fieldmap": fm, "mask": mask.astype(np.int8) if int8_mask else mask, "fov": np.asarray(fov, np.float32),
                                                "bvf": np.array([8.0], np.float32)})
    (tmp_path / "se.ini").write_text(INI.format(S=S))
    return case, mask, fm, fov, xyz0


def test_cli_compat_equals_reference_kernel(cli, oracle, tmp_path):
    if not oracle.have_ref_cuda():
        pytest.skip("oracle/_ref/libswref_cuda.so not present")
    S = 2048
    case, mask, fm, fov, xyz0 = _stage(tmp_path, S, int8_mask=True)  # int8 mask as MATLAB / h5py users write it (README.md:145-163)
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "--compat", "--sums"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Simulation completed successfully" in r.stdout
    out = str(tmp_path / "out" / "se_test_phantom.h5")
    assert sorted(h5util.names(out)) == ["M", "T", "TE", "XYZ", "scales", "sums", "swk_mode", "swk_seed"]
    assert h5util.read(out, "swk_mode").ravel().tolist() == [0] and h5util.read(out, "swk_seed").ravel().tolist() == [case.seed]  # which arithmetic wrote the file
    # the INI route rounds DIFFUSIVITY through float (std::stof, config_reader.cpp:164), as the reference does
    case.diffusivity = [float(np.float32(1.0e-9))] * 2
    ref = oracle.run_ref_cuda(case, fm, mask, xyz0)  # xyz0 = the oracle's restatement of the mt19937 start positions
    M, X, T = h5util.read(out, "M"), h5util.read(out, "XYZ"), h5util.read(out, "T")
    assert M.shape == (3, S, 1, 3) and X.shape == (3, S, 1, 3) and T.shape == (3, S, 1, 1) and T.dtype == np.uint8
    assert np.array_equal(T[..., 0], ref["T"])
    assert np.array_equal(X.view(np.uint32), ref["XYZ1"].view(np.uint32))
    assert np.abs(M - ref["M1"]).max() <= 2e-6
    assert np.array_equal(h5util.read(out, "scales").ravel(), np.array([0.1, 1.0, 8.0], np.float32))
    assert np.allclose(h5util.read(out, "TE").ravel(), [0.02])
    sums = h5util.read(out, "sums")
    assert sums.shape == (3, 1, 2, 4) and sums[..., 3].sum() == 3 * S


def test_cli_two_engines_equal_one(cli, tmp_path):
    """-d 0,0: two engines (threads) share the host arrays through swk_set_host_rows; the spins are keyed by their global id,
    so the output file equals the single-engine one bit for bit (fast mode)."""
    S = 3001
    _stage(tmp_path, S)
    outs = []
    for dev in ("0", "0,0"):
        r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "-d", dev, "-q"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = str(tmp_path / "out" / "se_test_phantom.h5")
        outs.append({k: h5util.read(out, k) for k in ("M", "XYZ", "T")})
    for k in ("M", "XYZ", "T"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert (outs[0]["M"] != 0).any()


def test_cli_sums_only_and_device_positions(cli, tmp_path):
    """--sums-only: no per-spin arrays anywhere (what lets BASELINE's 1e9-spin configuration run through the command line); the sums equal
    those of the full run (fixed-point accumulation: exactly), on one or several engines, more engines than spins included.
    --device-positions: start positions drawn on the GPU; every spin is counted at the echo."""
    S = 5000
    _stage(tmp_path, S)
    ini, out = str(tmp_path / "se.ini"), str(tmp_path / "out" / "se_test_phantom.h5")
    r = subprocess.run([cli, "sim", "-c", ini, "--sums", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    full = h5util.read(out, "sums")
    M, T = h5util.read(out, "M"), h5util.read(out, "T")
    for sub in range(2):  # the sums are the per-spin outputs added up per substrate
        w = T[..., 0] == sub
        assert np.array_equal(full[:, :, sub, 3], w.sum(axis=1).astype(np.float64))
        assert np.allclose(full[:, :, sub, :3], (M.astype(np.float64) * w[..., None]).sum(axis=1), rtol=0, atol=S * 2.0 ** -22)
    for dev in ("0", "0,0,0"):
        r = subprocess.run([cli, "sim", "-c", ini, "--sums-only", "-d", dev, "-q"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert sorted(h5util.names(out)) == ["TE", "scales", "sums", "swk_mode", "swk_seed"]
        assert np.array_equal(h5util.read(out, "sums"), full), dev
        assert h5util.read(out, "swk_mode").ravel().tolist() == [1]
    r = subprocess.run([cli, "sim", "-c", ini, "--sums-only", "--device-positions", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dp = h5util.read(out, "sums")
    assert (dp[..., 3].sum(axis=2) == S).all() and not np.array_equal(dp, full)
    tot_a, tot_b = full.sum(axis=2)[:, 0], dp.sum(axis=2)[:, 0]
    assert np.abs(np.hypot(tot_a[:, 0], tot_a[:, 1]) - np.hypot(tot_b[:, 0], tot_b[:, 1])).max() / S < 0.03  # same ensemble, other start positions
    # more devices than spins: no empty shard
    (tmp_path / "two.ini").write_text(INI.format(S=2))
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "two.ini"), "--sums-only", "-d", "0,0,0", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (h5util.read(str(tmp_path / "out" / "se_test_phantom.h5"), "sums")[..., 3].sum(axis=2) == 2).all()


def test_output_file_opens_with_libhdf5(cli, tmp_path):
    """the h5lite-written output through a real libhdf5 (h5py), when this image has one — the side of the codec tests/h5spec.py stands in for"""
    h5py = pytest.importorskip("h5py")
    S = 300
    _stage(tmp_path, S)
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "--sums", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "out" / "se_test_phantom.h5")
    with h5py.File(out, "r") as f:
        assert f["M"].shape == (3, S, 1, 3) and f["XYZ"].shape == (3, S, 1, 3) and f["T"].shape == (3, S, 1, 1) and f["T"].dtype == np.uint8
        for k in ("M", "XYZ", "T", "scales", "TE", "sums"):
            assert np.array_equal(f[k][...], h5util.read(out, k)), k


def test_cli_error_conventions(cli, tmp_path):
    _stage(tmp_path, 64)
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "-p"], capture_output=True, text=True)
    assert r.returncode == 0 and "no CPU path" in r.stderr and "warning" in r.stderr  # -p: accepted, warned about, run on the GPU
    for bad in ("0,", "a", "0,,1", "-1"):
        r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "-d", bad], capture_output=True, text=True)
        assert r.returncode == 1 and "not a device id" in r.stderr, bad
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "se.ini"), "-d", "999"], capture_output=True, text=True)
    assert r.returncode == 1 and "not available" in r.stderr
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "missing.ini")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr
    (tmp_path / "one.ini").write_text(INI.format(S=64).replace("DIFFUSIVITY[1] = 1.0e-9\n", "").replace("T1[1] = 2200\n", "").replace("T2[1] = 41\n", "")
                                      .replace("P_XY[0] = 1.0 0.0\nP_XY[1] = 0.0 1.0", "P_XY[0] = 1.0"))
    r = subprocess.run([cli, "sim", "-c", str(tmp_path / "one.ini")], capture_output=True, text=True)
    assert r.returncode == 1 and "Simulation failed" in r.stderr and "substrate" in r.stderr  # monte_carlo.cu:113-118


# ---- the generator subcommands: phantom (GPU voxel fill), config, dwi — and the whole reference workflow through the CLI ----

def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cli_phantom_files_hold_the_reference_generator_arrays(cli, tmp_path):
    """`spinwalk phantom -c / -s / -t`: the phantom file carries the datasets of phantom_base::save (phantom_base.cpp:69-103) and the
    mask / field map are, byte for byte, what the reference's generator makes (goldens: SHA-256 of its arrays)."""
    from phantom_cases import CASES

    gold_dir = os.path.join(h5util.ROOT, "tests", "golden", "phantom")
    flags = {0: "-c", 1: "-s", 2: "-t"}
    for name in ("cyl_bold", "cyl_mask_only", "sph_fixed", "twopools_odd"):
        kw = CASES[name]
        out = str(tmp_path / "made" / (name + ".h5"))  # the directory does not exist yet (phantom_base.cpp:72-80)
        cmd = [cli, "phantom", flags[kw["shape"]], "-f", str(kw["fov_um"]), "-z", str(kw["resolution"]), "-o", out]
        if kw["shape"] != 2:
            cmd += ["-r", str(kw["radius_um"]), "-v", str(kw["volume_fraction"]), "-y", str(kw["Y"]), "-e", str(kw["seed"]), "-n", str(kw.get("orientation_deg", 90.0))]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Done." in r.stdout
        gold = np.load(os.path.join(gold_dir, name + ".npz"))
        has_field = "fieldmap_sha256" in gold
        assert sorted(h5util.names(out)) == sorted(["mask", "fov", "bvf"] + (["fieldmap"] if has_field else []))
        n = kw["resolution"]
        shape, dt, _ = h5util.info(out, "mask")
        assert shape == (n, n, n) and np.dtype(dt) == np.uint8
        assert _sha(h5util.read(out, "mask")) == str(gold["mask_sha256"])
        if has_field:
            shape, dt, _ = h5util.info(out, "fieldmap")
            assert shape == (n, n, n) and np.dtype(dt) == np.float32
            assert _sha(h5util.read(out, "fieldmap")) == str(gold["fieldmap_sha256"])
        assert np.array_equal(h5util.read(out, "fov"), np.full(3, np.float32(kw["fov_um"]) * np.float32(1e-6), np.float32))
        assert h5util.read(out, "bvf").ravel()[0] == gold["bvf"]


def test_cli_phantom_refusals(cli, tmp_path):
    out = str(tmp_path / "x.h5")
    r = subprocess.run([cli, "phantom", "-c", "-f", "10", "-z", "8", "-r", "6", "-e", "1", "-o", out], capture_output=True, text=True)
    assert r.returncode == 1 and "Phantom generation failed" in r.stderr and "too large" in r.stderr  # phantom_cylinder.cpp:87-91
    r = subprocess.run([cli, "phantom", "-c", "-z", "8", "-o", out], capture_output=True, text=True)
    assert r.returncode == 1 and "--fov is required" in r.stderr
    r = subprocess.run([cli, "phantom", "-p", "-i", str(tmp_path / "missing.ply"), "-f", "10", "-z", "8", "-o", out], capture_output=True, text=True)
    assert r.returncode == 1 and "File does not exist" in r.stderr
    assert not os.path.exists(out)


def test_cli_whole_workflow_config_phantom_dwi_sim(cli, tmp_path):
    """The demo notebooks' chain, every step through this CLI: phantom -> config -> dwi -> sim.  Free diffusion (P_XY = 1 everywhere,
    no relaxation) must give the Stejskal-Tanner answer exp(-b D) per b-value (demo/spinwalk_dwi.ipynb)."""
    ph = str(tmp_path / "phantoms" / "spheres.h5")
    # FoV 1 mm: with CROSS_FOV = 1 a spin that wraps around the FoV jumps by one FoV in the gradient's frame and is lost to the signal
    # (~1 % of the spins at 9 um rms displacement; the notebook's own fit of the reference's output has a = 0.9945 for the same reason)
    r = subprocess.run([cli, "phantom", "-s", "-r", "-60", "-v", "5", "-f", "1000", "-z", "64", "-y", "-1", "-e", "9", "-o", ph, "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cfg = str(tmp_path / "cfg" / "pgse.ini")
    r = subprocess.run([cli, "config", "-s", "GRE", "-p", ph, "-e", "40000", "-t", "50", "-o", cfg], capture_output=True, text=True)
    assert r.returncode == 0 and "Configuration file is generated in" in r.stdout, r.stderr
    b = [200, 600, 1000, 1500]
    r = subprocess.run([cli, "dwi", "-b", *map(str, b), "-v", "1", "0", "0", "-d", "5", "10", "20", "-c", cfg], capture_output=True, text=True)
    assert r.returncode == 0 and "generated successfully" in r.stdout, r.stderr
    # the notebook then edits the tissue parameters by hand: free diffusion, no relaxation, a fixed seed, more spins
    default = str(tmp_path / "cfg" / "default_config.ini")
    txt = open(default).read()
    for old, new in (("P_XY[0] = 1.0 0.0", "P_XY[0] = 1.0 1.0"), ("P_XY[1] = 0.0 1.0", "P_XY[1] = 1.0 1.0"), ("T1[0] = 2200", "T1[0] = -1"), ("T1[1] = 2200", "T1[1] = -1"),
                     ("T2[0] = 41", "T2[0] = -1"), ("T2[1] = 41", "T2[1] = -1"), ("SEED = 0", "SEED = 7"), ("NUMBER_OF_SPINS = 1e5", "NUMBER_OF_SPINS = 2e5"),
                     ("CROSS_FOV = 0", "CROSS_FOV = 1")):
        assert old in txt
        txt = txt.replace(old, new)
    open(default, "w").write(txt)
    r = subprocess.run([cli, "sim", "-c", cfg, "-q"], capture_output=True, text=True, cwd=str(tmp_path / "cfg"))
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "cfg" / "outputs" / "gre_spheres.h5")  # {OUTPUT_DIR}/{SEQ_NAME}_{phantom stem}.h5 (config_reader.cpp:266-271)
    M = h5util.read(out, "M")
    assert M.shape == (4, 200000, 1, 3)
    sig = np.hypot(M[..., 0].astype(np.float64).mean(axis=1), M[..., 1].astype(np.float64).mean(axis=1)).ravel()
    slope, intercept = np.polyfit(np.asarray(b, np.float64) * 1e6, np.log(sig), 1)  # b in s/mm^2 -> s/m^2; simulated D = 1e-9 m^2/s
    # same windows as tests/test_engine_gpu.py::test_fast_pgse_free_diffusion_known_answer (Monte-Carlo error of |S| ~ 2e-3 at 2e5 spins)
    assert abs(-slope / 1e-9 - 1.0) < 0.04, f"fitted D = {-slope:.3e}, a = {np.exp(intercept):.4f}, S = {sig}"
    assert abs(np.exp(intercept) - 1.0) < 0.015, f"fitted D = {-slope:.3e}, a = {np.exp(intercept):.4f}, S = {sig}"


def test_cli_mesh_phantom(cli, tmp_path):
    """`spinwalk phantom -p -i mesh.ply`: the file holds /mask, /fov and /bvf = 0 (the reference never fills the volume fraction of a mesh
    phantom, phantom_ply.cpp:187) and no field map; the mask is the reference's (golden SHA-256)."""
    import meshes

    v, f = meshes.MESHES["torus"]()
    ply = str(tmp_path / "torus.ply")
    meshes.write_ply(ply, v, f, fmt="binary_little_endian", vertex_type="double", extra=True)
    out = str(tmp_path / "mesh.h5")
    r = subprocess.run([cli, "phantom", "-p", "-i", ply, "-f", "100", "-z", "37", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "576 triangles" in r.stdout and "Done." in r.stdout
    assert sorted(h5util.names(out)) == ["bvf", "fov", "mask"]
    gold = np.load(os.path.join(h5util.ROOT, "tests", "golden", "phantom", "mesh_torus.npz"))
    assert _sha(h5util.read(out, "mask")) == str(gold["sha256_37"])
    assert h5util.read(out, "bvf").ravel()[0] == 0.0
    r = subprocess.run([cli, "phantom", "-p", "-f", "100", "-z", "37", "-o", out], capture_output=True, text=True)
    assert r.returncode == 1 and "--ply needs --ply_file" in r.stderr


def test_cli_gpu_info(cli):
    """`spinwalk -g`: the lines of sim::print_device_info (device_helper.cu:47-75), then exit 0 (src/spinwalk.cpp:50-51)."""
    r = subprocess.run([cli, "-g"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for must in ("The latest version of CUDA supported by the driver:", "Number of devices:", "-Compute Capability: 10.0", "-Free GPU Memory:"):
        assert must in r.stdout, r.stdout
