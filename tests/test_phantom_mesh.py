"""CPU tests of the mesh-phantom path (`spinwalk phantom -p`): the oracle's restatement of src/phantom/phantom_ply.cpp against goldens written by
the reference itself (and against the reference library where it is built), and the host PLY reader (host/ply_reader.cpp) against the values the
files were written with — and, through the reference's happly, against what the reference reads from the same files."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import h5util
import meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "phantom")
SIZES = ((60.0, 24), (100.0, 37))


@pytest.fixture(scope="module")
def pp(oracle):
    from oracle import pyphantom

    return pyphantom


def read_ply(path):
    lib = h5util.lib()
    nv, nf = C.c_uint64(0), C.c_uint64(0)
    buf = C.create_string_buffer(4096)
    if lib.swkh_ply_read(path.encode(), None, C.byref(nv), None, C.byref(nf), buf, len(buf)) != 0:
        raise RuntimeError(buf.value.decode())
    v = np.zeros((nv.value, 3), np.float64)
    f = np.zeros((nf.value, 3), np.uint64)
    assert lib.swkh_ply_read(path.encode(), v.ctypes.data_as(C.c_void_p), C.byref(nv), f.ctypes.data_as(C.c_void_p), C.byref(nf), buf, len(buf)) == 0
    return v, f


@pytest.mark.parametrize("name", sorted(meshes.MESHES))
def test_mesh_oracle_matches_reference_golden(pp, name):
    v, f = meshes.MESHES[name]()
    gold = np.load(os.path.join(GOLD, "mesh_" + name + ".npz"))
    for fov, res in SIZES:
        mask = pp.oracle_mesh(fov, res, v, f)
        assert mask.sum() == gold[f"inside_{res}"] and mask.sum() > 0
        assert np.array_equal(mask[:, :, res // 2], gold[f"slice_{res}"])
        assert hashlib.sha256(mask.tobytes()).hexdigest() == str(gold[f"sha256_{res}"])


FORMATS = [dict(fmt="ascii", vertex_type="float"), dict(fmt="ascii", vertex_type="double", index_type="uint", list_name="vertex_index"),
           dict(fmt="binary_little_endian", vertex_type="double", extra=True), dict(fmt="binary_big_endian", vertex_type="float", index_type="ushort"),
           dict(fmt="binary_little_endian", vertex_type="float", index_type="uchar", extra=True)]


@pytest.mark.parametrize("kw", FORMATS, ids=lambda k: "-".join(str(v) for v in k.values()))
def test_ply_reader_returns_the_stored_mesh(tmp_path, kw):
    v, f = meshes.MESHES["two_bodies"]()
    path = str(tmp_path / "m.ply")
    meshes.write_ply(path, v, f, **kw)
    gv, gf = read_ply(path)
    assert np.array_equal(gf, f)
    assert np.array_equal(gv, meshes.stored_vertices(v, kw["vertex_type"]))


def test_ply_reader_and_oracle_equal_the_reference_on_the_same_file(pp, tmp_path):
    """What the reference makes of a PLY FILE (happly + phantom::ply) == oracle(mesh as read by host/ply_reader.cpp), for every body format."""
    if not pp.have_ref():
        pytest.skip("oracle/_ref/libswref_gen.so not built (no reference tree here)")
    for name in ("torus", "two_bodies"):
        v, f = meshes.MESHES[name]()
        for i, kw in enumerate(FORMATS):
            if kw.get("index_type") == "uchar" and len(v) > 255:
                continue
            path = str(tmp_path / f"{name}{i}.ply")
            meshes.write_ply(path, v, f, **kw)
            ref, _ = pp.reference_mesh(80.0, 29, path)
            gv, gf = read_ply(path)
            assert np.array_equal(ref, pp.oracle_mesh(80.0, 29, gv, gf)), (name, kw)


def test_ply_reader_refusals(tmp_path):
    v, f = meshes.box()
    p = str(tmp_path / "quad.ply")
    open(p, "w").write("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
                       "property list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    with pytest.raises(RuntimeError, match="Only triangular mesh is supported"):  # phantom_ply.cpp:150-154
        read_ply(p)
    with pytest.raises(RuntimeError, match="could not open"):
        read_ply(str(tmp_path / "missing.ply"))
    p = str(tmp_path / "ints.ply")
    open(p, "w").write("ply\nformat ascii 1.0\nelement vertex 3\nproperty int x\nproperty int y\nproperty int z\nelement face 1\n"
                       "property list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n1 1 0\n3 0 1 2\n")
    with pytest.raises(RuntimeError, match="float or double"):
        read_ply(p)
    p = str(tmp_path / "short.ply")
    meshes.write_ply(p, v, f, fmt="binary_little_endian")
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-5])
    with pytest.raises(RuntimeError, match="ends early"):
        read_ply(p)
    p = str(tmp_path / "range.ply")
    meshes.write_ply(p, v, f + np.uint64(3), fmt="ascii")
    with pytest.raises(RuntimeError, match="out of range"):
        read_ply(p)
