"""host/h5lite.cpp — the HDF5 codec behind the drop-in `spinwalk sim` host (reference: src/sim/h5_helper.h:48-129).

No libhdf5 / h5py exists in this toolchain, so the codec is pinned two ways:
  * the READER against a file written by a real libhdf5: scipy ships MATLAB's `testhdf5_7.4_GLNX86.mat` (HDF5 behind a
    512-byte user block, v0 superblock, v1 object header with an attribute, symbol-table group), whose content is known
    (scipy/io/matlab/tests: testdouble = 0 : pi/4 : 2 pi);
  * the WRITER through the reader (round trips of every type / rank, many datasets => several symbol-table nodes) and
    through byte-level checks of the structures libhdf5 validates when it opens a file (signature, end-of-file address,
    sorted symbol table, B-tree keys, heap names);
  * a SECOND, independent parser (tests/h5spec.py: pure Python, written from the format specification, strict about every
    reserved byte, alignment, key interval and message count libhdf5 enforces) that is itself pinned on the libhdf5-written file
    and must accept, and decode identically, everything the Writer emits."""
import os
import struct

import numpy as np
import pytest

import h5spec
import h5util


def _scipy_mat():
    try:
        import scipy.io

        p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
        return p if os.path.exists(p) else None
    except Exception:
        return None


def test_reads_a_libhdf5_written_file():
    p = _scipy_mat()
    if p is None:
        pytest.skip("scipy's MATLAB v7.3 sample is not installed")
    assert h5util.names(p) == ["testdouble"]
    shape, dt, layout = h5util.info(p, "testdouble")
    assert shape == (9, 1) and dt is np.float64 and layout == 1
    v = h5util.read(p, "testdouble")
    assert np.allclose(v[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    # H5Dread-style conversion to the caller's element type (h5_helper.h:69 read_raw<T>)
    assert np.array_equal(h5util.read(p, "testdouble", np.float32), v.astype(np.float32))
    assert np.array_equal(h5util.read(p, "testdouble", np.uint8), v.astype(np.uint8))


def test_round_trip_reference_output_layout(tmp_path):
    """the five datasets the reference writes (monte_carlo.cu:168-197) with their ranks and types."""
    rng = np.random.default_rng(0)
    K, S, E, trj = 3, 11, 2, 1
    ds = {"M": rng.standard_normal((K, S, E, 3)).astype(np.float32), "XYZ": rng.standard_normal((K, S, trj, 3)).astype(np.float32),
          "T": rng.integers(0, 3, (K, S, E, 1)).astype(np.uint8), "scales": np.arange(K, dtype=np.float32).reshape(K, 1, 1, 1),
          "TE": np.array([0.01, 0.02], np.float32).reshape(E, 1, 1, 1)}
    p = str(tmp_path / "out.h5")
    h5util.write(p, ds)
    assert h5util.names(p) == sorted(ds)
    for k, v in ds.items():
        shape, dt, layout = h5util.info(p, k)
        assert shape == v.shape and dt is v.dtype.type and layout == 1
        assert np.array_equal(h5util.read(p, k), v)


@pytest.mark.parametrize("n_datasets", [1, 8, 9, 40])
def test_round_trip_all_types_and_many_datasets(tmp_path, n_datasets):
    rng = np.random.default_rng(n_datasets)
    ds = {}
    for i in range(n_datasets):
        dt = h5util.DTYPES[i % len(h5util.DTYPES)]
        shape = tuple(int(x) for x in rng.integers(1, 5, size=1 + i % 4))
        ds[f"d{i:02d}_{'x' * (i % 7)}"] = (rng.standard_normal(shape) * 100).astype(dt)
    p = str(tmp_path / "many.h5")
    h5util.write(p, ds)
    assert h5util.names(p) == sorted(ds)
    for k, v in ds.items():
        assert np.array_equal(h5util.read(p, k), v), k


def test_written_file_structure(tmp_path):
    """what libhdf5 checks on open: signature, superblock v0 fields, EOF address == file size, root symbol-table entry with
    cached B-tree / heap addresses, TREE / HEAP / SNOD signatures, names sorted, object headers 8-byte aligned."""
    p = str(tmp_path / "s.h5")
    h5util.write(p, {"mask": np.ones((4, 5, 6), np.uint8), "fieldmap": np.zeros((4, 5, 6), np.float32), "fov": np.array([1e-4] * 3, np.float32)})
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    leaf_k, int_k = struct.unpack("<HH", b[16:20])
    assert (leaf_k, int_k) == (4, 16)
    base, _, eof, _ = struct.unpack("<4Q", b[24:56])
    assert base == 0 and eof == len(b)
    _, root_oh, cache, _, btree, heap = struct.unpack("<QQIIQQ", b[56:96])
    assert cache == 1 and root_oh % 8 == 0
    assert b[root_oh] == 1 and struct.unpack("<H", b[root_oh + 16:root_oh + 18])[0] == 0x0011  # v1 header, symbol-table message
    assert b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    n_children = struct.unpack("<H", b[btree + 6:btree + 8])[0]
    assert n_children == 1
    snod = struct.unpack("<Q", b[btree + 32:btree + 40])[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack("<H", b[snod + 6:snod + 8])[0] == 3
    heap_data = struct.unpack("<Q", b[heap + 24:heap + 32])[0]
    names, ohs = [], []
    for k in range(3):
        off, oh = struct.unpack("<QQ", b[snod + 8 + 40 * k:snod + 24 + 40 * k])
        names.append(b[heap_data + off:b.index(b"\0", heap_data + off)].decode())
        ohs.append(oh)
    assert names == ["fieldmap", "fov", "mask"] and all(o % 8 == 0 and b[o] == 1 for o in ohs)
    key_last = struct.unpack("<Q", b[btree + 40:btree + 48])[0]
    assert b[heap_data + key_last:heap_data + key_last + 5] == b"mask\0"  # right key = largest name of the child


def test_reader_errors(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(RuntimeError, match="not an HDF5 file"):
        h5util.names(str(p))
    q = str(tmp_path / "y.h5")
    h5util.write(q, {"a": np.zeros(3, np.float32)})
    with pytest.raises(RuntimeError, match="does not exist"):
        h5util.read(q, "b")
    trunc = tmp_path / "z.h5"
    trunc.write_bytes(open(q, "rb").read()[:-4])
    with pytest.raises(RuntimeError, match="past the end"):
        h5util.read(str(trunc), "a")


def test_independent_parser_reads_the_libhdf5_written_file():
    """pins tests/h5spec.py itself: user block of 512 bytes, base address 512, absolute end-of-file address, v2 layout message."""
    p = _scipy_mat()
    if p is None:
        pytest.skip("scipy's MATLAB v7.3 sample is not installed")
    f = h5spec.File(p)
    assert (f.user_block, f.base, f.eof) == (512, 512, os.path.getsize(p)) and list(f.links) == ["testdouble"]
    v = f.dataset("testdouble")
    assert v.dtype == np.dtype("<f8") and v.shape == (9, 1) and np.allclose(v[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


@pytest.mark.parametrize("n_datasets", [0, 1, 8, 9, 17, 40, 64, 256])
def test_writer_output_passes_the_independent_parser(tmp_path, n_datasets):
    """every file structure of the Writer (1 .. 32 symbol-table nodes under one B-tree node, all ten element types, ranks 1-4)
    is accepted by the strict parser and decodes to the arrays that were written."""
    rng = np.random.default_rng(100 + n_datasets)
    ds = {}
    for i in range(n_datasets):
        dt = h5util.DTYPES[i % len(h5util.DTYPES)]
        shape = tuple(int(x) for x in rng.integers(1, 6, size=1 + i % 4))
        ds[f"{'ZMa_'[i % 4]}{i:03d}{'y' * (i % 9)}"] = (rng.standard_normal(shape) * 100).astype(dt)
    p = str(tmp_path / "w.h5")
    h5util.write(p, ds)
    f = h5spec.File(p)
    assert sorted(f.links) == sorted(ds) and f.base == 0 and f.eof == os.path.getsize(p)
    for k, v in ds.items():
        a = f.dataset(k)
        assert a.dtype == v.dtype and a.shape == v.shape and np.array_equal(a, v), k


def test_reference_file_layouts_pass_the_independent_parser(tmp_path):
    """the phantom file (phantom_base.cpp:72-105) and the sim output file (monte_carlo.cu:168-197) as this host writes them."""
    rng = np.random.default_rng(7)
    ph = {"mask": (rng.random((7, 5, 6)) < 0.2).astype(np.uint8), "fieldmap": rng.standard_normal((7, 5, 6)).astype(np.float32),
          "fov": np.array([7e-6, 5e-6, 6e-6], np.float32), "bvf": np.array([20.0], np.float32)}
    K, S, E = 2, 13, 3
    out = {"M": rng.standard_normal((K, S, E, 3)).astype(np.float32), "XYZ": rng.standard_normal((K, S, 1, 3)).astype(np.float32),
           "T": rng.integers(0, 2, (K, S, E, 1)).astype(np.uint8), "scales": np.ones((K, 1, 1, 1), np.float32), "TE": np.full((E, 1, 1, 1), 0.02, np.float32)}
    for name, ds in (("phantom.h5", ph), ("out.h5", out)):
        p = str(tmp_path / name)
        h5util.write(p, ds)
        f = h5spec.File(p)
        assert sorted(f.links) == sorted(ds)
        for k, v in ds.items():
            assert np.array_equal(f.dataset(k), v) and f.dataset(k).dtype == v.dtype


def test_independent_parser_rejects_what_libhdf5_rejects(tmp_path):
    """the second opinion is only worth something if it is strict: flip the fields libhdf5 validates and expect a refusal."""
    p = str(tmp_path / "ok.h5")
    h5util.write(p, {"a": np.arange(6, dtype=np.float32).reshape(2, 3), "b": np.arange(4, dtype=np.int16)})
    good = bytearray(open(p, "rb").read())
    h5spec.File(p)

    def broken(edit):
        b = bytearray(good)
        edit(b)
        q = tmp_path / "bad.h5"
        q.write_bytes(bytes(b))
        with pytest.raises(h5spec.FormatError):
            f = h5spec.File(str(q))
            for k in f.links:
                f.dataset(k)

    root_oh = struct.unpack_from("<Q", good, 64)[0]
    broken(lambda b: struct.pack_into("<Q", b, 40, len(b) + 8))           # end-of-file address beyond the file
    broken(lambda b: struct.pack_into("<H", b, root_oh + 2, 2))           # message count of the root object header
    broken(lambda b: b.__setitem__(slice(root_oh + 40, root_oh + 44), b"XXXX"))  # B-tree signature (the node follows the 40-byte header)
    a_oh = h5spec.File(p).links["a"]
    broken(lambda b: b.__setitem__(a_oh + 21, 1))                         # reserved byte of a message header
    broken(lambda b: struct.pack_into("<I", b, a_oh + 8, 100))            # header size not matching its messages
