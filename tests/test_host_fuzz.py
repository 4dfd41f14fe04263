"""Malformed input files must be refused with a message, never crash or hang the host: seeded byte-level mutations of valid HDF5, PLY and
INI files go through the readers behind `spinwalk sim` / `spinwalk phantom -p` (host/h5lite.cpp, host/ply_reader.cpp, host/sim_config.cpp).
(The reference leaves this to libhdf5, happly and std::stof, the last two of which throw through main().)  6000 mutations per reader ran clean
during development; the committed runs are sized for seconds."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

import h5util
import meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 300


def test_hdf5_reader_survives_mutated_files(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "base.h5")
    h5util.write(p, {"mask": (rng.random((4, 5, 6)) < 0.3).astype(np.uint8), "fieldmap": rng.standard_normal((4, 5, 6)).astype(np.float32),
                     "fov": np.array([1e-4] * 3, np.float32), "XYZ": rng.random((7, 3)).astype(np.float32)})
    bases = [open(p, "rb").read()]
    try:
        import scipy.io

        mat = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
        if os.path.exists(mat):
            bases.append(open(mat, "rb").read())
    except Exception:
        pass
    q = str(tmp_path / "m.h5")
    refused = 0
    for it in range(N):
        b = bytearray(bases[it % len(bases)])
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, min(len(b), 2400)))  # the metadata region
            mode = int(rng.integers(0, 3))
            b[pos] = int(rng.integers(0, 256)) if mode == 0 else (0xFF if mode == 1 else b[pos] ^ (1 << int(rng.integers(0, 8))))
        open(q, "wb").write(bytes(b))
        try:
            for nm in h5util.names(q):
                shape, _, _ = h5util.info(q, nm)
                if np.prod(shape, dtype=np.float64) < 1e7:
                    h5util.read(q, nm)
        except (RuntimeError, UnicodeDecodeError):
            refused += 1
    assert 0 < refused < N  # some mutations hit unused bytes, many must be caught


def test_ply_reader_survives_mutated_files(tmp_path):
    rng = np.random.default_rng(2)
    lib = h5util.lib()
    v, f = meshes.icosphere(1)
    bases = []
    for kw in (dict(fmt="ascii", vertex_type="float"), dict(fmt="binary_little_endian", vertex_type="double", extra=True),
               dict(fmt="binary_big_endian", vertex_type="float", index_type="ushort")):
        p = str(tmp_path / "b.ply")
        meshes.write_ply(p, v, f, **kw)
        bases.append(open(p, "rb").read())
    q = str(tmp_path / "m.ply")
    refused = 0
    for it in range(N):
        b = bytearray(bases[it % 3])
        for _ in range(int(rng.integers(1, 5))):
            pos = int(rng.integers(0, len(b)))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                b[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                del b[pos]
            elif mode == 2:
                b.insert(pos, int(rng.integers(32, 127)))
            else:
                del b[pos:]
            if not b:
                b = bytearray(b"p")
        open(q, "wb").write(bytes(b))
        nv, nf, buf = C.c_uint64(0), C.c_uint64(0), C.create_string_buffer(4096)
        if lib.swkh_ply_read(q.encode(), None, C.byref(nv), None, C.byref(nf), buf, len(buf)) != 0:
            assert buf.value, "a refusal carries a message"
            refused += 1
    assert refused > N // 2


def test_config_reader_survives_mutated_files(tmp_path):
    src = os.path.join(ROOT, "tests", "golden", "config")
    names = sorted(f for f in os.listdir(src) if f.endswith(".ini")) if os.path.isdir(src) else []
    if not names:
        pytest.skip("no committed config files")
    rng = np.random.default_rng(3)
    lib = h5util.lib()
    texts = {f: open(os.path.join(src, f), "rb").read() for f in names}
    junk = [b"", b"abc", b"1e999", b"-", b"nan", b" 1 2 x", b"0x10", b"1.5.2", b"9999999999999999999999", b"[", b"]", b"=", b";", b"\x00", b"\xff\xfe"]
    d = str(tmp_path)
    outcomes = set()
    for it in range(N):
        for f, t in texts.items():
            open(os.path.join(d, f), "wb").write(t)
        f = names[int(rng.integers(0, len(names)))]
        lines = texts[f].split(b"\n")
        for _ in range(int(rng.integers(1, 4))):
            i = int(rng.integers(0, len(lines)))
            mode = int(rng.integers(0, 4))
            if mode == 0 and b"=" in lines[i]:
                lines[i] = lines[i].split(b"=")[0] + b"= " + junk[int(rng.integers(0, len(junk)))]
            elif mode == 1:
                lines[i] = junk[int(rng.integers(0, len(junk)))]
            elif mode == 2 and len(lines) > 1:
                del lines[i]
            elif lines[i]:
                bb = bytearray(lines[i])
                bb[int(rng.integers(0, len(bb)))] = int(rng.integers(0, 256))
                lines[i] = bytes(bb)
        open(os.path.join(d, f), "wb").write(b"\n".join(lines))
        buf = C.create_string_buffer(1 << 20)
        outcomes.add(lib.swkh_config_json(os.path.join(d, f).encode(), 0, buf, len(buf)) == 0)
    assert outcomes == {True, False}
