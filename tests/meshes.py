"""Small closed triangle meshes for the `spinwalk phantom -p` tests, and PLY writers (ascii / binary, the layouts happly reads)."""
import struct

import numpy as np


def icosphere(subdiv=2, radius=1.0, centre=(0.0, 0.0, 0.0)):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.asarray(v) * radius + np.asarray(centre, np.float64), np.asarray(f, np.uint64)


def box(size=(1.0, 0.7, 0.4)):
    s = np.asarray(size, np.float64) / 2
    v = np.asarray([(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], np.float64) * s
    f = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
    return v, np.asarray(f, np.uint64)


def torus(R=1.0, r=0.35, nu=24, nv=12, tilt=0.3):
    u = np.arange(nu) * 2 * np.pi / nu
    w = np.arange(nv) * 2 * np.pi / nv
    U, W = np.meshgrid(u, w, indexing="ij")
    v = np.stack([(R + r * np.cos(W)) * np.cos(U), (R + r * np.cos(W)) * np.sin(U), r * np.sin(W)], -1).reshape(-1, 3)
    c, s = np.cos(tilt), np.sin(tilt)
    v = v @ np.asarray([[1, 0, 0], [0, c, -s], [0, s, c]]).T @ np.asarray([[c, 0, s], [0, 1, 0], [-s, 0, c]]).T
    f = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, i * nv + (j + 1) % nv
            c2, d = ((i + 1) % nu) * nv + j, ((i + 1) % nu) * nv + (j + 1) % nv
            f += [(a, c2, d), (a, d, b)]
    return v, np.asarray(f, np.uint64)


def two_bodies():
    """two separate closed surfaces in one mesh (a sphere and a box): the ray parity rule handles several crossings per row"""
    v1, f1 = icosphere(1, 0.5, (-0.6, 0.1, 0.0))
    v2, f2 = box((0.6, 0.9, 0.5))
    v2 = v2 + np.asarray([0.7, -0.1, 0.2])
    return np.concatenate([v1, v2]), np.concatenate([f1, f2 + np.uint64(len(v1))])


MESHES = {
    "icosphere": lambda: icosphere(2, 0.02),     # 320 faces, 40 um across (vertex units are mm, phantom_ply.cpp:162)
    "box": lambda: tuple(a * s for a, s in zip(box(), (0.05, 1))),
    "torus": lambda: tuple(a * s for a, s in zip(torus(), (0.02, 1))),
    "two_bodies": lambda: tuple(a * s for a, s in zip(two_bodies(), (0.03, 1))),
}


def write_ply(path, v, f, fmt="ascii", vertex_type="float", index_type="int", list_name="vertex_indices", extra=False):
    """fmt: ascii | binary_little_endian | binary_big_endian.  extra=True adds a per-vertex `quality` property, a comment and an `edge` element
    the reader must skip."""
    v = np.asarray(v, np.float64)
    f = np.asarray(f, np.int64)
    vt = {"float": "f", "double": "d"}[vertex_type]
    it = {"int": "i", "uint": "I", "uchar": "B", "ushort": "H"}[index_type]
    head = ["ply", f"format {fmt} 1.0", "comment made by tests/meshes.py", f"element vertex {len(v)}"]
    head += [f"property {vertex_type} x", f"property {vertex_type} y", f"property {vertex_type} z"]
    if extra:
        head += ["property uchar quality"]
    head += [f"element face {len(f)}", f"property list uchar {index_type} {list_name}"]
    if extra:
        head += ["element edge 1", "property int vertex1", "property int vertex2"]
    head += ["end_header"]
    with open(path, "wb") as o:
        o.write(("\n".join(head) + "\n").encode())
        if fmt == "ascii":
            for p in v:
                o.write((" ".join(repr(float(np.float32(c) if vertex_type == "float" else c)) for c in p) + (" 7" if extra else "") + "\n").encode())
            for t in f:
                o.write(("3 " + " ".join(str(int(i)) for i in t) + "\n").encode())
            if extra:
                o.write(b"0 1\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            for p in v:
                o.write(struct.pack(e + "3" + vt, *p))
                if extra:
                    o.write(struct.pack("B", 7))
            for t in f:
                o.write(struct.pack(e + "B3" + it, 3, *[int(i) for i in t]))
            if extra:
                o.write(struct.pack(e + "2i", 0, 1))


def stored_vertices(v, vertex_type="float"):
    """the values a reader gets back (float32 round trip for `float` files)"""
    v = np.asarray(v, np.float64)
    return v.astype(np.float32).astype(np.float64) if vertex_type == "float" else v
