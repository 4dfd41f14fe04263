"""Synthetic phantoms of the shapes the reference's `phantom` subcommand makes (host-side, numpy).

These are bench/test inputs, not part of the accelerated path.  Formulas restate
src/phantom/phantom_cylinder.cpp:85-130,183-275 (parallel cylinders along z, analytic dB of an
infinite cylinder, doi:10.1016/j.neuroimage.2017.09.015) and src/phantom/phantom_sphere.cpp:79-198
(random spheres, dipole field outside, doi:10.1002/nbm.1079).  Voxel centres follow
src/phantom/phantom_base.cpp:107-143.  The random placement uses numpy's generator, so shapes are
statistically — not bitwise — those of the reference generator.
Layout: row-major [x][y][z] (x slowest), mask uint8, fieldmap float32 in Tesla at B0 = 1 T.
"""
from __future__ import annotations

import numpy as np


def _centres(fov_um: float, n: int) -> np.ndarray:
    return ((np.arange(n, dtype=np.float64) + 0.5) * (fov_um / n)).astype(np.float32)


def cylinder_phantom(n: int, fov_um: float, radius_um: float = 8.0, bvf_pct: float = 4.0, Y: float = 0.78,
                     dchi: float = 0.11e-6, orientation_deg: float = 90.0, seed: int = 0, fieldmap: bool = True,
                     nz: int | None = None, planar: bool = False):
    """Random non-overlapping cylinders parallel to z.  Returns (mask[n,n,nz] u8, fieldmap f32 or None, fov_m[3] f32).

    planar=True returns the 2-D [n,n] slice only (the field is z-invariant) so that a caller can
    broadcast it on the GPU without materialising n^3 voxels on the host."""
    nz = n if nz is None else nz
    rng = np.random.default_rng(seed)
    g = _centres(fov_um, n).astype(np.float64)
    X, Yg = np.meshgrid(g, g, indexing="ij")
    mask2 = np.zeros((n, n), np.uint8)
    pts, radii = [], []
    target = bvf_pct / 100.0
    tries = 0
    while mask2.mean() < target and tries < 100000:
        tries += 1
        r = radius_um if radius_um > 0 else rng.random() * -radius_um
        c = rng.random(2) * (fov_um + 2 * r) - r
        if any((c[0] - p[0]) ** 2 + (c[1] - p[1]) ** 2 < (r + q) ** 2 for p, q in zip(pts, radii)):
            continue
        d2 = (X - c[0]) ** 2 + (Yg - c[1]) ** 2
        add = (d2 <= r * r)
        if (mask2 | add).mean() > 1.02 * target or add.sum() == 0:
            continue
        mask2 |= add.astype(np.uint8)
        pts.append(c)
        radii.append(r)
    fm2 = None
    if fieldmap:
        th = np.deg2rad(orientation_deg)
        c2, s2 = np.cos(th) ** 2, 1.0 - np.cos(th) ** 2
        b0p = np.array([np.sin(th), 0.0])  # roty(orientation) of (0,0,1), projected on the xy plane
        nb = np.linalg.norm(b0p)
        b0p = b0p / nb if nb > 0 else b0p
        fm2 = np.zeros((n, n), np.float64)
        k = 2 * np.pi * (1 - Y) * dchi
        for c, r in zip(pts, radii):
            dx, dy = X - c[0], Yg - c[1]
            d2 = dx * dx + dy * dy
            box = (np.abs(dx) <= 20 * (r + fov_um / n)) & (np.abs(dy) <= 20 * (r + fov_um / n))
            with np.errstate(divide="ignore", invalid="ignore"):
                cphi = (dx * b0p[0] + dy * b0p[1]) / np.sqrt(d2)
                outside = k * (r * r / d2) * (2 * cphi * cphi - 1) * s2
            inside = k * (c2 - 1.0 / 3.0)
            fm2 += np.where(box, np.where(d2 > r * r, np.nan_to_num(outside), inside), 0.0)
        fm2 = fm2.astype(np.float32)
    fov_m = np.array([fov_um, fov_um, fov_um * nz / n], np.float32) * np.float32(1e-6)
    if planar:
        return mask2, fm2, fov_m
    mask = np.ascontiguousarray(np.broadcast_to(mask2[:, :, None], (n, n, nz)))
    fm = None if fm2 is None else np.ascontiguousarray(np.broadcast_to(fm2[:, :, None], (n, n, nz)))
    return mask, fm, fov_m


def sphere_phantom(n: int, fov_um: float, radius_um: float = -20.0, vf_pct: float = 40.0, Y: float = 0.78,
                   dchi: float = 0.11e-6, seed: int = 0, fieldmap: bool = False, max_spheres: int = 100000):
    """Random non-overlapping spheres.  Returns (mask[n,n,n] u8, fieldmap f32 or None, fov_m[3] f32)."""
    rng = np.random.default_rng(seed)
    g = _centres(fov_um, n)
    mask = np.zeros((n, n, n), np.uint8)
    fm = np.zeros((n, n, n), np.float32) if fieldmap else None
    h = fov_um / n
    pts, radii, vol = [], [], 0.0
    target = vf_pct / 100.0 * fov_um**3
    tries = 0
    while 0.95 * vol < target and len(pts) < max_spheres and tries < 2000000:
        tries += 1
        r = radius_um if radius_um > 0 else max(rng.random() * -radius_um, 0.5 * h)
        c = rng.random(3) * fov_um
        if pts:
            P = np.asarray(pts)
            if np.any(((P - c) ** 2).sum(1) < (np.asarray(radii) + r) ** 2):
                continue
        pts.append(c)
        radii.append(r)
        vol += 4 * np.pi / 3 * r**3
    k = 4 * np.pi * (1 - Y) * dchi
    for c, r in zip(pts, radii):
        reach = 20 * (r + h) if fieldmap else (r + 2 * h)
        lo = np.maximum(0, np.floor((c - reach) / h).astype(int))
        hi = np.minimum(n, np.ceil((c + reach) / h).astype(int) + 1)
        if np.any(hi <= lo):
            continue
        sx, sy, sz = (slice(lo[i], hi[i]) for i in range(3))
        dx = (g[sx] - c[0])[:, None, None]
        dy = (g[sy] - c[1])[None, :, None]
        dz = (g[sz] - c[2])[None, None, :]
        d2 = dx * dx + dy * dy + dz * dz
        inside = d2 <= r * r
        mask[sx, sy, sz] |= inside.astype(np.uint8)
        if fieldmap:
            with np.errstate(divide="ignore", invalid="ignore"):
                f = k * r**3 / (d2 * np.sqrt(d2)) * (dz * dz / d2 - 1.0 / 3.0)
            fm[sx, sy, sz] += np.where(inside, 0.0, np.nan_to_num(f)).astype(np.float32)
    fov_m = np.full(3, fov_um, np.float32) * np.float32(1e-6)
    return mask, fm, fov_m


def sphere_lattice_phantom(n: int, fov_um: float, cell_um: float = 40.0, vf_pct: float = 40.0, seed: int = 0):
    """Non-overlapping spheres, one per cell of a cubic lattice, random radius and random jitter inside the cell, scaled
    to the requested volume fraction (<= ~45 %).  O(voxels) and vectorised: the random sequential placement of
    `sphere_phantom` (phantom_sphere.cpp:79-119) does not reach 40 % in reasonable time at 400^3.  Same kind of substrate
    (closed permeable/impermeable compartments of radius <= cell/2) for the PGSE bench workload.
    Returns (mask[n,n,n] u8, None, fov_m[3] f32)."""
    rng = np.random.default_rng(seed)
    nc = max(1, int(round(fov_um / cell_um)))
    a = fov_um / nc
    r = rng.uniform(0.6, 1.0, size=(nc, nc, nc))
    r *= (vf_pct / 100.0 * a**3 / (4 * np.pi / 3 * (r**3).mean())) ** (1 / 3)
    r = np.minimum(r, 0.5 * a)
    c = (rng.random((3, nc, nc, nc)) - 0.5) * 2 * (0.5 * a - r)  # jitter keeps the sphere inside its cell
    g = _centres(fov_um, n)
    ci = np.minimum((g / a).astype(np.int64), nc - 1)
    loc = g - (ci + 0.5) * a  # position relative to the cell centre
    mask = np.zeros((n, n, n), np.uint8)
    for ix in range(nc):  # slab by slab along x keeps the temporaries small
        sx = np.nonzero(ci == ix)[0]
        dx = loc[sx][:, None, None] - c[0, ix][ci][:, ci][None, :, :]
        dy = loc[None, :, None] - c[1, ix][ci][:, ci][None, :, :]
        dz = loc[None, None, :] - c[2, ix][ci][:, ci][None, :, :]
        rr = r[ix][ci][:, ci][None, :, :]
        mask[sx] = (dx * dx + dy * dy + dz * dz <= rr * rr).astype(np.uint8)
    fov_m = np.full(3, fov_um, np.float32) * np.float32(1e-6)
    return mask, None, fov_m


def icosphere_mesh(subdiv: int = 5, radius_mm: float = 0.2):
    """A closed triangle mesh (subdivided icosahedron) as `spinwalk phantom -p` input: (vertices float64 [nv,3] in mm, faces uint64 [nf,3]),
    20 * 4**subdiv triangles.  Synthetic bench / test input."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = np.asarray([(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
                    (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], np.int64)
    v = np.asarray(v, np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        uniq, inv = np.unique(edges, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = inv.reshape(3, -1) + len(v)  # midpoint index of edges ab, bc, ca of every face
        v = np.concatenate([v, mid])
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = m[0], m[1], m[2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    return v * radius_mm, f.astype(np.uint64)


def write_ply(path: str, vertices, faces) -> None:
    """binary_little_endian PLY (double x/y/z, int vertex_indices) — the file `spinwalk phantom -p -i` reads."""
    v = np.ascontiguousarray(vertices, "<f8")
    f = np.ascontiguousarray(faces, np.int64)
    rec = np.zeros(len(f), dtype=[("n", "u1"), ("i", "<i4", 3)])
    rec["n"] = 3
    rec["i"] = f
    with open(path, "wb") as o:
        o.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(v)}\nproperty double x\nproperty double y\nproperty double z\n"
                 f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        o.write(v.tobytes())
        o.write(rec.tobytes())
