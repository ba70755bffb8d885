"""Synthetic closed triangle meshes for benchmarks and parity tests.

All generators return ``(vertices float32 [Nv,3], triangles uint32 [T,3])``: indexed (shared
vertices), closed, outward oriented, deterministic for a given seed.  ``workload(name)`` builds the
BASELINE.json configurations (mesh + grid placement).

The reference has no mesh generators; its tests use one 36-triangle STL and a procedural unit cube
(/root/reference/tests/test_correctness.cpp:30-62, python/tests/test_sdfgen.py:15-58).
"""
from __future__ import annotations

import numpy as np

__all__ = ["icosphere", "uv_sphere", "blob", "torus", "unit_cube", "shuffle_triangles", "stacked_workload",
           "workload", "WORKLOADS"]


def _unique_midpoints(tris: np.ndarray, nverts: int):
    """Midpoint vertex ids for the three edges of every triangle (shared edges share a midpoint)."""
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], axis=0).astype(np.int64)
    e.sort(axis=1)
    key = e[:, 0] * np.int64(nverts) + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    pairs = np.stack([uniq // nverts, uniq % nverts], axis=1)
    t = tris.shape[0]
    mid = nverts + inv.reshape(3, t).T          # [T,3]: midpoints of edges (01, 12, 20)
    return pairs, mid


def icosphere(level: int, radius: float = 1.0, center=(0.0, 0.0, 0.0)):
    """Icosahedron subdivided ``level`` times and projected to the sphere: T = 20 * 4**level."""
    g = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, g, 0], [1, g, 0], [-1, -g, 0], [1, -g, 0],
                  [0, -1, g], [0, 1, g], [0, -1, -g], [0, 1, -g],
                  [g, 0, -1], [g, 0, 1], [-g, 0, -1], [-g, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        pairs, mid = _unique_midpoints(f, v.shape[0])
        m = v[pairs[:, 0]] + v[pairs[:, 1]]
        m /= np.linalg.norm(m, axis=1, keepdims=True)
        v = np.concatenate([v, m], axis=0)
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = mid[:, 0], mid[:, 1], mid[:, 2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0)
    v = v * radius + np.asarray(center, dtype=np.float64)
    return v.astype(np.float32), f.astype(np.uint32)


def uv_sphere(nlat: int, nlon: int, radius: float = 1.0, center=(0.0, 0.0, 0.0)):
    """Latitude/longitude sphere: T = 2 * nlon * (nlat - 1), Nv = 2 + (nlat-1)*nlon."""
    th = np.pi * np.arange(1, nlat) / nlat                      # interior rings
    ph = 2.0 * np.pi * np.arange(nlon) / nlon
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)), np.outer(np.sin(th), np.sin(ph)),
                     np.outer(np.cos(th), np.ones(nlon))], axis=-1).reshape(-1, 3)
    v = np.concatenate([[[0, 0, 1.0]], ring, [[0, 0, -1.0]]], axis=0)
    south = v.shape[0] - 1
    idx = lambda r, c: 1 + r * nlon + (c % nlon)
    c = np.arange(nlon)
    tris = [np.stack([np.zeros(nlon, np.int64), idx(0, c), idx(0, c + 1)], 1)]
    for r in range(nlat - 2):
        tris.append(np.stack([idx(r, c), idx(r + 1, c), idx(r + 1, c + 1)], 1))
        tris.append(np.stack([idx(r, c), idx(r + 1, c + 1), idx(r, c + 1)], 1))
    tris.append(np.stack([np.full(nlon, south, np.int64), idx(nlat - 2, c + 1), idx(nlat - 2, c)], 1))
    f = np.concatenate(tris, axis=0)
    v = v * radius + np.asarray(center, dtype=np.float64)
    return v.astype(np.float32), f.astype(np.uint32)


def blob(nlat: int, nlon: int, radius: float, center=(0.0, 0.0, 0.0), amp: float = 0.15, seed: int = 1234):
    """'Bunny-scale blob': UV sphere with radial perturbation r*(1 + amp*sum_3 sin(k_m.p + phi_m))/..."""
    v, f = uv_sphere(nlat, nlon, 1.0)
    rng = np.random.default_rng(seed)
    kvec = rng.uniform(-4.0, 4.0, size=(3, 3))
    phase = rng.uniform(0.0, 2.0 * np.pi, size=3)
    p = v.astype(np.float64)
    bump = np.sin(p @ kvec.T + phase).sum(axis=1)
    p = p * (radius * (1.0 + amp * bump / 3.0))[:, None] + np.asarray(center, dtype=np.float64)
    return p.astype(np.float32), f


def torus(nu: int, nv: int, R: float, r: float, center=(0.0, 0.0, 0.0), jitter: float = 0.0, seed: int = 2025):
    """Torus of nu x nv quads (T = 2*nu*nv); ``jitter`` displaces vertices along the normal by
    jitter * (local edge length) * U(-1,1)."""
    u = 2.0 * np.pi * np.arange(nu) / nu
    w = 2.0 * np.pi * np.arange(nv) / nv
    cu, su = np.cos(u)[:, None], np.sin(u)[:, None]
    cw, sw = np.cos(w)[None, :], np.sin(w)[None, :]
    rr = np.full((nu, nv), r, dtype=np.float64)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        edge = min(2.0 * np.pi * (R - r) / nu, 2.0 * np.pi * r / nv)
        rr = rr + jitter * edge * rng.uniform(-1.0, 1.0, size=(nu, nv))
    x = (R + rr * cw) * cu
    y = (R + rr * cw) * su
    z = rr * sw * np.ones_like(cu)
    v = np.stack([x, y, z], axis=-1).reshape(-1, 3) + np.asarray(center, dtype=np.float64)
    a = np.arange(nu)[:, None]
    b = np.arange(nv)[None, :]
    i00 = (a * nv + b).ravel()
    i10 = (((a + 1) % nu) * nv + b).ravel()
    i01 = (a * nv + (b + 1) % nv).ravel()
    i11 = (((a + 1) % nu) * nv + (b + 1) % nv).ravel()
    f = np.concatenate([np.stack([i00, i10, i11], 1), np.stack([i00, i11, i01], 1)], axis=0)
    return v.astype(np.float32), f.astype(np.uint32)


def unit_cube(lo: float = 0.0, hi: float = 1.0):
    """12-triangle axis-aligned cube (the shape the reference's tests build procedurally)."""
    v = np.array([[lo, lo, lo], [hi, lo, lo], [hi, hi, lo], [lo, hi, lo],
                  [lo, lo, hi], [hi, lo, hi], [hi, hi, hi], [lo, hi, hi]], dtype=np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                  [2, 3, 7], [2, 7, 6], [0, 4, 7], [0, 7, 3], [1, 2, 6], [1, 6, 5]], dtype=np.uint32)
    return v, f


def shuffle_triangles(f: np.ndarray, seed: int):
    """Permute triangle order (tests tie-break determinism / removes memory locality)."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(f[rng.permutation(f.shape[0])])


def _placed(n, L=1.0, off=0.37):
    """Cubic n^3 grid of side L centred at the origin, shifted by a non-lattice offset of off*dx."""
    dx = np.float32(L / n)
    origin = (np.float32(-0.5 * L) + np.float32(off) * dx) * np.ones(3, dtype=np.float32)
    return origin, dx


# name -> description used in bench.py's config.workload
WORKLOADS = {
    "c1_blob_256": "synthetic 70312-triangle perturbed UV-sphere blob at 256^3 (BASELINE configs[1])",
    "c2_icosphere_512": "synthetic 1310720-triangle icosphere (level 8) at 512^3 (BASELINE configs[2], metric config)",
    "c3_torus_1024": "synthetic 5005448-triangle jittered torus at 1024^3 (BASELINE configs[3])",
    "c4_mix_2048": "synthetic 10.0M-triangle icosphere(level 9)+inner torus at 2048^3 (BASELINE configs[4])",
}


def workload(name: str, n: int | None = None, shuffle: bool = False):
    """Build a BASELINE.json configuration.  ``n`` overrides the grid edge (down-scaled twins for
    parity runs keep the same mesh).  Returns dict(vertices, triangles, origin, dx, ni, nj, nk, name)."""
    L = 1.0
    if name == "c1_blob_256":
        n = n or 256
        v, f = blob(188, 188, 0.35 * L, seed=1234)
    elif name == "c2_icosphere_512":
        n = n or 512
        v, f = icosphere(8, 0.4 * L)
    elif name == "c3_torus_1024":
        n = n or 1024
        v, f = torus(1582, 1582, 0.30 * L, 0.12 * L, jitter=0.2, seed=2025)
    elif name == "c4_mix_2048":
        n = n or 2048
        v1, f1 = icosphere(9, 0.42 * L)
        v2, f2 = torus(1550, 1550, 0.20 * L, 0.08 * L, jitter=0.2, seed=2026)
        v = np.concatenate([v1, v2], axis=0)
        f = np.concatenate([f1, f2 + np.uint32(v1.shape[0])], axis=0)
    else:
        raise ValueError(f"unknown workload {name!r}; choose from {sorted(WORKLOADS)}")
    if shuffle:
        f = shuffle_triangles(f, 99)
    origin, dx = _placed(n, L)
    return dict(name=name, vertices=np.ascontiguousarray(v), triangles=np.ascontiguousarray(f),
                origin=origin, dx=float(dx), ni=n, nj=n, nk=n)


def stacked_workload(copies: int, n: int = 512, level: int = 8):
    """Weak-scaling workload for `copies` GPUs: `copies` icospheres (level 8, radius 0.4 L, the C2 mesh)
    stacked along z, one per n^3 block of an n x n x (copies*n) grid.  With z-slab sharding every rank gets
    one block and one sphere's worth of surface, so per-GPU work equals the single-GPU C2 configuration."""
    L = 1.0
    v0, f0 = icosphere(level, 0.4 * L)
    vs, fs = [], []
    for g in range(copies):
        vs.append(v0 + np.array([0.0, 0.0, g * L], dtype=np.float32))
        fs.append(f0 + np.uint32(g * v0.shape[0]))
    v = np.ascontiguousarray(np.concatenate(vs, axis=0), dtype=np.float32)
    f = np.ascontiguousarray(np.concatenate(fs, axis=0), dtype=np.uint32)
    origin, dx = _placed(n, L)
    return dict(name=f"c2_icosphere_{n}_stack{copies}", vertices=v, triangles=f, origin=origin, dx=float(dx),
                ni=n, nj=n, nk=n * copies)
