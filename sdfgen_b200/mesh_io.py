"""Mesh / .sdf file side-cars of the hot path (SURVEY.md section 8f "next" rows), numpy only.

These are the data formats either side of make_level_set3, kept to what the Python boundary needs:

    load_mesh   OBJ (triangles, quads/polygons as a fan; 'v' and 'f' records only) and STL (binary or
                ASCII, autodetected; three unshared vertices per facet) -- same vertex/face order as
                /root/reference/common/mesh_io_obj.cpp:21-157 and mesh_io_stl.cpp:157-165, 309-332.
    save_sdf    36-byte header (3 x int32 dims, 3 x float32 min, 3 x float32 max) + float32 values in
    load_sdf    k-fastest order, /root/reference/common/sdf_io.cpp:10-74 and :76-147.
"""
from __future__ import annotations

import os
import struct

import numpy as np


def _bounds(v):
    return tuple(float(x) for x in v.min(axis=0)), tuple(float(x) for x in v.max(axis=0))


def _load_obj(path):
    verts, faces = [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if len(line) < 2:
                continue
            if line[0] == "v" and line[1] in " \t":
                p = line.split()
                try:
                    verts.append((float(p[1]), float(p[2]), float(p[3])))
                except (IndexError, ValueError):
                    continue
            elif line[0] == "f" and line[1] in " \t":
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                for i in range(1, len(idx) - 1):
                    faces.append((idx[0] - 1, idx[i] - 1, idx[i + 1] - 1))
    if not verts or not faces:
        raise RuntimeError("Failed to load mesh: " + path)
    v = np.asarray(verts, dtype=np.float32)
    t = (np.asarray(faces, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint32)
    return v, t


def _load_stl(path):
    raw = open(path, "rb").read()
    is_binary = False
    if len(raw) >= 84:
        n = struct.unpack_from("<I", raw, 80)[0]
        is_binary = (84 + 50 * n == len(raw))
    if is_binary:
        rec = np.frombuffer(raw, dtype=np.uint8, count=50 * n, offset=84).reshape(n, 50)
        v = rec[:, 12:48].copy().view(np.float32).reshape(n * 3, 3)
    else:
        pts = []
        for line in raw.decode("ascii", "replace").splitlines():
            s = line.split()
            if len(s) == 4 and s[0].lower() == "vertex":
                pts.append((float(s[1]), float(s[2]), float(s[3])))
        if not pts or len(pts) % 3:
            raise RuntimeError("Failed to load mesh: " + path)
        v = np.asarray(pts, dtype=np.float32)
    if v.shape[0] == 0:
        raise RuntimeError("Failed to load mesh: " + path)
    t = np.arange(v.shape[0], dtype=np.uint32).reshape(-1, 3)
    return np.ascontiguousarray(v), t


def load_mesh(filename: str):
    """Returns (vertices float32 [N,3], triangles uint32 [M,3], (min_xyz, max_xyz))."""
    ext = os.path.splitext(filename)[1].lower()
    if not os.path.exists(filename):
        raise RuntimeError("Failed to load mesh: " + filename)
    if ext == ".obj":
        v, t = _load_obj(filename)
    elif ext == ".stl":
        v, t = _load_stl(filename)
    else:
        raise RuntimeError("Failed to load mesh: " + filename)
    return v, t, _bounds(v)


def save_sdf(filename: str, sdf_array, origin, dx) -> None:
    a = np.asarray(sdf_array)
    if a.ndim != 3:
        raise ValueError("SDF array must be 3-dimensional")
    if 0 in a.shape:
        raise ValueError("SDF array dimensions cannot be zero")
    a = np.ascontiguousarray(a, dtype=np.float32)
    o = np.asarray([origin[0], origin[1], origin[2]], dtype=np.float32)
    mx = (o + np.asarray(a.shape, dtype=np.float32) * np.float32(dx)).astype(np.float32)
    with open(filename, "wb") as f:
        f.write(struct.pack("<3i", *a.shape))
        f.write(o.tobytes())
        f.write(mx.tobytes())
        f.write(a.tobytes())          # C order (nx,ny,nz) == k fastest


def load_sdf(filename: str):
    """Returns (sdf float32 [nx,ny,nz], origin, dx, (min, max))."""
    with open(filename, "rb") as f:
        hdr = f.read(36)
        if len(hdr) != 36:
            raise RuntimeError("Failed to read SDF file: " + filename)
        nx, ny, nz = struct.unpack_from("<3i", hdr, 0)
        mn = struct.unpack_from("<3f", hdr, 12)
        mx = struct.unpack_from("<3f", hdr, 24)
        if nx <= 0 or ny <= 0 or nz <= 0:
            raise RuntimeError("Failed to read SDF file: " + filename)
        data = np.fromfile(f, dtype=np.float32, count=nx * ny * nz)
    if data.size != nx * ny * nz:
        raise RuntimeError("Failed to read SDF file: " + filename)
    dx = (np.float32(mx[0]) - np.float32(mn[0])) / np.float32(nx)
    return data.reshape(nx, ny, nz), tuple(mn), float(dx), (tuple(mn), tuple(mx))
