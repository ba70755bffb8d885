"""sdfgen_b200 -- B200-native drop-in for SDFGenFast's mesh -> signed-distance-grid path.

Python mirror of the reference's ``sdfgen`` package for this one path, same names, arguments and
error behaviour:

    generate_sdf         /root/reference/python/sdfgen_py.cpp:160-218, :337-343
    is_gpu_available     /root/reference/common/sdfgen_unified.cpp:19-28
    generate_from_mesh   /root/reference/python/sdfgen.py:47-142
    generate_from_file   /root/reference/python/sdfgen.py:145-265   (loaders: sdfgen_b200.mesh_io)

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/sdfb.h
(sdfgen_b200/libsdfb.so).  There is no CPU backend and no fallback: ``backend="cpu"`` raises, a
missing library is an ImportError, a missing B200 is a RuntimeError.  Results follow the reference's
single-threaded CPU semantics.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import Plan, SdfbError

__version__ = "0.2.0"

__all__ = ["generate_sdf", "generate_sdf_batch", "generate_sdf_file", "generate_sdf_debug", "generate_from_mesh", "generate_from_file",
           "is_gpu_available", "load_mesh", "save_sdf", "load_sdf", "Plan", "SdfbError", "launch_count", "trim_memory"]


def is_gpu_available() -> bool:
    """True when a usable sm_100 device is present (sdfgen::is_gpu_available)."""
    return _lib.lib().sdfb_device_count() > 0


def trim_memory() -> None:
    """Return the device memory libsdfb keeps for reuse between calls to the driver (sdfb_trim_memory)."""
    _lib.check(_lib.lib().sdfb_trim_memory())


def launch_count() -> int:
    """Kernels launched by libsdfb in this process so far."""
    return int(_lib.lib().sdfb_launch_count())


def _validate(vertices, triangles, dx, nx, ny, nz, backend):
    # same checks, order and messages as python/sdfgen_py.cpp:171-182,195-202
    v = np.asarray(vertices)
    t = np.asarray(triangles)
    if v.ndim != 2 or v.shape[1] != 3 or t.ndim != 2 or t.shape[1] != 3:
        raise TypeError("vertices must have shape (N, 3) and triangles shape (M, 3)")
    if v.shape[0] == 0 or t.shape[0] == 0:
        raise ValueError("Cannot generate SDF from empty mesh (vertices or triangles are empty)")
    if nx <= 0 or ny <= 0 or nz <= 0:
        raise ValueError("Grid dimensions must be positive (nx, ny, nz > 0)")
    if not dx > 0.0:
        raise ValueError("Cell spacing dx must be positive")
    if backend == "cpu":
        raise ValueError("backend 'cpu' is not available: sdfgen_b200 is the GPU path only and has no CPU fallback")
    if backend not in ("auto", "gpu"):
        raise ValueError("Invalid backend: " + str(backend) + " (must be 'auto', 'cpu', or 'gpu')")
    # nanobind converts int32/float64 inputs implicitly (python/tests/test_sdfgen.py:770-800)
    v = np.ascontiguousarray(v, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.uint32)
    return v, t


def _origin3(origin):
    o = np.asarray([float(origin[0]), float(origin[1]), float(origin[2])], dtype=np.float32)
    return o


def generate_sdf(vertices, triangles, origin, dx, nx, ny, nz, exact_band: int = 1,
                 backend: str = "auto", num_threads: int = 0, num_gpus: int = 1) -> np.ndarray:
    """Signed distance field of a triangle mesh on an nx x ny x nz grid.

    Same contract as ``sdfgen_ext.generate_sdf``: float32 ``vertices`` (N,3), uint32 ``triangles``
    (M,3), ``origin`` 3-tuple, returns float32 array of shape (nx, ny, nz), C-contiguous.
    ``num_threads`` is accepted and ignored (it only affects the reference's CPU backend).
    ``num_gpus`` (not in the reference, which is single-device): cut the grid into k-slabs over this many GPUs of the
    process (0 = all); the result is bit-identical to one GPU (sdfb_make_level_set3_multi).
    """
    v, t = _validate(vertices, triangles, dx, int(nx), int(ny), int(nz), backend)
    o = _origin3(origin)
    phi = np.empty((int(nx), int(ny), int(nz)), dtype=np.float32)
    if int(num_gpus) == 1:
        rc = _lib.lib().sdfb_make_level_set3(t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0], o.ctypes.data,
                                             float(dx), int(nx), int(ny), int(nz), int(exact_band),
                                             phi.ctypes.data, None, None, _lib.OUT_KFASTEST)
    else:
        rc = _lib.lib().sdfb_make_level_set3_multi(t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0], o.ctypes.data,
                                                   float(dx), int(nx), int(ny), int(nz), int(exact_band),
                                                   phi.ctypes.data, None, None, int(num_gpus), _lib.OUT_KFASTEST)
    _lib.check(rc)
    return phi


def generate_sdf_batch(items, concurrency: int = 4):
    """Many independent ``generate_sdf`` problems in one call (sdfb_make_level_set3_batch): ``items`` is a sequence of
    dicts with the keys of generate_sdf's arguments (vertices, triangles, origin, dx, nx, ny, nz and optionally
    exact_band); up to ``concurrency`` of them run at a time on separate streams.  Returns the list of (nx, ny, nz)
    float32 arrays, identical to what generate_sdf returns for each item."""
    items = list(items)
    arr = (_lib.BatchItem * max(len(items), 1))()
    keep, outs = [], []
    for b, it in zip(arr, items):
        nx, ny, nz = int(it["nx"]), int(it["ny"]), int(it["nz"])
        v, t = _validate(it["vertices"], it["triangles"], it["dx"], nx, ny, nz, it.get("backend", "auto"))
        phi = np.empty((nx, ny, nz), dtype=np.float32)
        keep.append((v, t))
        outs.append(phi)
        b.tri, b.ntri, b.xyz, b.nvert = t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0]
        b.origin[:] = [float(x) for x in _origin3(it["origin"])]
        b.dx, b.ni, b.nj, b.nk, b.exact_band = float(it["dx"]), nx, ny, nz, int(it.get("exact_band", 1))
        b.phi_out, b.status = phi.ctypes.data, 0
    _lib.check(_lib.lib().sdfb_make_level_set3_batch(arr, len(items), int(concurrency), _lib.OUT_KFASTEST))
    return outs


def generate_sdf_file(vertices, triangles, origin, dx, nx, ny, nz, filename: str, exact_band: int = 1) -> int:
    """Mesh -> binary .sdf file without the grid ever becoming a host array: what the reference's CLI does with
    make_level_set3 + write_sdf_binary (app/main.cpp:273,336), with the k-fastest layout and the inside count
    produced on the device (sdfb_plan_write_sdf).  Returns the inside count (cells with phi < 0)."""
    v, t = _validate(vertices, triangles, dx, int(nx), int(ny), int(nz), "gpu")
    o = _origin3(origin)
    plan = _lib.Plan(int(nx), int(ny), int(nz), flags=_lib.OUT_KFASTEST)
    try:
        plan.set_mesh_host(v, t)
        plan.run(o, float(dx), int(exact_band))
        return plan.write_sdf(filename, o, float(dx))
    finally:
        plan.close()


def generate_sdf_debug(vertices, triangles, origin, dx, nx, ny, nz, exact_band: int = 1, flags: int = 0, num_gpus: int = 1):
    """Like generate_sdf but returns ``(phi, closest_tri, intersection_count)`` as FLAT arrays in the
    reference's internal i-fastest order (index i + nx*(j + ny*k)); used by the parity tests, which
    are graded on closest_tri and intersection_count as well (the reference keeps them as locals)."""
    v, t = _validate(vertices, triangles, dx, int(nx), int(ny), int(nz), "gpu")
    o = _origin3(origin)
    V = int(nx) * int(ny) * int(nz)
    phi, tri, cnt = np.empty(V, np.float32), np.empty(V, np.int32), np.empty(V, np.int32)
    a = (t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0], o.ctypes.data, float(dx), int(nx), int(ny), int(nz),
         int(exact_band), phi.ctypes.data, tri.ctypes.data, cnt.ctypes.data)
    if int(num_gpus) == 1:
        rc = _lib.lib().sdfb_make_level_set3(*a, int(flags))
    else:
        rc = _lib.lib().sdfb_make_level_set3_multi(*a, int(num_gpus), int(flags))
    _lib.check(rc)
    return phi, tri, cnt


def _size_grid(extents, nx, ny, nz, dx):
    """The reference's grid sizing rules in one place (python/sdfgen.py:101-112 and :219-247 state them twice):
    without dx the spacing follows from the cell counts -- x extent / nx when a count is missing (proportional
    sizing), the largest extent / count ratio when all three are given -- and every missing count is
    ceil(extent / dx)."""
    counts = (nx, ny, nz)
    if dx is None:
        if nx is None:
            raise ValueError("Must specify either 'dx' or 'nx' (or 'nx', 'ny', 'nz') for grid sizing")
        dx = extents[0] / nx if None in counts[1:] else max(e / n for e, n in zip(extents, counts))
    return [int(np.ceil(e / dx)) if n is None else n for e, n in zip(extents, counts)], dx


def _padded_sdf(vertices, triangles, min_box, max_box, nx, ny, nz, dx, padding, exact_band, backend, num_threads):
    """Size the grid around the box, add `padding` cells on every side (python/sdfgen.py:114-121,249-255), run
    generate_sdf and describe the grid with the reference's metadata keys (:137-142)."""
    (nx, ny, nz), dx = _size_grid(max_box - min_box, nx, ny, nz, dx)
    origin = min_box - padding * dx
    sdf = generate_sdf(vertices, triangles, tuple(origin), dx, nx + 2 * padding, ny + 2 * padding, nz + 2 * padding,
                       exact_band=exact_band, backend=backend, num_threads=num_threads)
    return sdf, {"origin": tuple(origin), "dx": dx, "bounds": (tuple(min_box), tuple(max_box)), "backend": backend}


def generate_from_mesh(vertices: np.ndarray, triangles: np.ndarray, nx: int, ny: Optional[int] = None,
                       nz: Optional[int] = None, dx: Optional[float] = None, padding: int = 1,
                       exact_band: int = 1, backend: str = "auto", num_threads: int = 0) -> Tuple[np.ndarray, dict]:
    """SDF of an in-memory mesh on a grid sized from its bounding box, python/sdfgen.py:47-142 (same arguments,
    same sizing, same metadata keys)."""
    vertices = np.asarray(vertices)
    return _padded_sdf(vertices, triangles, vertices.min(axis=0), vertices.max(axis=0), nx, ny, nz, dx,
                       padding, exact_band, backend, num_threads)


def generate_from_file(filename: str, nx: Optional[int] = None, ny: Optional[int] = None, nz: Optional[int] = None,
                       dx: Optional[float] = None, padding: int = 1, exact_band: int = 1, backend: str = "auto",
                       num_threads: int = 0) -> Tuple[np.ndarray, dict]:
    """Load a mesh file and generate its SDF, python/sdfgen.py:145-265 (same sizing modes; the box is the loader's,
    as float32)."""
    vertices, triangles, bounds = load_mesh(filename)
    return _padded_sdf(vertices, triangles, np.array(bounds[0], dtype=np.float32), np.array(bounds[1], dtype=np.float32),
                       nx, ny, nz, dx, padding, exact_band, backend, num_threads)


def load_mesh(filename: str):
    from .mesh_io import load_mesh as _load
    return _load(filename)


def save_sdf(filename: str, sdf_array, origin, dx) -> None:
    from .mesh_io import save_sdf as _save
    _save(filename, sdf_array, origin, dx)


def load_sdf(filename: str):
    from .mesh_io import load_sdf as _load
    return _load(filename)
