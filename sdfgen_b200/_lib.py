"""ctypes binding of libsdfb.so (include/sdfb.h).  No fallback: a missing library is an ImportError
that says how to build it, a missing GPU is a RuntimeError from the library itself."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDFB_LIB_PATH") or os.path.join(_HERE, "libsdfb.so")   # override: development builds only

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_LIMIT, ERR_IO = 0, -1, -2, -3, -4, -5, -6, -7
OUT_KFASTEST, SWEEP_LEVELS, NO_SIGN, SWEEP_RELAX, SWEEP_COLUMNS = 0x1, 0x2, 0x4, 0x10, 0x20
LINK_HANDLE_BYTES = 128

_lib = None


class BatchItem(C.Structure):
    """sdfb_batch_item (include/sdfb.h)."""
    _fields_ = [("tri", C.c_void_p), ("ntri", C.c_uint64), ("xyz", C.c_void_p), ("nvert", C.c_uint64),
                ("origin", C.c_float * 3), ("dx", C.c_float),
                ("ni", C.c_int32), ("nj", C.c_int32), ("nk", C.c_int32), ("exact_band", C.c_int32),
                ("phi_out", C.c_void_p), ("status", C.c_int32)]


class SdfbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsdfb error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build the CUDA extension first: "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C sdfgen_b200/csrc`. "
            "sdfgen_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32, u32, f32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32, C.c_float
    L.sdfb_version.restype = C.c_char_p
    L.sdfb_last_error.restype = C.c_char_p
    L.sdfb_device_count.restype = C.c_int
    L.sdfb_launch_count.restype = u64
    L.sdfb_make_level_set3.restype = C.c_int
    L.sdfb_make_level_set3.argtypes = [vp, u64, vp, u64, vp, f32, i32, i32, i32, i32, vp, vp, vp, u32]
    L.sdfb_plan_create.restype = C.c_int
    L.sdfb_plan_create.argtypes = [C.POINTER(vp), C.c_int, i32, i32, i32, i32, i32, u32]
    L.sdfb_plan_destroy.restype = C.c_int
    L.sdfb_plan_destroy.argtypes = [vp]
    for name in ("sdfb_plan_set_mesh_host", "sdfb_plan_set_mesh_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, u64, vp, u64, vp]
    L.sdfb_plan_band.restype = C.c_int
    L.sdfb_plan_band.argtypes = [vp, vp, f32, i32, vp]
    L.sdfb_plan_sweep.restype = C.c_int
    L.sdfb_plan_sweep.argtypes = [vp, i32, i32, vp]
    L.sdfb_plan_sign.restype = C.c_int
    L.sdfb_plan_sign.argtypes = [vp, vp]
    L.sdfb_plan_run.restype = C.c_int
    L.sdfb_plan_run.argtypes = [vp, vp, f32, i32, vp]
    L.sdfb_plan_device_ptrs.restype = C.c_int
    L.sdfb_plan_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sdfb_plan_changed.restype = C.c_int
    L.sdfb_plan_changed.argtypes = [vp, vp, C.POINTER(u64)]
    L.sdfb_plan_halo_refresh.restype = C.c_int
    L.sdfb_plan_halo_refresh.argtypes = [vp, vp]
    L.sdfb_plan_counters.restype = C.c_int
    L.sdfb_plan_counters.argtypes = [vp, vp, C.POINTER(u64 * 2)]
    L.sdfb_plan_download.restype = C.c_int
    L.sdfb_plan_download.argtypes = [vp, vp, vp, vp, vp]
    L.sdfb_plan_download_phi_async.restype = C.c_int
    L.sdfb_plan_download_phi_async.argtypes = [vp, vp, vp]
    L.sdfb_plan_set_concurrency.restype = C.c_int
    L.sdfb_plan_set_concurrency.argtypes = [vp, i32]
    L.sdfb_trim_memory.restype = C.c_int
    L.sdfb_trim_memory.argtypes = []
    L.sdfb_make_level_set3_batch.restype = C.c_int
    L.sdfb_make_level_set3_batch.argtypes = [C.POINTER(BatchItem), i32, i32, u32]
    L.sdfb_plan_write_sdf.restype = C.c_int
    L.sdfb_plan_write_sdf.argtypes = [vp, C.c_char_p, vp, f32, C.POINTER(C.c_int64), vp]
    L.sdfb_plan_phase_ms.restype = C.c_int
    L.sdfb_plan_phase_ms.argtypes = [vp, C.POINTER(f32 * 4)]
    L.sdfb_plan_verify.restype = C.c_int
    L.sdfb_plan_verify.argtypes = [vp, vp, C.POINTER(u64 * 4)]
    L.sdfb_slab_bounds.restype = C.c_int
    L.sdfb_slab_bounds.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.sdfb_plan_link_export.restype = C.c_int
    L.sdfb_plan_link_export.argtypes = [vp, vp]
    L.sdfb_plan_link_import.restype = C.c_int
    L.sdfb_plan_link_import.argtypes = [vp, i32, vp]
    L.sdfb_plan_link_trace.restype = C.c_int
    L.sdfb_plan_link_trace.argtypes = [vp, vp, C.POINTER(u64 * 32)]
    L.sdfb_plan_unlink.restype = C.c_int
    L.sdfb_plan_unlink.argtypes = [vp]
    L.sdfb_plan_download_global.restype = C.c_int
    L.sdfb_plan_download_global.argtypes = [vp, vp, vp, vp, vp]
    L.sdfb_make_level_set3_multi.restype = C.c_int
    L.sdfb_make_level_set3_multi.argtypes = [vp, u64, vp, u64, vp, f32, i32, i32, i32, i32, vp, vp, vp, i32, u32]
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        msg = lib().sdfb_last_error().decode("utf-8", "replace")
        if rc == ERR_INVALID:
            raise ValueError(msg)          # the reference raises std::invalid_argument -> ValueError
        if rc == ERR_OOM:
            raise MemoryError(msg)
        if rc == ERR_IO:
            raise OSError(msg)
        raise SdfbError(rc, msg)


def _addr(a):
    """Host address of a numpy array, or a raw integer address (device pointer / pinned buffer)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


class Plan:
    """Device-resident state for one k-slab of a grid (include/sdfb.h plan API)."""

    def __init__(self, ni, nj, nk, k_lo=0, k_hi=None, device=0, flags=0):
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.k_lo, self.k_hi = int(k_lo), int(nk if k_hi is None else k_hi)
        self.device, self.flags = int(device), int(flags)
        self._h = C.c_void_p()
        check(lib().sdfb_plan_create(C.byref(self._h), self.device, self.ni, self.nj, self.nk,
                                     self.k_lo, self.k_hi, self.flags))
        self._keep = None

    @property
    def slab_voxels(self):
        return self.ni * self.nj * (self.k_hi - self.k_lo)

    def close(self):
        if self._h:
            lib().sdfb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh_host(self, vertices: np.ndarray, triangles: np.ndarray, stream=0):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
        check(lib().sdfb_plan_set_mesh_host(self._h, t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0], stream or None))
        self._keep = (v, t)

    def set_mesh_host_ptr(self, tri_addr, ntri, xyz_addr, nvert, stream=0):
        check(lib().sdfb_plan_set_mesh_host(self._h, tri_addr, ntri, xyz_addr, nvert, stream or None))

    def set_mesh_device(self, d_tri_addr, ntri, d_xyz_addr, nvert, stream=0, keepalive=None):
        check(lib().sdfb_plan_set_mesh_device(self._h, d_tri_addr, ntri, d_xyz_addr, nvert, stream or None))
        self._keep = keepalive

    def _origin(self, origin):
        self._o = np.ascontiguousarray(origin, dtype=np.float32).reshape(3)
        return self._o.ctypes.data

    def band(self, origin, dx, exact_band=1, stream=0):
        check(lib().sdfb_plan_band(self._h, self._origin(origin), float(dx), int(exact_band), stream or None))

    def sweep(self, first=0, count=16, stream=0):
        check(lib().sdfb_plan_sweep(self._h, int(first), int(count), stream or None))

    def sign(self, stream=0):
        check(lib().sdfb_plan_sign(self._h, stream or None))

    def run(self, origin, dx, exact_band=1, stream=0):
        check(lib().sdfb_plan_run(self._h, self._origin(origin), float(dx), int(exact_band), stream or None))

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().sdfb_plan_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def changed(self, stream=0) -> int:
        n = C.c_uint64()
        check(lib().sdfb_plan_changed(self._h, stream or None, C.byref(n)))
        return int(n.value)

    def halo_refresh(self, stream=0):
        check(lib().sdfb_plan_halo_refresh(self._h, stream or None))

    def counters(self, stream=0):
        """(changed cells, distance evaluations) of the sweeps since the last band/counters call."""
        out = (C.c_uint64 * 2)()
        check(lib().sdfb_plan_counters(self._h, stream or None, C.byref(out)))
        return int(out[0]), int(out[1])

    def download(self, phi=True, tri=False, counts=False, stream=0, phi_out=None):
        """Blocking copy to host.  Returns (phi, tri, counts) flat arrays (None where not requested).
        ``phi_out`` may be a preallocated float32 array or a raw (e.g. pinned) host address."""
        V = self.slab_voxels
        p = phi_out if phi_out is not None else (np.empty(V, np.float32) if phi else None)
        t = np.empty(V, np.int32) if tri else None
        c = np.empty(V, np.int32) if counts else None
        check(lib().sdfb_plan_download(self._h, _addr(p), _addr(t), _addr(c), stream or None))
        return p, t, c

    def download_phi_async(self, phi_out, copy_stream):
        """Enqueue the D2H copy of the last sign pass's phi on `copy_stream` (raw cudaStream_t, not the compute
        stream); phi_out is a pinned host address or array.  Synchronise that stream before reading."""
        check(lib().sdfb_plan_download_phi_async(self._h, _addr(phi_out), copy_stream or None))

    def set_concurrency(self, plans_in_flight: int):
        """This many plans run on the device at once (own streams): the sweep kernels take their share of the SMs."""
        check(lib().sdfb_plan_set_concurrency(self._h, int(plans_in_flight)))

    def write_sdf(self, path, min_box, dx, stream=0) -> int:
        """Write the signed phi of the last sign pass as a binary .sdf file straight from the device
        (write_sdf_binary, common/sdf_io.cpp:10-74); returns the inside count."""
        inside = C.c_int64()
        check(lib().sdfb_plan_write_sdf(self._h, os.fsencode(path), self._origin(min_box), float(dx), C.byref(inside), stream or None))
        return int(inside.value)

    # ---- exact multi-GPU mode (include/sdfb.h: linked k-slabs) -------------------------------------------------
    def link_export(self) -> bytes:
        """Allocate this slab's inbound hand-over buffers and return the opaque handle the neighbours import."""
        buf = C.create_string_buffer(LINK_HANDLE_BYTES)
        check(lib().sdfb_plan_link_export(self._h, buf))
        return buf.raw

    def link_import(self, side: int, handle: bytes):
        """Map the inbound buffers of the plan holding the slab below (side 0) or above (side 1)."""
        buf = C.create_string_buffer(bytes(handle), LINK_HANDLE_BYTES)
        check(lib().sdfb_plan_link_import(self._h, int(side), buf))

    def link_trace(self, stream=0):
        """[(start_ns, end_ns)] of the 16 sweeps on this slab since the last call (SDFB_LINK_TRACE=1 at plan creation)."""
        out = (C.c_uint64 * 32)()
        check(lib().sdfb_plan_link_trace(self._h, stream or None, C.byref(out)))
        return [(int(out[2 * s]), int(out[2 * s + 1])) for s in range(16)]

    def unlink(self):
        check(lib().sdfb_plan_unlink(self._h))

    def download_global(self, phi=None, tri=None, counts=None, stream=0):
        """Blocking copy of the slab into arrays of the WHOLE grid (flat, the plan's layout), at the slab's place."""
        check(lib().sdfb_plan_download_global(self._h, _addr(phi), _addr(tri), _addr(counts), stream or None))

    def verify(self, stream=0):
        """Device-side self-consistency check and checksums (sdfb_plan_verify): dict(inconsistent, without_triangle,
        checksum_cells, checksum_values)."""
        out = (C.c_uint64 * 4)()
        check(lib().sdfb_plan_verify(self._h, stream or None, C.byref(out)))
        return dict(inconsistent=int(out[0]), without_triangle=int(out[1]), checksum_cells=int(out[2]), checksum_values=int(out[3]))

    def phase_ms(self):
        out = (C.c_float * 4)()
        check(lib().sdfb_plan_phase_ms(self._h, C.byref(out)))
        return dict(band=out[0], sweeps=out[1], sign=out[2], total=out[3])
