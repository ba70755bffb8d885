"""ctypes doors to the parity checkers (TEST INFRASTRUCTURE ONLY -- see oracle/sdf_oracle.c header).

``oracle.port``  -> oracle/liboracle.so, the plain-C restatement (always buildable: gcc only).
``oracle.ref``   -> oracle/_ref/libsdfgen_ref.so, the unmodified reference CPU code compiled in place
                    from /root/reference (present wherever it was built; it travels to the GPU box
                    as a prebuilt file).  ``oracle.have_ref()`` says whether it can be loaded.

Only tests/, __graft_entry__ (build/smoke) and bench.py's CPU-baseline legs may import this package.
Nothing under sdfgen_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libsdfgen_ref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """(Re)build the checkers with oracle/Makefile (the _ref target is a no-op without /root/reference)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class _Outputs(C.Structure):
    _fields_ = [("phi_band", C.POINTER(C.c_float)), ("tri_band", C.POINTER(C.c_int32)),
                ("counts", C.POINTER(C.c_int32)), ("phi_swept", C.POINTER(C.c_float)),
                ("tri_final", C.POINTER(C.c_int32)), ("stats", C.POINTER(C.c_int64))]


@dataclass
class Staged:
    """All arrays are flat, i fastest (index i + ni*(j + nj*k)), like the reference's Array3."""
    phi: np.ndarray          # signed result
    phi_band: np.ndarray     # |phi| after the exact band
    tri_band: np.ndarray     # closest_tri after the exact band
    counts: np.ndarray       # intersection_count
    phi_swept: np.ndarray    # |phi| after the sweeps
    tri_final: np.ndarray    # closest_tri after the sweeps
    stats: np.ndarray | None = None


def _prep(vertices, triangles, origin):
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    o = np.ascontiguousarray(origin, dtype=np.float32).reshape(3)
    return v, t, o


class _Port:
    _lib = None

    def lib(self):
        if self._lib is None:
            srcs = [os.path.join(_HERE, f) for f in ("sdf_oracle.c", "columns_emu.c")]
            if not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < max(os.path.getmtime(f) for f in srcs):
                build()
            L = C.CDLL(_PORT_SO)
            L.sdfo_make_level_set3.restype = C.c_int
            L.sdfo_make_level_set3.argtypes = [_u32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_float,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p,
                                               C.POINTER(_Outputs)]
            L.sdfo_band_counts_slab.restype = C.c_int
            L.sdfo_band_counts_slab.argtypes = [_u32p, C.c_uint64, _f32p, _f32p, C.c_float,
                                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                _f32p, np.ctypeslib.ndpointer(np.int32), np.ctypeslib.ndpointer(np.int32)]
            L.sdfo_point_triangle_distance.restype = C.c_float
            L.sdfo_point_triangle_distance.argtypes = [_f32p, _f32p, _f32p, _f32p]
            L.sdfo_emu_sweep_columns.restype = C.c_long
            L.sdfo_emu_sweep_columns.argtypes = [_u32p, _f32p, _f32p, _u32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
            L.sdfo_emu_flag_violations.restype = C.c_long
            L.sdfo_emu_sweep_relax.restype = C.c_long
            L.sdfo_emu_sweep_relax.argtypes = [_u32p, _f32p, _f32p, _u32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_long), C.POINTER(C.c_long)]
            _u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
            L.sdfo_emu_sweep_relax_from.restype = C.c_long
            L.sdfo_emu_sweep_relax_from.argtypes = [_u32p, _f32p, _f32p, _u32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                                    C.c_int, C.c_int, C.c_int, C.c_uint64, _u8p, _u8p,
                                                    C.POINTER(C.c_long), C.POINTER(C.c_long)]
            L.sdfo_emu_look_scan.restype = C.c_long
            L.sdfo_emu_look_scan.argtypes = [_u32p, _f32p, _f32p, _u32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, _u8p]
            L.sdfo_emu_look_mark.restype = C.c_long
            L.sdfo_emu_look_mark.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
            self._lib = L
        return self._lib

    def emu_sweep_lookahead(self, vertices, triangles, origin, dx, ni, nj, nk, phi_band, tri_band, nsweeps=16,
                            look_from=8, window=8, dedupe=1, seed=1):
        """CPU emulation of the production mix WITH the lookahead window of sdfb_sweep_relax.cu: column emulation below
        `look_from`; from there on windows of `window` sweeps -- one scan of the cells as they are before the window
        (sdfo_emu_look_scan), then every sweep's relaxation starts from its marks plus the cells the window's earlier
        sweeps changed and their downstream neighbours (sdfo_emu_look_mark) instead of from every voxel.
        Returns (phi_swept, tri_final, evals_per_sweep, scan_evals, round0_cells_per_sweep)."""
        v, t, o = _prep(vertices, triangles, origin)
        plane = ni * nj
        ncell = plane * (nk + 2)
        init = np.float32(np.float32(ni + nj + nk) * np.float32(dx))
        cphi = np.full(ncell, init, np.float32)
        clo = np.full(ncell, 0xFFFFFFFF, np.uint32)
        cphi[plane:plane * (nk + 1)] = phi_band
        tb = np.asarray(tri_band)
        clo[plane:plane * (nk + 1)] = np.where(tb < 0, np.uint32(0xFFFFFFFF), tb.astype(np.uint32))
        evals, scan_evals, r0 = [], [], []
        L = self.lib()
        marks = log = None
        w_lo = w_hi = -1
        for s in range(nsweeps):
            ch, rd = C.c_long(), C.c_long()
            if s < look_from:
                e = L.sdfo_emu_sweep_columns(t, v, cphi, clo, o, dx, ni, nj, nk, 0, nk, s, C.byref(ch))
                r0.append(-1)
            else:
                if not (w_lo <= s < w_hi):
                    w_lo, w_hi = s, min(s + window, nsweeps)
                    marks = np.zeros((w_hi - w_lo) * ncell, np.uint8)
                    log = np.zeros(ncell, np.uint8)
                    se = L.sdfo_emu_look_scan(t, v, cphi, clo, o, dx, ni, nj, nk, w_lo, w_hi, dedupe, marks)
                    assert se >= 0
                    scan_evals.append(int(se))
                round0 = np.zeros(ncell, np.uint8)
                n0 = L.sdfo_emu_look_mark(marks[(s - w_lo) * ncell:(s - w_lo + 1) * ncell], log, ni, nj, nk, s, round0)
                r0.append(int(n0))
                e = L.sdfo_emu_sweep_relax_from(t, v, cphi, clo, o, dx, ni, nj, nk, 0, nk, s, seed + s, round0, log,
                                                C.byref(ch), C.byref(rd))
            evals.append(int(e))
        lo = clo[plane:plane * (nk + 1)]
        tri = np.where((lo & 0x07FFFFFF) == 0x07FFFFFF, -1, (lo & 0x07FFFFFF).astype(np.int64)).astype(np.int32)
        return cphi[plane:plane * (nk + 1)].copy(), tri, evals, scan_evals, r0

    def emu_sweep_columns(self, vertices, triangles, origin, dx, ni, nj, nk, phi_band, tri_band, nsweeps=16,
                          k_lo=0, k_hi=None, shape=(8, 16)):
        """CPU emulation of the CUDA column schedule (oracle/columns_emu.c) starting from band results; shape = the
        column cross-section (the library builds 8 x 16 and 8 x 12).
        Returns (phi_swept, tri_final, evals_per_sweep, changed_per_sweep, flag_violations)."""
        v, t, o = _prep(vertices, triangles, origin)
        self.lib().sdfo_emu_set_column_shape(int(shape[0]), int(shape[1]))
        k_hi = nk if k_hi is None else k_hi
        plane = ni * nj
        nkl = k_hi - k_lo
        init = np.float32(np.float32(ni + nj + nk) * np.float32(dx))
        cphi = np.full(plane * (nkl + 2), init, np.float32)
        clo = np.full(plane * (nkl + 2), 0xFFFFFFFF, np.uint32)
        cphi[plane:plane * (nkl + 1)] = phi_band
        tb = np.asarray(tri_band)
        clo[plane:plane * (nkl + 1)] = np.where(tb < 0, np.uint32(0xFFFFFFFF), tb.astype(np.uint32))
        evals, changed = [], []
        for s in range(nsweeps):
            ch = C.c_long()
            e = self.lib().sdfo_emu_sweep_columns(t, v, cphi, clo, o, dx, ni, nj, nk, k_lo, k_hi, s, C.byref(ch))
            evals.append(int(e)); changed.append(int(ch.value))
        lo = clo[plane:plane * (nkl + 1)]
        tri = np.where((lo & 0x07FFFFFF) == 0x07FFFFFF, -1, (lo & 0x07FFFFFF).astype(np.int64)).astype(np.int32)
        self.lib().sdfo_emu_set_column_shape(8, 16)
        return cphi[plane:plane * (nkl + 1)].copy(), tri, evals, changed, int(self.lib().sdfo_emu_flag_violations())

    def emu_sweep_mixed(self, vertices, triangles, origin, dx, ni, nj, nk, phi_band, tri_band, nsweeps=16, relax_from=8, seed=1):
        """CPU emulation of the production schedule mix: column emulation for sweeps < relax_from, relaxation
        emulation (oracle/relax_emu.c, seeded random order) from there on.
        Returns (phi_swept, tri_final, evals_per_sweep, changed_per_sweep, rounds_per_sweep)."""
        v, t, o = _prep(vertices, triangles, origin)
        plane = ni * nj
        init = np.float32(np.float32(ni + nj + nk) * np.float32(dx))
        cphi = np.full(plane * (nk + 2), init, np.float32)
        clo = np.full(plane * (nk + 2), 0xFFFFFFFF, np.uint32)
        cphi[plane:plane * (nk + 1)] = phi_band
        tb = np.asarray(tri_band)
        clo[plane:plane * (nk + 1)] = np.where(tb < 0, np.uint32(0xFFFFFFFF), tb.astype(np.uint32))
        evals, changed, rounds = [], [], []
        for s in range(nsweeps):
            ch, rd = C.c_long(), C.c_long()
            if s < relax_from:
                e = self.lib().sdfo_emu_sweep_columns(t, v, cphi, clo, o, dx, ni, nj, nk, 0, nk, s, C.byref(ch))
            else:
                e = self.lib().sdfo_emu_sweep_relax(t, v, cphi, clo, o, dx, ni, nj, nk, 0, nk, s, seed + s, C.byref(ch), C.byref(rd))
            evals.append(int(e)); changed.append(int(ch.value)); rounds.append(int(rd.value))
        lo = clo[plane:plane * (nk + 1)]
        tri = np.where((lo & 0x07FFFFFF) == 0x07FFFFFF, -1, (lo & 0x07FFFFFF).astype(np.int64)).astype(np.int32)
        return cphi[plane:plane * (nk + 1)].copy(), tri, evals, changed, rounds

    def make_level_set3(self, vertices, triangles, origin, dx, ni, nj, nk, exact_band=1):
        """Signed phi only (flat, i fastest)."""
        v, t, o = _prep(vertices, triangles, origin)
        phi = np.empty(ni * nj * nk, dtype=np.float32)
        rc = self.lib().sdfo_make_level_set3(t, t.shape[0], v, v.shape[0], o, dx, ni, nj, nk, exact_band,
                                             16, phi, None)
        if rc != 0:
            raise RuntimeError(f"oracle port failed rc={rc}")
        return phi

    def staged(self, vertices, triangles, origin, dx, ni, nj, nk, exact_band=1, nsweeps=16, stats=False) -> Staged:
        v, t, o = _prep(vertices, triangles, origin)
        V = ni * nj * nk
        s = Staged(np.empty(V, np.float32), np.empty(V, np.float32), np.empty(V, np.int32),
                   np.empty(V, np.int32), np.empty(V, np.float32), np.empty(V, np.int32),
                   np.zeros(80, np.int64) if stats else None)
        out = _Outputs(_ptr(s.phi_band, C.c_float), _ptr(s.tri_band, C.c_int32), _ptr(s.counts, C.c_int32),
                       _ptr(s.phi_swept, C.c_float), _ptr(s.tri_final, C.c_int32), _ptr(s.stats, C.c_int64))
        rc = self.lib().sdfo_make_level_set3(t, t.shape[0], v, v.shape[0], o, dx, ni, nj, nk, exact_band,
                                             nsweeps, s.phi, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"oracle port failed rc={rc}")
        return s

    def band_counts_slab(self, vertices, triangles, origin, dx, ni, nj, nk, k_lo, k_hi, exact_band=1):
        v, t, o = _prep(vertices, triangles, origin)
        V = ni * nj * (k_hi - k_lo)
        phi, tri, cnt = np.empty(V, np.float32), np.empty(V, np.int32), np.empty(V, np.int32)
        rc = self.lib().sdfo_band_counts_slab(t, t.shape[0], v, o, dx, ni, nj, nk, k_lo, k_hi, exact_band,
                                              phi, tri, cnt)
        if rc != 0:
            raise RuntimeError(f"oracle slab failed rc={rc}")
        return phi, tri, cnt

    def point_triangle_distance(self, x0, x1, x2, x3) -> float:
        a = [np.ascontiguousarray(x, dtype=np.float32).reshape(3) for x in (x0, x1, x2, x3)]
        return float(self.lib().sdfo_point_triangle_distance(*a))


class _Ref:
    _lib = None

    def available(self) -> bool:
        return os.path.exists(_REF_SO)

    def lib(self):
        if self._lib is None:
            if not self.available():
                build()
            if not self.available():
                raise RuntimeError("oracle/_ref/libsdfgen_ref.so is absent and /root/reference is not here to build it")
            L = C.CDLL(_REF_SO)
            L.ref_make_level_set3.restype = C.c_int
            L.ref_make_level_set3.argtypes = [_u32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_float,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
            L.ref_make_level_set3_staged.restype = C.c_int
            L.ref_make_level_set3_staged.argtypes = [_u32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_float,
                                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
            L.ref_point_triangle_distance.restype = C.c_float
            L.ref_point_triangle_distance.argtypes = [_f32p, _f32p, _f32p, _f32p]
            L.ref_hardware_concurrency.restype = C.c_int
            self._lib = L
        return self._lib

    def hardware_concurrency(self) -> int:
        return int(self.lib().ref_hardware_concurrency())

    def make_level_set3(self, vertices, triangles, origin, dx, ni, nj, nk, exact_band=1, num_threads=1):
        """sdfgen::cpu::make_level_set3 itself; num_threads=1 is the parity oracle (the threaded
        path is racy), num_threads=0 is the reference's auto-threaded timing baseline."""
        v, t, o = _prep(vertices, triangles, origin)
        phi = np.empty(ni * nj * nk, dtype=np.float32)
        rc = self.lib().ref_make_level_set3(t, t.shape[0], v, v.shape[0], o, dx, ni, nj, nk, exact_band,
                                            num_threads, phi)
        if rc != 0:
            raise RuntimeError(f"reference failed rc={rc}")
        return phi

    def staged(self, vertices, triangles, origin, dx, ni, nj, nk, exact_band=1, nsweeps=16) -> Staged:
        v, t, o = _prep(vertices, triangles, origin)
        V = ni * nj * nk
        s = Staged(np.empty(V, np.float32), np.empty(V, np.float32), np.empty(V, np.int32),
                   np.empty(V, np.int32), np.empty(V, np.float32), np.empty(V, np.int32))
        rc = self.lib().ref_make_level_set3_staged(
            t, t.shape[0], v, v.shape[0], o, dx, ni, nj, nk, exact_band, nsweeps,
            s.phi_band.ctypes.data, s.tri_band.ctypes.data, s.counts.ctypes.data,
            s.phi_swept.ctypes.data, s.tri_final.ctypes.data, s.phi.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"reference staged failed rc={rc}")
        return s

    def point_triangle_distance(self, x0, x1, x2, x3) -> float:
        a = [np.ascontiguousarray(x, dtype=np.float32).reshape(3) for x in (x0, x1, x2, x3)]
        return float(self.lib().ref_point_triangle_distance(*a))


class _RefGpu:
    """The reference's OWN CUDA file (gpu_lib/makelevelset3_gpu.cu), unmodified, recompiled in place for sm_100a
    (oracle/_ref/libsdfgen_refgpu.so).  It computes a different far field (Jacobi Eikonal) -- a TIMING comparator for
    bench.py ("the existing GPU kernel"), never a parity oracle."""
    _lib = None
    _SO = os.path.join(_HERE, "_ref", "libsdfgen_refgpu.so")

    def available(self) -> bool:
        return os.path.exists(self._SO)

    def make_level_set3(self, vertices, triangles, origin, dx, ni, nj, nk, exact_band=1, want_phi=True):
        """Returns (phi or None, seconds of the whole call: 8 cudaMalloc, H2D, kernels, D2H, 8 cudaFree)."""
        if self._lib is None:
            L = C.CDLL(self._SO)
            L.sdfref_gpu_make_level_set3.restype = C.c_int
            L.sdfref_gpu_make_level_set3.argtypes = [_u32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_float, C.c_int, C.c_int,
                                                     C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
            self._lib = L
        v, t, o = _prep(vertices, triangles, origin)
        phi = np.empty(ni * nj * nk, np.float32) if want_phi else None
        sec = C.c_double()
        rc = self._lib.sdfref_gpu_make_level_set3(t, t.shape[0], v, v.shape[0], o, dx, ni, nj, nk, exact_band,
                                                  phi.ctypes.data if want_phi else None, C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"reference GPU path failed rc={rc}")
        return phi, float(sec.value)


port = _Port()
ref = _Ref()
refgpu = _RefGpu()


def have_ref() -> bool:
    return ref.available() or os.path.exists("/root/reference/cpu_lib/makelevelset3.cpp")


def best():
    """The strongest available checker: the compiled reference if present, else the C port."""
    return ref if have_ref() else port
