"""CPU stand-in for sdfgen_b200.dist.CudaSlabEngine used by the gloo tests: same interface, arithmetic by
the oracle (band/counts per slab, the column-schedule emulator for sweeps).  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np
import torch

import oracle


class OracleSlabEngine:
    def __init__(self, vertices, triangles, ni, nj, nk, k_lo, k_hi, relax_from=None):
        # relax_from: sweeps >= this index use the relaxation emulator (seeded random order) instead of the column one
        self.relax_from = relax_from
        self.v = np.ascontiguousarray(vertices, np.float32)
        self.t = np.ascontiguousarray(triangles, np.uint32)
        self.ni, self.nj, self.nk, self.k_lo, self.k_hi = ni, nj, nk, k_lo, k_hi
        self.plane, self.nkl = ni * nj, k_hi - k_lo
        n = self.plane * (self.nkl + 2)
        self.cphi = np.zeros(n, np.float32)
        self.clo = np.zeros(n, np.uint32)
        self.counts = None
        self.phi = None
        self._changed = 0
        # staging tensors for the exchange: int64 = (phi bits << 32 | lo), like the device cells
        self.stage = {k: torch.zeros(self.plane, dtype=torch.int64) for k in ("lo_s", "hi_s", "lo_r", "hi_r")}

    def band(self, origin, dx, exact_band):
        self.o, self.dx = np.ascontiguousarray(origin, np.float32), float(dx)
        phi, tri, cnt = oracle.port.band_counts_slab(self.v, self.t, self.o, self.dx, self.ni, self.nj, self.nk,
                                                     self.k_lo, self.k_hi, exact_band)
        init = np.float32(np.float32(self.ni + self.nj + self.nk) * np.float32(self.dx))
        self.cphi[:] = init
        self.clo[:] = 0xFFFFFFFF
        p = self.plane
        self.cphi[p:p * (self.nkl + 1)] = phi
        self.clo[p:p * (self.nkl + 1)] = np.where(tri < 0, np.uint32(0xFFFFFFFF), tri.astype(np.uint32))
        self.counts = cnt
        self._changed = 0

    def sweep(self, first, count):
        L = oracle.port.lib()
        for s in range(first, first + count):
            ch = C.c_long()
            if self.relax_from is not None and s >= self.relax_from:
                rd = C.c_long()
                L.sdfo_emu_sweep_relax(self.t, self.v, self.cphi, self.clo, self.o, self.dx, self.ni, self.nj, self.nk,
                                       self.k_lo, self.k_hi, s, 7 + s, C.byref(ch), C.byref(rd))
            else:
                L.sdfo_emu_sweep_columns(self.t, self.v, self.cphi, self.clo, self.o, self.dx, self.ni, self.nj, self.nk,
                                         self.k_lo, self.k_hi, s, C.byref(ch))
            self._changed += int(ch.value)

    def changed(self):
        c, self._changed = self._changed, 0
        return c

    def _pack(self, sl):
        return torch.from_numpy((self.cphi[sl].view(np.uint32).astype(np.int64) << 32) | self.clo[sl].astype(np.int64))

    def boundary_planes(self):
        p = self.plane
        self.stage["lo_s"].copy_(self._pack(slice(p, 2 * p)))
        self.stage["hi_s"].copy_(self._pack(slice(self.nkl * p, (self.nkl + 1) * p)))
        return self.stage["lo_s"], self.stage["hi_s"]

    def halo_planes(self):
        return self.stage["lo_r"], self.stage["hi_r"]

    def halo_refresh(self):
        p = self.plane
        for key, sl, present in (("lo_r", slice(0, p), self.k_lo > 0),
                                 ("hi_r", slice((self.nkl + 1) * p, (self.nkl + 2) * p), self.k_hi < self.nk)):
            if not present:
                continue
            x = self.stage[key].numpy()
            lo = (x & 0xFFFFFFFF).astype(np.uint32)
            lo = np.where((lo & 0x07FFFFFF) != 0x07FFFFFF, lo | np.uint32(31 << 27), lo)
            self.cphi[sl] = ((x >> 32) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
            self.clo[sl] = lo

    def halo_commit(self, lower):
        """Exact mode: unpack one received plane as it is (stamps kept)."""
        p = self.plane
        key, sl = ("lo_r", slice(0, p)) if lower else ("hi_r", slice((self.nkl + 1) * p, (self.nkl + 2) * p))
        x = self.stage[key].numpy()
        self.cphi[sl] = ((x >> 32) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
        self.clo[sl] = (x & 0xFFFFFFFF).astype(np.uint32)

    def counter_tensor(self, value):
        return torch.tensor([value], dtype=torch.int64)

    def sign(self):
        p = self.plane
        self.phi = self.cphi[p:p * (self.nkl + 1)].copy()
        oracle.port.lib().sdfo_apply_sign.argtypes = [C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.int32),
                                                      np.ctypeslib.ndpointer(np.float32)]
        oracle.port.lib().sdfo_apply_sign.restype = None
        oracle.port.lib().sdfo_apply_sign(self.ni, self.nj, self.nkl, self.counts, self.phi)

    def tri(self):
        p = self.plane
        lo = self.clo[p:p * (self.nkl + 1)]
        return np.where((lo & 0x07FFFFFF) == 0x07FFFFFF, -1, (lo & 0x07FFFFFF).astype(np.int64)).astype(np.int32)
