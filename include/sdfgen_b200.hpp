// sdfgen_b200.hpp -- C++ shim: the reference's C++ entry points for the make_level_set3 path, implemented on
// top of the C ABI of sdfb.h (libsdfb.so, sm_100a).  Header-only; meant to be compiled INSIDE the reference
// tree (it includes the reference's own containers by name, nothing is copied):
//
//   sdfgen::gpu::make_level_set3   replaces /root/reference/gpu_lib/makelevelset3_gpu.h:40-42 (same signature)
//   sdfgen::make_level_set3        replaces /root/reference/common/sdfgen_unified.cpp:30-71 (same signature;
//                                  HardwareBackend::CPU is rejected: this build has no CPU path or fallback)
//   sdfgen::is_gpu_available       replaces /root/reference/common/sdfgen_unified.cpp:19-28
//
// Behaviour kept: phi is resized to nx*ny*nz and filled in the reference's i-fastest order
// (common/array3.h:111-115); the call is synchronous on the current device; tri / x are borrowed.
// Behaviour changed on purpose: errors throw std::runtime_error / std::invalid_argument carrying
// sdfb_last_error() instead of exit(EXIT_FAILURE) (gpu_lib/makelevelset3_gpu.cu:14-20); results follow the
// reference's single-threaded CPU semantics bit for bit (the reference's own CUDA file computes a
// different far field).
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "array3.h"   // reference: common/array3.h  (Array3f)
#include "vec.h"      // reference: common/vec.h     (Vec3f, Vec3ui)
#include "sdfb.h"

namespace sdfgen {

#ifndef SDFGEN_B200_HAVE_BACKEND_ENUM
#define SDFGEN_B200_HAVE_BACKEND_ENUM
// same enumerators as common/sdfgen_unified.h:16-20
enum class HardwareBackend { Auto, CPU, GPU };
#endif

inline bool is_gpu_available() { return sdfb_device_count() > 0; }

namespace gpu {

// Not in the reference (it is single-device, README.md:220 lists multi-GPU as future work): number of GPUs of this
// process the grid is cut over, k-slabs with the sweeps' serial order kept across the faces (sdfb_make_level_set3_multi,
// bit-identical results).  1 = one device (default), 0 = all usable devices.  The reference's signature has no room for
// it, so it is a process-wide setting: sdfgen::gpu::num_gpus() = 8;
inline int& num_gpus() { static int n = 1; return n; }

inline void make_level_set3(const std::vector<Vec3ui>& tri, const std::vector<Vec3f>& x, const Vec3f& origin, float dx,
                            int nx, int ny, int nz, Array3f& phi, const int exact_band = 1)
{
    static_assert(sizeof(Vec3ui) == 12 && sizeof(Vec3f) == 12, "Vec3ui / Vec3f must be packed 12-byte PODs");
    if (nx <= 0 || ny <= 0 || nz <= 0) throw std::invalid_argument("Grid dimensions must be positive (nx, ny, nz > 0)");
    phi.resize(nx, ny, nz);
    const float o[3] = {origin[0], origin[1], origin[2]};
    const int rc = num_gpus() == 1
        ? sdfb_make_level_set3(reinterpret_cast<const uint32_t*>(tri.data()), tri.size(),
                               reinterpret_cast<const float*>(x.data()), x.size(), o, dx, nx, ny, nz, exact_band,
                               &phi.a[0], nullptr, nullptr, 0u)
        : sdfb_make_level_set3_multi(reinterpret_cast<const uint32_t*>(tri.data()), tri.size(),
                                     reinterpret_cast<const float*>(x.data()), x.size(), o, dx, nx, ny, nz, exact_band,
                                     &phi.a[0], nullptr, nullptr, num_gpus(), 0u);
    if (rc == SDFB_ERR_INVALID) throw std::invalid_argument(sdfb_last_error());
    if (rc != SDFB_OK) throw std::runtime_error(std::string("sdfgen::gpu::make_level_set3: ") + sdfb_last_error());
}

}  // namespace gpu

inline void make_level_set3(const std::vector<Vec3ui>& tri, const std::vector<Vec3f>& x, const Vec3f& origin, float dx,
                            int nx, int ny, int nz, Array3f& phi, int exact_band = 1,
                            HardwareBackend backend = HardwareBackend::Auto, int num_threads = 0)
{
    (void)num_threads;   // only meaningful for the reference's CPU backend
    if (backend == HardwareBackend::CPU)
        throw std::runtime_error("CPU backend requested but this build is the B200 GPU path only (no CPU fallback)");
    if (!is_gpu_available())
        throw std::runtime_error("GPU backend requested but no sm_100 CUDA device is available (no CPU fallback)");
    gpu::make_level_set3(tri, x, origin, dx, nx, ny, nz, phi, exact_band);
}

}  // namespace sdfgen
