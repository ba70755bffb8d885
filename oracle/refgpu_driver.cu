// refgpu_driver.cu -- C door to the reference's OWN CUDA implementation, compiled in place for sm_100a
// (TEST / BENCH INFRASTRUCTURE ONLY: the "existing GPU kernel" comparator of BASELINE.md section 4.5; never linked into
// the product).  The translation unit textually includes /root/reference/gpu_lib/makelevelset3_gpu.cu (path given by
// -DSDFGEN_REFGPU_TU=...), unmodified, built with the reference's own flags (--fmad=false, gpu_lib/CMakeLists.txt:20-27)
// for the one architecture this box has.  It computes a DIFFERENT far field than the CPU path (Jacobi Eikonal,
// gpu_lib/makelevelset3_gpu.cu:487-551,690-699): it is a timing comparator, not a parity oracle.
#include SDFGEN_REFGPU_TU
#include <chrono>
#include <cstdint>

extern "C" int sdfref_gpu_make_level_set3(const uint32_t *tri, uint64_t ntri, const float *x, uint64_t nvert,
                                          const float origin[3], float dx, int ni, int nj, int nk, int exact_band,
                                          float *phi_out, double *seconds_out)
{
    std::vector<Vec3ui> t(ntri);
    std::vector<Vec3f> v(nvert);
    for (uint64_t q = 0; q < ntri; ++q) t[q] = Vec3ui(tri[3 * q], tri[3 * q + 1], tri[3 * q + 2]);
    for (uint64_t q = 0; q < nvert; ++q) v[q] = Vec3f(x[3 * q], x[3 * q + 1], x[3 * q + 2]);
    Array3f phi;
    const auto t0 = std::chrono::steady_clock::now();
    sdfgen::gpu::make_level_set3(t, v, Vec3f(origin[0], origin[1], origin[2]), dx, ni, nj, nk, phi, exact_band);
    cudaDeviceSynchronize();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (phi_out) for (size_t q = 0; q < phi.a.size(); ++q) phi_out[q] = phi.a[q];
    return 0;
}
