// shim_caller.cpp -- a caller written like the reference's own call sites (app/main.cpp:273, tests/test_utils.cpp:27),
// compiled against the REFERENCE's headers (common/array3.h, common/vec.h, included by name, nothing copied) and
// include/sdfgen_b200.hpp.  Built on the CPU box by oracle/Makefile into oracle/_ref/shim_caller (the reference's
// headers do not exist on the GPU box; the binary travels there like oracle/_ref/libsdfgen_ref.so) and run by the
// gpu-marked tests in tests/test_shim.py.
//
//   shim_caller <case.bin> <phi_out.bin> [num_gpus]
// case.bin : int32 nvert, ntri, nx, ny, nz, exact_band; float32 origin[3], dx; float32 x[nvert][3]; uint32 tri[ntri][3]
// phi_out  : float32[nx*ny*nz], the Array3f storage as the call left it (i fastest, common/array3.h:111-115)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sdfgen_b200.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: shim_caller case.bin phi_out.bin [num_gpus]\n"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    int hdr[6]; float par[4];
    if (std::fread(hdr, sizeof(hdr), 1, f) != 1 || std::fread(par, sizeof(par), 1, f) != 1) return 2;
    std::vector<Vec3f> x(hdr[0]);
    std::vector<Vec3ui> tri(hdr[1]);
    if (std::fread(x.data(), sizeof(Vec3f), x.size(), f) != x.size() || std::fread(tri.data(), sizeof(Vec3ui), tri.size(), f) != tri.size()) return 2;
    std::fclose(f);
    if (argc > 3) sdfgen::gpu::num_gpus() = std::atoi(argv[3]);
    std::printf("available=%d\n", (int)sdfgen::is_gpu_available());
    Array3f phi;
    try {
        // the unified entry, as app/main.cpp:273 calls it (backend Auto, num_threads 0)
        sdfgen::make_level_set3(tri, x, Vec3f(par[0], par[1], par[2]), par[3], hdr[2], hdr[3], hdr[4], phi, hdr[5],
                                sdfgen::HardwareBackend::Auto, 0);
    } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 3; }
    std::printf("ok %d %d %d\n", phi.ni, phi.nj, phi.nk);
    // the direct GPU slot must give the same array (gpu_lib/makelevelset3_gpu.h:40-42)
    Array3f phi2;
    sdfgen::gpu::make_level_set3(tri, x, Vec3f(par[0], par[1], par[2]), par[3], hdr[2], hdr[3], hdr[4], phi2, hdr[5]);
    for (size_t i = 0; i < phi.a.size(); ++i) if (phi.a[i] != phi2.a[i] && !(phi.a[i] != phi.a[i])) { std::printf("slot mismatch at %zu\n", i); return 4; }
    try { sdfgen::make_level_set3(tri, x, Vec3f(0, 0, 0), 0.1f, 4, 4, 4, phi2, 1, sdfgen::HardwareBackend::CPU); std::printf("cpu accepted?!\n"); return 5; }
    catch (const std::runtime_error& e) { std::printf("cpu rejected: %s\n", e.what()); }
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) { std::perror(argv[2]); return 2; }
    std::fwrite(&phi.a[0], sizeof(float), phi.a.size(), o);
    std::fclose(o);
    return 0;
}
