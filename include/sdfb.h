/*
 * sdfb.h -- C ABI of the B200-native make_level_set3 (libsdfb.so, sm_100a only).
 *
 * This is the drop-in boundary for ONE hot path of SDFGenFast: triangle mesh -> signed distance
 * grid.  The reference has no C ABI for it; the slot it fills is the C++ symbol
 *     sdfgen::gpu::make_level_set3          /root/reference/gpu_lib/makelevelset3_gpu.h:40-42
 * reached through
 *     sdfgen::make_level_set3 (dispatch)    /root/reference/common/sdfgen_unified.cpp:30-71
 *     sdfgen::is_gpu_available              /root/reference/common/sdfgen_unified.cpp:19-28
 *     Python sdfgen_ext.generate_sdf        /root/reference/python/sdfgen_py.cpp:160-218
 * include/sdfgen_b200.hpp is the C++ shim with the reference's signatures on top of this file and
 * sdfgen_b200/__init__.py is the Python mirror; INTEGRATION.md shows the reference-side edits.
 *
 * Semantics are those of the reference's single-threaded CPU path
 * (/root/reference/cpu_lib/makelevelset3.cpp:192-304): exact band with lowest-triangle-index
 * tie-break, SOS-robust x-ray crossing counts, 2 x 8 Gauss-Seidel closest-triangle sweeps in the
 * reference's order, parity sign.  There is no CPU fallback: every entry point fails with
 * SDFB_ERR_NO_DEVICE when no sm_100 device is usable.
 *
 * Conventions: plain pointers and sizes only; all functions return 0 (SDFB_OK) or a negative
 * SDFB_ERR_* code and leave a message for sdfb_last_error() (thread-local).  Grids are dense,
 * "i fastest" (index i + ni*(j + nj*k), /root/reference/common/array3.h:111-115) unless
 * SDFB_OUT_KFASTEST is given.  Triangles are uint32[ntri][3], vertices float[nvert][3]
 * (Vec3ui / Vec3f PODs, /root/reference/common/vec.h:25-29,180-181).
 */
#ifndef SDFB_H
#define SDFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFB_OK                 0
#define SDFB_ERR_INVALID       -1   /* bad argument (null pointer, non-positive size, dx<=0 ...)   */
#define SDFB_ERR_NO_DEVICE     -2   /* no CUDA device / not an sm_100 part / driver failure        */
#define SDFB_ERR_CUDA          -3   /* a CUDA call failed; message has the CUDA error string       */
#define SDFB_ERR_OOM           -4   /* device or host allocation failed                            */
#define SDFB_ERR_STATE         -5   /* call order violated (e.g. run before a mesh was set)        */
#define SDFB_ERR_LIMIT         -6   /* more than SDFB_MAX_TRIANGLES triangles                      */
#define SDFB_ERR_IO            -7   /* a file could not be opened or written                       */

/* closest-triangle ids share a 32-bit word with a 5-bit sweep stamp */
#define SDFB_MAX_TRIANGLES     134217726u

/* flags */
#define SDFB_OUT_KFASTEST       0x1u  /* phi/tri/count outputs in C order [ni][nj][nk] (k fastest), the
                                         layout sdfgen_ext.generate_sdf returns (sdfgen_py.cpp:80-86)
                                         and the .sdf file stores (common/sdf_io.cpp:49-57)          */
/* Sweep schedules.  All reproduce the reference's serial Gauss-Seidel order bit for bit; they differ in how
 * the work is laid out.  Default: pipelined columns for the first pass of 8 sweeps (most voxels change),
 * fixed-point relaxation for every later sweep (almost nothing changes).  The flags force one schedule for
 * all sweeps (cross-checks and experiments). */
#define SDFB_SWEEP_LEVELS       0x2u  /* one launch per anti-diagonal level (slow, trivially exact)         */
#define SDFB_SWEEP_STRIPS       0x8u  /* retired experiment of round 1: plans created with it are refused   */
#define SDFB_SWEEP_RELAX        0x10u /* fixed-point relaxation for every sweep                             */
#define SDFB_SWEEP_COLUMNS      0x20u /* pipelined columns for every sweep                                  */
#define SDFB_NO_SIGN            0x4u  /* stop before the sign pass (phi stays unsigned)              */

const char *sdfb_version(void);
const char *sdfb_last_error(void);

/* Number of usable sm_100 devices; 0 when none (never negative).  Replaces
 * sdfgen::is_gpu_available (common/sdfgen_unified.cpp:19-28): available == (count > 0). */
int sdfb_device_count(void);

/* Device memory of destroyed plans (and of finished one-shot / batch calls) stays in a library-private pool so that
 * the next plan does not pay cudaMalloc/cudaFree again (tens to hundreds of ms per call); at most SDFB_POOL_RETAIN_MB
 * (environment, default 16384) are kept across synchronisations.  This hands all of it back to the driver -- the
 * state sdfgen::gpu::make_level_set3 leaves behind (gpu_lib/makelevelset3_gpu.cu:767-774 frees everything). */
int sdfb_trim_memory(void);

/* Kernel launches issued by this library in the calling process so far (all plans). */
uint64_t sdfb_launch_count(void);

/*
 * One-shot, host buffers in and out, current device, blocking.  Drop-in for
 * sdfgen::gpu::make_level_set3 (gpu_lib/makelevelset3_gpu.cu:595-777).
 *   phi_out                 ni*nj*nk floats, required
 *   closest_tri_out         ni*nj*nk int32 or NULL  (-1 where no triangle was ever assigned)
 *   intersection_count_out  ni*nj*nk int32 or NULL
 * The two nullable outputs exist because parity is graded on them and the reference keeps them as
 * locals (cpu_lib/makelevelset3.cpp:198-199).
 * A triangle that names a vertex index >= nvert is undefined behaviour in the reference (x[] is indexed unchecked,
 * cpu_lib/makelevelset3.cpp:205); here it is SDFB_ERR_INVALID ("triangle T names a vertex index >= nvert"), found on the
 * device and reported by the calls that deliver results (this one, the batch call, sdfb_plan_download,
 * sdfb_plan_write_sdf); the CUDA context stays usable.
 * Large phi-only calls (>= 64 MB of output) copy phi to the host while the last eight sweeps run and then patch the few
 * values those sweeps changed; the result is the same as a copy after the last sweep.  While the call
 * runs, up to 12 worker threads of the library touch / fill phi_out and stage the mesh.
 */
int sdfb_make_level_set3(const uint32_t *tri, uint64_t ntri, const float *xyz, uint64_t nvert,
                         const float origin[3], float dx, int32_t ni, int32_t nj, int32_t nk,
                         int32_t exact_band, float *phi_out, int32_t *closest_tri_out,
                         int32_t *intersection_count_out, uint32_t flags);

/*
 * Many meshes / grids in one call (the batch API the reference lists as wanted, README.md:216-221; its callers loop
 * over sdfgen::make_level_set3).  Every item is an independent sdfb_make_level_set3 with host buffers; `concurrency`
 * items (<= 0: 4, at most 16) are in flight at a time, each on its own stream with its share of the SMs, so that
 * small grids -- which are launch- and latency-bound one at a time -- overlap.  Results are those of the one-shot
 * call, bit for bit.  item.status receives each item's code; the call returns the first failing item's code.
 * flags: as for sdfb_make_level_set3 (SDFB_OUT_KFASTEST ...).
 */
typedef struct sdfb_batch_item {
    const uint32_t *tri;   uint64_t ntri;     /* uint32[ntri][3]                                     */
    const float    *xyz;   uint64_t nvert;    /* float[nvert][3]                                     */
    float           origin[3];
    float           dx;
    int32_t         ni, nj, nk, exact_band;
    float          *phi_out;                  /* ni*nj*nk floats, host                               */
    int32_t         status;                   /* out                                                 */
} sdfb_batch_item;
int sdfb_make_level_set3_batch(sdfb_batch_item *items, int32_t n, int32_t concurrency, uint32_t flags);

/* ---- plan API: device-resident state, explicit phases, k-slabs for multi-GPU ---------------- */

typedef struct sdfb_plan sdfb_plan;

/*
 * A plan owns all device state for the k-slab [k_lo,k_hi) of a global ni x nj x nk grid on
 * `device` (k_lo=0,k_hi=nk for a single GPU).  State: 8-byte cells {phi, stamp|closest_tri},
 * int32 crossing counts, float output; plus one halo plane of cells below and above the slab.
 */
int sdfb_plan_create(sdfb_plan **plan, int device, int32_t ni, int32_t nj, int32_t nk,
                     int32_t k_lo, int32_t k_hi, uint32_t flags);
int sdfb_plan_destroy(sdfb_plan *plan);

/* Tell the plan that `plans_in_flight` plans run on this device at the same time, each on its own stream: the
 * persistent sweep kernels then take only their share of the SMs so that the plans' kernels overlap instead of
 * queueing behind each other (the sweeps are latency-bound: two 512^3 grids in flight finish sooner than one after
 * the other).  1 = the whole device (default).  Results do not depend on it. */
int sdfb_plan_set_concurrency(sdfb_plan *plan, int32_t plans_in_flight);

/* Mesh from host memory (copied, may be pageable) or already on the plan's device (borrowed until
 * the next set/destroy).  Builds the per-triangle records. `stream` is a cudaStream_t (NULL = default). */
int sdfb_plan_set_mesh_host(sdfb_plan *plan, const uint32_t *tri, uint64_t ntri,
                            const float *xyz, uint64_t nvert, void *stream);
int sdfb_plan_set_mesh_device(sdfb_plan *plan, const uint32_t *d_tri, uint64_t ntri,
                              const float *d_xyz, uint64_t nvert, void *stream);

/* Phase A: init + exact band + crossing counts (cpu_lib/makelevelset3.cpp:196-236). Asynchronous. */
int sdfb_plan_band(sdfb_plan *plan, const float origin[3], float dx, int32_t exact_band, void *stream);
/* Phase B: sweeps first..first+count-1 of the reference's 16 (index s uses direction s%8 of
 * cpu_lib/makelevelset3.cpp:245-248).  Asynchronous.  Sweeps must be run in order since the last sdfb_plan_band:
 * `first` may repeat an index already run but not skip one (SDFB_ERR_STATE) -- the candidate memo relies on every
 * earlier sweep having examined the cells.  Halo planes, if the slab has neighbours and is not linked, must
 * have been filled by the caller (sdfb_plan_device_ptrs + sdfb_plan_halo_refresh) before each call. */
int sdfb_plan_sweep(sdfb_plan *plan, int32_t first, int32_t count, void *stream);
/* Phase C: parity sign + unpack to the float output (cpu_lib/makelevelset3.cpp:295-303). Asynchronous. */
int sdfb_plan_sign(sdfb_plan *plan, void *stream);
/* All three phases with the reference's 16 sweeps. Asynchronous. */
int sdfb_plan_run(sdfb_plan *plan, const float origin[3], float dx, int32_t exact_band, void *stream);

/* Device pointers (valid until destroy): cells uint64[(k_hi-k_lo+2) planes], the first and last
 * plane being the halos; counts int32[slab]; phi float[slab] (layout per plan flags). */
int sdfb_plan_device_ptrs(sdfb_plan *plan, void **cells, void **counts, void **phi);
/* Multi-GPU: call after new contents were written into the two halo planes (plane 0 and plane
 * k_hi-k_lo+1 of `cells`) and before the next sdfb_plan_sweep.  Marks the received cells as freshly
 * changed so the sweep re-examines them (their previous copy may have been stale). Asynchronous.
 * NOT used in the exact multi-slab order: there the caller writes, before sweep s, only the halo plane on the
 * upstream side of that sweep (k direction + - - + + - - +, cpu_lib/makelevelset3.cpp:245-248) with the neighbour's
 * boundary plane as it is AFTER the neighbour's sweep s, stamps untouched, and calls sdfb_plan_sweep(plan, s, 1);
 * the slabs then hold exactly the cells of one whole grid (sdfgen_b200/dist.py: run_sharded_exact). */
int sdfb_plan_halo_refresh(sdfb_plan *plan, void *stream);
/* Number of cells whose closest triangle changed during the sweeps since the last sdfb_plan_band or
 * sdfb_plan_changed call (synchronises the stream; used by the multi-GPU fixed-point loop). */
int sdfb_plan_changed(sdfb_plan *plan, void *stream, uint64_t *changed);
/* Same, plus out[1] = point_triangle_distance evaluations done by the sweeps (column schedule only);
 * out[0] = changed cells.  Both counters are reset. */
int sdfb_plan_counters(sdfb_plan *plan, void *stream, uint64_t out[2]);

/* Verification on the device (blocking), for grids no CPU reference can check in reasonable time or at all (the
 * reference's int index overflows at 2^31 voxels, common/array3.h:59-61):
 *   out[0] = cells of the slab whose distance is NOT bit-identical to point_triangle_distance(voxel, the triangle the
 *            cell names) (cpu_lib/makelevelset3.cpp:49-70), or whose distance is not the initial one although no triangle
 *            was assigned -- 0 for any correct state after band or sweeps;
 *   out[1] = cells without a triangle;
 *   out[2] = order-independent checksum (sum mod 2^64 of a hash of global voxel index and cell word) of the slab: the
 *            slabs of a sharded run add up to the checksum of one plan on the whole grid iff all cells are identical;
 *   out[3] = the same over (distance, triangle) only, without the sweep stamps. */
int sdfb_plan_verify(sdfb_plan *plan, void *stream, uint64_t out[4]);

/* Blocking copies of the slab results to host memory (any may be NULL).  phi is the output of
 * sdfb_plan_sign (or the unsigned cell phi if the sign pass has not run). */
int sdfb_plan_download(sdfb_plan *plan, float *phi_out, int32_t *closest_tri_out,
                       int32_t *intersection_count_out, void *stream);

/* Asynchronous copy of the signed phi of the last sdfb_plan_sign to (pinned) host memory on `copy_stream`, a stream
 * other than the one the phases run on: the copy starts when that sign pass has finished and overlaps whatever is
 * enqueued on the compute stream afterwards (the next mesh upload, band and sweeps); the next sdfb_plan_sign waits for
 * it before it overwrites phi.  The caller synchronises copy_stream (or the device) before reading phi_out.  This is
 * the streaming form of the D2H copy sdfgen::gpu::make_level_set3 does at gpu_lib/makelevelset3_gpu.cu:747-749. */
int sdfb_plan_download_phi_async(sdfb_plan *plan, float *phi_out, void *copy_stream);

/* Writes the signed phi of the last sdfb_plan_sign as a binary .sdf file straight from the device: 36-byte header
 * (3 x int32 dims, 3 x float32 min_box, 3 x float32 min_box + dims*dx) followed by the float32 values, k fastest.
 * Byte-for-byte the file write_sdf_binary(filename, phi_grid, min_box, dx, &inside) produces
 * (/root/reference/common/sdf_io.cpp:10-74, called from app/main.cpp:336), without the host-side Array3f: the
 * k-fastest layout and the inside count (values < 0) are produced on the device and the values stream to the file
 * through pinned staging buffers while the next chunk is copied.  Whole-grid plans only (k_lo = 0, k_hi = nk).
 * inside_count_out may be NULL.  Blocking. */
int sdfb_plan_write_sdf(sdfb_plan *plan, const char *path, const float min_box[3], float dx,
                        int64_t *inside_count_out, void *stream);

/* ---- several GPUs: k-slabs whose sweeps keep the reference's serial order across the slab faces ------------------
 *
 * The reference is single-device; its README lists multi-GPU as future work (README.md:220).  A
 * sweep reads only k offsets 0 and -dk (cpu_lib/makelevelset3.cpp:143-149), so the grid is cut into contiguous k-slabs,
 * one plan per GPU, and each slab's last plane is handed to the next slab column by column WHILE the sweep runs (peer
 * stores over NVLink + system-scope flags inside the sweep kernel; no host synchronisation, no collective).  The result
 * is bit-identical to one plan on the whole grid.  Phases A and C are slab-local (x-rays run along i).
 *
 * One process driving all GPUs: sdfb_make_level_set3_multi.  One process per GPU (torch.distributed / MPI style):
 * every rank creates its slab's plan, calls sdfb_plan_link_export, sends the SDFB_LINK_HANDLE_BYTES bytes to the ranks
 * holding the slabs below and above (any transport), imports theirs with sdfb_plan_link_import, and from then on runs
 *     sdfb_plan_band -> sdfb_plan_sweep(plan, 0, 16) -> sdfb_plan_sign
 * in lockstep with the other ranks (same number of band calls on every slab; the device-side waits carry a watchdog,
 * SDFB_LINK_TIMEOUT_S, default 20 s, after which the kernel traps instead of hanging).  Linked plans always use the
 * column schedule.  Slabs that touch the grid's first or last plane need at least 2 planes.
 */
#define SDFB_LINK_HANDLE_BYTES 128

/* Contiguous k range of slab `index` of `slabs` (the first nk % slabs slabs get one extra plane). */
int sdfb_slab_bounds(int32_t nk, int32_t slabs, int32_t index, int32_t *k_lo, int32_t *k_hi);
/* Allocates this plan's inbound hand-over buffers (16 planes of cells + flags, cudaMalloc) and writes an opaque handle
 * (CUDA IPC handle + slab geometry) to handle_out[SDFB_LINK_HANDLE_BYTES]. */
int sdfb_plan_link_export(sdfb_plan *plan, void *handle_out);
/* Maps the inbound buffers of the plan that holds the neighbouring slab: side 0 = the slab below (its k_hi == this
 * k_lo), 1 = the slab above.  Works across processes (cudaIpcOpenMemHandle) and inside one (peer access is enabled
 * when the devices differ; two slabs may also share a device: then cap their grids with sdfb_plan_set_concurrency). */
int sdfb_plan_link_import(sdfb_plan *plan, int32_t side, const void *handle);
/* Diagnostics (plans created with SDFB_LINK_TRACE=1 in the environment): out[2s] / out[2s+1] = device time (globaltimer,
 * ns) at which the first column of sweep s started / the last one ended on this slab since the last call; 0 = not run. */
int sdfb_plan_link_trace(sdfb_plan *plan, void *stream, uint64_t out[32]);
/* Drops this plan's mappings of its neighbours' buffers (waits for the plan's device first); its own inbound buffers
 * stay allocated until sdfb_plan_destroy.  Order for a clean shutdown: every rank finishes its last run -> barrier -> every
 * rank calls sdfb_plan_unlink -> barrier -> sdfb_plan_destroy (a neighbour's sweep kernel stores into this plan's buffers,
 * and an exported block should outlive the mappings of it). */
int sdfb_plan_unlink(sdfb_plan *plan);
/* sdfb_plan_download into arrays of the WHOLE grid: the slab is written at its place (either layout). */
int sdfb_plan_download_global(sdfb_plan *plan, float *phi_grid, int32_t *closest_tri_grid,
                              int32_t *intersection_count_grid, void *stream);
/* sdfb_make_level_set3 on num_gpus devices (0 = all usable) of the calling process, devices 0 .. num_gpus-1.  Same
 * arguments, same results bit for bit; the SURVEY's proposed `num_gpus` parameter of the C entry
 * (gpu_lib/makelevelset3_gpu.h:40-42 has no such notion).  Falls back to one device for grids with fewer than 2 planes
 * per device. */
int sdfb_make_level_set3_multi(const uint32_t *tri, uint64_t ntri, const float *xyz, uint64_t nvert,
                               const float origin[3], float dx, int32_t ni, int32_t nj, int32_t nk,
                               int32_t exact_band, float *phi_out, int32_t *closest_tri_out,
                               int32_t *intersection_count_out, int32_t num_gpus, uint32_t flags);

/* Device time of the phases of the last completed run, in ms: out[0]=band (init+records+band+counts),
 * out[1]=sweeps, out[2]=sign/unpack, out[3]=total.  Blocks until the work has finished. */
int sdfb_plan_phase_ms(sdfb_plan *plan, float out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SDFB_H */
