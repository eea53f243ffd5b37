/* vrt_b200.h -- C ABI of the B200-native volume ray marcher.
 *
 * This is the drop-in boundary underneath the reference's C++ class TraceRaysCu<DiffType>
 * (reference: src/cuda_volume_raytracer.h:61-115).  Plain pointers and sizes only; no C++ or torch types.
 * Every entry point cites the reference interface it replaces ("ref:" paths are relative to the
 * reference's src/ directory; "cu:" = cuda_volume_raytracer.cu).  INTEGRATION.md shows the bindings a
 * reference maintainer adds on top (the TraceRaysCu<> shim, ctypes, JNI).
 *
 * Data contract (ref: types.h:5-11, cu:103-113):
 *   positions   uint32 16.16 fixed point, `dim` per ray, packed [ray][axis]; CROPPED-volume coordinates
 *               (API coordinate - 0x10000, the caller shifts: image_util.cpp:692,710,770-771)
 *   directions  float32, or int16 with unit 0x100 (dir_t), packed [ray][axis], already multiplied by n(start)
 *   volume      gradient field, `dim`+1 channels per voxel {d0,..,d(dim-1),extra}, float32 or int16 (diff_t),
 *               axis 0 slowest; extra = (0x7FFFFFFF - translucency)/0x10000 (cu:654-659); a ray stops where the
 *               interpolated extra channel is > 0 (cu:343)
 *   translucency uint32 per voxel of the cropped volume (0xFFFFFFFF = clear)
 *
 * All functions return VRT_OK (0) or a VRT_ERR_* code; vrt_last_error() gives the message of the calling
 * thread's last failure.  There is no CPU fallback: without a CUDA device every compute call fails.
 * A scene is immutable after creation; concurrent vrt_trace calls on one scene from several host threads
 * are allowed (ref boundary contract: the JNI binding can be entered from any Java thread).
 */
#ifndef VRT_B200_H
#define VRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VRT_API __attribute__((visibility("default")))
#else
#define VRT_API
#endif

typedef struct vrt_scene vrt_scene; /* opaque; owns the device copy of the volume */

enum { VRT_OK = 0, VRT_ERR_INVALID = 1, VRT_ERR_CUDA = 2, VRT_ERR_NOMEM = 3, VRT_ERR_UNSUPPORTED = 4 };

/* element types */
enum { VRT_F32 = 0, VRT_I16 = 1, VRT_U32 = 2 };

/* vrt_trace flags */
enum {
    VRT_TRACE_DEFAULT           = 0,
    /* attenuate by the translucency plane every step and stop below min_brightness (the template code at
     * cu:337-341,370-373, which the reference's shipped call sites compile out: cu:853,859,...).  Without this
     * flag remaining_light is 0xFFFFFFFF for every ray and min_brightness is ignored, exactly as shipped (cu:785). */
    VRT_TRACE_LIVE_TRANSLUCENCY = 1u << 0,
    /* write the per-step polyline (ref trace_paths; cu:333,348,352-358): path[ray][iterations][dim], REVERSE order */
    VRT_TRACE_PATHS             = 1u << 1,
    /* round like the reference's CPU build: pos += static_cast<int32_t>(std::round(step)) (tuple_math.h:270-278: ties away from
     * zero; NaN / out of range -> INT32_MIN) for the position update (cu:347) and the int16 direction narrowing (cu:359-363), and
     * the CPU build's FMA contraction of the 2-D lerps.  Default: the reference's CUDA build (cvt.rni, ties to even, saturating).
     * With this flag results are bit-identical to the reference's shipped Python module / CI, which link the CPU object
     * (ref Makefile:78-79).  Single-launch marcher only (region mode and the layout variants are not used). */
    VRT_TRACE_ROUND_HOST        = 1u << 2
};

/* vrt_scene_create* flags */
enum {
    VRT_SCENE_DEFAULT = 0,
    VRT_SCENE_BORROW  = 1u << 0, /* vrt_scene_create_device: use the caller's device buffers in place (no copy, not freed) */
    /* layout study: store the volume as 2x2x2-voxel bricks (one 128-byte line per brick for a float scene) instead of the
     * reference's linear interleaved order.  3-D only, no path output; results are bit-identical, only the memory
     * behaviour changes (fewer, fuller lines per cell: helps incoherent ray batches).  Not with VRT_SCENE_BORROW. */
    VRT_SCENE_LAYOUT_BRICK = 1u << 1,
    /* By default the device copy of an int16 (diff_t) 3-D scene is widened to float when it fits (<= 24 GiB): the marcher
     * converts int16 to float before its first multiply anyway (cu:164), so results are bit-identical while the conversions
     * leave the hot loop (+37 % on coherent bundles, at twice the device memory).  This flag keeps the 8-byte voxels. */
    VRT_SCENE_KEEP_I16 = 1u << 2,
    /* LAYOUT STUDY ONLY (library built with -DVRT_STUDY, `make -C volumeraytracer_b200/csrc study`; the shipped library answers
     * VRT_ERR_UNSUPPORTED): additionally stage the (float) volume in a CUDA 3-D array (block-linear tiling) and fetch the corners
     * through a point-sampled float4 texture object; filtering stays in software (hardware trilinear has 8-bit weights, the
     * reference uses 16).  3-D only, no path output; results bit-identical.  Not with VRT_SCENE_BORROW / _KEEP_I16 / _BRICK. */
    VRT_SCENE_LAYOUT_TEXTURE = 1u << 3,
    /* LAYOUT STUDY ONLY (-DVRT_STUDY, see above): store, for every cell, the voxel and its z neighbour side by side (32 bytes per cell, float scenes, 3-D):
     * the two z-adjacent corners of a row become one aligned 256-bit load, a cell change 4 loads / 4 sectors instead of 8
     * loads / ~6 sectors, at twice the device memory.  Results bit-identical; works with path output, live translucency and
     * region mode.  Not with VRT_SCENE_BORROW / _KEEP_I16 / _BRICK / _TEXTURE. */
    VRT_SCENE_LAYOUT_PAIR = 1u << 4
};

/* vrt_scene_set_option keys (tuning; defaults are what bench.py measures) */
enum {
    VRT_OPT_KERNEL        = 0,  /* settable: 0 default (= 3), 1 reference-like (reload every step), 2 register cell cache, 3 cell cache +
                                   packed f32x2 + fast loop for cells without a possibly opaque corner, 6 = 3 + empty-space fast path
                                   (opt-in: coherent bundles through mostly empty volumes).  Chosen implicitly, not settable: 9 = 3
                                   specialised for invscale == (1,1,1) (same bits, fewer instructions; what 0/3 resolve to in that
                                   case), 11 = 9 without the per-cell clear test (scenes without any possibly opaque voxel, VRT_OPT_ALL_CLEAR_KERNEL), 4 = brick layout (VRT_SCENE_LAYOUT_BRICK), 8 = host rounding (VRT_TRACE_ROUND_HOST), 2 for
                                   path output; 5 / 7 = texture / z-pair layouts (study build only) */
    VRT_OPT_BLOCK_THREADS = 1,  /* 32..256, multiple of 32 (the marcher is compiled with __launch_bounds__(256, 4)) */
    VRT_OPT_REFILL        = 2,  /* 0: one ray per thread, no refill; 1..32: a warp fetches new rays when >= this many lanes are idle */
    VRT_OPT_CHUNK_RAYS    = 3,  /* vrt_trace (host buffers): rays per pipelined chunk, 0 = auto */
    VRT_OPT_STEPS_PER_POLL= 4,  /* marching steps between two refill polls */
    VRT_OPT_REGION_LOG2   = 6,  /* round 1's name for VRT_OPT_WAVE_LOG2 (its region mode -- a cub radix sort and a launch per round -- was replaced by
                                   the wavefront marcher at the same marching rate); still accepted: 5..9 = wavefront marcher with that brick log2 */
    /* WAVEFRONT marcher for INCOHERENT batches (csrc/vrt_wave.cuh): one cooperative launch sorts the rays by 2^k-voxel brick with an
       in-kernel counting sort, marches them brick by brick -- persistent warps refilled from the sorted list, every ray until it leaves
       its brick -- and repeats until no ray is left; bit-identical results.  3..8: always, with that k.  -1: never.  0 (default): decided
       per batch by a coherence probe of the ray buffers -- on the host for vrt_trace, by a probe KERNEL for vrt_trace_device (the
       single-launch and the wavefront marcher are then both enqueued, gated on the probe's flag; no synchronisation).  Only large
       3-D batches (>= 2^18 rays) over volumes that do not fit L2 are probed. */
    VRT_OPT_WAVE_LOG2     = 8,
    VRT_OPT_WAVE_MARGIN   = 9,  /* voxels a ray may travel beyond its brick before it is re-bucketed (default 8) */
    VRT_OPT_WAVE_CHECK    = 10, /* marching steps between two refill polls of a warp (default 16) */
    VRT_OPT_WAVE_TAIL_PERMILLE = 11, /* when at most this share of the batch is still alive, the rest is marched without bricks (default 20) */
    VRT_OPT_WAVE_CTAS_PER_SM = 12, /* cap on resident CTAs per SM of the wavefront kernel (0 = the library's choice: the occupancy limit, one less for sparse batches of the all-clear variant) */
    VRT_OPT_WAVE_REFILL   = 14, /* a warp takes new rays from the brick-sorted list when at least this many lanes are idle (default 8) */
    VRT_OPT_ALL_CLEAR_KERNEL = 18, /* 1 (default): a scene in which NO voxel can make a sample opaque (channel 3 carries the sign bit everywhere; counted once at
                                   scene creation, VRT_INFO_ALL_CLEAR) is marched by a variant of the default kernel without the per-cell clear test (KVER 11,
                                   chosen implicitly like 9: same bits), and the wavefront marcher runs its variant that keeps no channel 3 in the cell
                                   cache (64 registers: 4 instead of 3 resident CTAs per SM for dense batches); 0: always the kernels with the test */
    VRT_INFO_ALL_CLEAR    = 103, /* read-only: 1 if no voxel of the scene has a non-negative channel 3 */
    VRT_INFO_WAVE_ROUNDS  = 102, /* read-only: rounds the last wavefront launch on this scene took (synchronises) */
    VRT_OPT_REGION_ROUNDS = 7,  /* accepted for compatibility, unused (the wavefront marcher runs as many rounds as the batch needs) */
    VRT_INFO_EMPTY_PERMILLE = 100, /* read-only (vrt_scene_get_option): share of voxels that are empty space (zero gradient, non-positive
                                      extra channel), in 1/1000 -- the figure to look at before choosing VRT_OPT_KERNEL 6 */
    VRT_INFO_NUM_SMS      = 101, /* read-only: multiprocessor count of the scene's device (cudaGetDeviceProperties) */
    /* read-only, VRT_OPT_KERNEL 10 only (an instrumented copy of the default float / unit-invscale kernel, for measurement: same
       results, slower): VRT_INFO_STAT_BASE + k, k = 0..11 = how many times block k of the marcher was ISSUED (per warp pass) since
       the option was set: 0 outer loop, 1 refill, 2 fast-loop step, 3 cell reload, 4 fast-loop exit checks, 5 generic step,
       6 retire/store, 7 lane-steps (per-thread, not per warp), 8 cell reloads taken by only a part of the active lanes (the
       reconvergence instruction then issues twice), 9..11 unused.  bench.py turns these into the issue-slot roofline. */
    VRT_INFO_STAT_BASE    = 200,
    VRT_OPT_MAX_CTAS_PER_SM = 5 /* persistent mode: cap on resident CTAs per SM (0 = occupancy limit); fewer rays in flight keep an
                                   incoherent batch's working set inside L1/L2 */
};

VRT_API const char *vrt_last_error(void);
VRT_API const char *vrt_version(void);

/* ref: init() cu:82-101 (device count, cached in the global `inited`). */
VRT_API int vrt_device_count(int *count);

/* ref: TraceRaysCu<DiffType>::TraceRaysCu(bounds, diff[dim] planar, translucency_cropped)  cuda_volume_raytracer.h:74-82,
 * cu:637-720: builds the extra channel, interleaves, uploads.  Host pointers.  dim is 2 or 3; bounds[d] <= 65535 and
 * prod(bounds) < 2^32 (the reference indexes in uint16/uint32: cu:113,321). */
VRT_API int vrt_scene_create(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                             const void *const *diff_planes, const uint32_t *translucency_cropped, unsigned flags);

/* Same, from an already interleaved HOST volume [nvox][dim+1] (the layout the ctor produces, cu:660-669). */
VRT_API int vrt_scene_create_interleaved(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                                         const void *volume_interleaved, const uint32_t *translucency_cropped, unsigned flags);

/* Same, from DEVICE memory on `device` (e.g. the buffer an NCCL broadcast just filled: the B200 replacement for the
 * per-device host upload at cu:676-686).  With VRT_SCENE_BORROW the buffers are used in place. */
VRT_API int vrt_scene_create_device(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                                    const void *d_volume_interleaved, const uint32_t *d_translucency_cropped, unsigned flags);

/* Scene prep on the GPU ("next" row f1; ref: RaytraceScene ctor image_util.cpp:501-643 = crop translucency,
 * log(n)*0x420000, 3x3x3 {14,47,162} stencil / (812*256), then the TraceRaysCu ctor).  ior is float32 (float scene) or
 * uint32 16.16 (int16 scene) over the UNCROPPED bounds; the scene's bounds are bounds-2.  ior/translucency are host
 * pointers unless ptrs_on_device != 0. */
VRT_API int vrt_scene_create_from_ior(vrt_scene **out, int device, int dim, const uint64_t *bounds, int ior_dtype,
                                      const void *ior, const uint32_t *translucency, int ptrs_on_device, unsigned flags);

/* ref: TraceRaysCu<DiffType>::~TraceRaysCu cu:974-989. */
VRT_API int vrt_scene_destroy(vrt_scene *scene);

/* Introspection: ref public member TraceRaysCu::_output_sizes (cuda_volume_raytracer.h:73).  Any out pointer may be NULL.
 * *d_volume_interleaved is the device buffer ONLY when it holds the reference's layout in the API element type
 * (volume_bytes = nvox*(dim+1)*sizeof(diff_dtype)); it is NULL when the staged copy differs (an int16 scene widened to
 * float, the brick / study layouts) -- use vrt_scene_storage_info for what is really stored, vrt_scene_export_device /
 * vrt_scene_download for a copy in the reference's layout, and vrt_scene_replicate / vrt_scene_broadcast to copy a scene
 * to other GPUs. */
VRT_API int vrt_scene_info(const vrt_scene *scene, int *device, int *dim, uint64_t *bounds, int *diff_dtype,
                           void **d_volume_interleaved, uint32_t **d_translucency, uint64_t *volume_bytes);
/* What the device copy really is: element type of the staged volume, the VRT_SCENE_LAYOUT_* / _KEEP_I16 flags in effect,
 * its size in bytes and its address. */
VRT_API int vrt_scene_storage_info(const vrt_scene *scene, int *storage_dtype, unsigned *layout_flags, uint64_t *storage_bytes,
                                   void **d_storage);

/* Multi-GPU replication of a staged scene, the B200 replacement for the reference's per-device host upload loop
 * (cu:676-686: one interleave + pageable cudaMemcpy of the whole volume per device).
 *
 * vrt_scene_replicate: ONE process driving several GPUs (the reference's own model, cu:804-843).  The staged buffers of `src`
 * (volume as stored, translucency plane, ior if kept) are copied device-to-device over NVLink peer copies to devices[0..n-1]
 * as a PIPELINED CHAIN src -> devices[0] -> devices[1] ... in 64 MiB slices, so every GPU receives and forwards at link speed
 * and the whole replication takes about one volume's transfer time whatever n is.  out[i] receives the new scene on
 * devices[i] (options copied from src).  *seconds (may be NULL) = wall time of the copies. */
VRT_API int vrt_scene_replicate(const vrt_scene *src, int n, const int *devices, vrt_scene **out, double *seconds);

/* One process per GPU (torchrun, MPI): NCCL.  libnccl.so.2 is loaded with dlopen on first use (the library has no link-time
 * NCCL dependency; without NCCL these calls return VRT_ERR_UNSUPPORTED).  vrt_comm_unique_id fills 128 bytes on one rank; the
 * caller hands them to every rank by whatever means it has (a file, MPI, torch.distributed) and each rank calls
 * vrt_comm_create, which also runs a 4-byte broadcast so that NCCL's lazy connection set-up is not billed to the first scene.
 * vrt_scene_broadcast replicates root's scene: on root `src` is the scene and *out == src; elsewhere src is ignored and *out
 * is a new scene on the communicator's device.  The staged buffers are broadcast in place (no export copy).  *seconds (may be
 * NULL) = device time of the payload broadcasts (CUDA events on the communicator's stream); the ranks meet in a 4-byte all-reduce
 * after the receivers have allocated their replica, so that time holds the transfer and not the receivers' cudaMalloc. */
typedef struct vrt_comm vrt_comm;
VRT_API int vrt_comm_unique_id(void *id_128_bytes);
VRT_API int vrt_comm_create(vrt_comm **out, int device, int rank, int world, const void *id_128_bytes);
VRT_API int vrt_comm_destroy(vrt_comm *comm);
VRT_API int vrt_scene_broadcast(vrt_comm *comm, int root, vrt_scene *src, vrt_scene **out, double *seconds);
/* Copy the staged volume back to the host (the reference keeps such a host copy itself: _diff_interleaved,
 * cuda_volume_raytracer.h:67).  host_volume: volume_bytes; host_translucency: nvox uint32; either may be NULL. */
VRT_API int vrt_scene_download(const vrt_scene *scene, void *host_volume, uint32_t *host_translucency);
/* Device-to-device copy of the staged volume into caller buffers on the scene's device (what rank 0 hands to
 * ncclBroadcast in the multi-GPU set-up; the reference instead re-uploads from the host per device, cu:676-686). */
VRT_API int vrt_scene_export_device(const vrt_scene *scene, void *d_volume_out, uint32_t *d_translucency_out, void *cuda_stream);
VRT_API int vrt_scene_set_option(vrt_scene *scene, int key, int64_t value);
VRT_API int vrt_scene_get_option(const vrt_scene *scene, int key, int64_t *value);

/* ref: TraceRaysCu<DiffType>::trace_rays_cu<DirType>(start_position, start_direction, end_position&, end_direction&,
 * end_iteration&, remaining_light&, path&, scale_vec, minimum_brightness, iterations, trace_paths, Options)
 * cuda_volume_raytracer.h:84-97, cu:722-972.  HOST buffers; blocking.  Outputs are caller-allocated:
 * end_pos n*dim u32, end_dir n*dim (dir_dtype), end_iter n u32, remaining_light n u32, path n*iterations*dim u32
 * (only with VRT_TRACE_PATHS, else may be NULL).  In-place (end_* == start_*) is allowed (the JNI binding does that:
 * java_binding.cpp:158-160).  end_iter follows cu:953-956: k+1 if the ray left the volume after k steps, k if it was
 * stopped in step k (opaque / below min_brightness), `iterations` if the cap was hit. */
VRT_API int vrt_trace(vrt_scene *scene, uint64_t n_rays, const uint32_t *start_pos, const void *start_dir, int dir_dtype,
                      const float *invscale, uint32_t min_brightness, uint32_t iterations, unsigned flags,
                      uint32_t *end_pos, void *end_dir, uint32_t *end_iter, uint32_t *remaining_light, uint32_t *path);

/* ref: the "Warning, maximum iterations hitted" scan over end_iteration (cu:507-515).  1 if any ray of the calling thread's last
 * vrt_trace ended at the iteration cap (end_iter == iterations), 0 if none did, -1 if unknown (no call yet / the call failed).
 * The marcher sets a flag when it retires such a ray, so callers need not scan millions of results on one host thread. */
VRT_API int vrt_trace_cap_hit(void);

/* Same with DEVICE buffers on the scene's device, enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream);
 * returns without synchronising.  This is the call bench.py times for the HBM-resident figure. */
VRT_API int vrt_trace_device(vrt_scene *scene, uint64_t n_rays, const uint32_t *d_start_pos, const void *d_start_dir,
                             int dir_dtype, const float *invscale, uint32_t min_brightness, uint32_t iterations,
                             unsigned flags, uint32_t *d_end_pos, void *d_end_dir, uint32_t *d_end_iter,
                             uint32_t *d_remaining_light, uint32_t *d_path, void *cuda_stream);

/* Ray pre-processing on the GPU ("next" row f2; ref: RaytraceScene::trace_rays image_util.cpp:675-719): range check
 * against the UNCROPPED bounds, pos -= 0x8000, n = trilinear(ior, pos), dir *= n, pos -= 0x8000.  Device buffers, in
 * place.  *first_bad_ray receives 1 + index of the first out-of-range ray (0 if none; the reference throws there).
 * Needs a scene made by vrt_scene_create_from_ior (which keeps ior on the device). */
VRT_API int vrt_normalise_rays_device(vrt_scene *scene, uint64_t n_rays, uint32_t *d_pos, void *d_dir, int dir_dtype,
                                      int64_t *first_bad_ray, void *cuda_stream);

/* Measurement helper for the roofline denominator that MEASURED_PEAKS.json does not hold:
 * random 32-byte-sector gather over a `bytes`-sized device buffer (L2 resident when bytes << L2 size), returns GB/s. */
VRT_API int vrt_measure_gather_bandwidth(int device, uint64_t bytes, int sector_bytes, int iters, double *gb_per_s);

/* Device self-test of the marcher's two exact shortcuts, each over its WHOLE input range: the short division sequence against
   IEEE division (div.rn.f32) of the reference's constant 0x42000000p0f (cu:346) by every float for which the marcher uses it
   (2^-95 <= d < 2^97, 1.6e9 values), and the add-a-constant float-to-integer rounding of the unit-invscale kernels against
   cvt.rni.s32.f32 (cu:347) for every float of magnitude below 2^22 (2.5e9 values).  *mismatches = number of differing results
   (must be 0). */
VRT_API int vrt_selftest_division(int device, uint64_t *mismatches);

/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
VRT_API uint64_t vrt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VRT_B200_H */
