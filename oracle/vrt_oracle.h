/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the one hot path of PaulStahr/VolumeRaytracer (the per-ray marcher) and of
 * the two thin steps either side of it (scene prep, ray pre-processing).  It exists so that the CUDA
 * product under volumeraytracer_b200/ can be checked against an independent CPU statement of the
 * reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every function below, bit for
 * bit, with the unmodified reference built from /root/reference/src (oracle/_ref/libvrt_ref.so) on
 * the reference's own known-answer inputs (scaling_test, interpolation_test) and on seeded random
 * scenes; tests/golden/ holds outputs of that reference build for use where /root/reference is absent.
 *
 * "ref:" citations are relative to /root/reference/src.
 */
#ifndef VRT_ORACLE_H
#define VRT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* How  pos += round(step)  rounds (ref: cuda_volume_raytracer.cu:347, tuple_math.h:270-278).
 *  DEVICE: cvt.rni.s32.f32 -- ties to even, saturating, NaN -> 0   (the reference's CUDA build)
 *  HOST  : std::round + cvttss2si -- ties away from zero, NaN / out of range -> INT32_MIN (its CPU build)
 * Everything else (FMA contraction pattern of the lerps, dot product and bend) was read off both
 * compilers' output (nvcc 12.9 PTX, g++ 13.3 -O2 -mfma) and is identical for the 3-D path. */
enum { VRT_ORACLE_ROUND_DEVICE = 0, VRT_ORACLE_ROUND_HOST = 1 };

typedef struct vrt_oracle_trace_args
{
    int            dim;            /* 2 or 3 */
    int            volume_is_i16;  /* interleaved gradient volume element type: 0 float, 1 int16 (diff_t) */
    int            dir_is_i16;     /* ray direction element type: 0 float, 1 int16 (dir_t, unit 0x100) */
    int            round_mode;     /* VRT_ORACLE_ROUND_* */
    uint32_t       bounds[3];      /* gradient-volume extent per axis (axis 0 slowest); compared as uint16 */
    float          invscale[3];
    uint32_t       iterations;     /* cap, ref: raydata_t::_iterations */
    uint32_t       min_brightness; /* only read when translucency != NULL */
    const void    *volume;         /* [nvox][dim+1] interleaved {d0..,extra}  (ref: cu:654-669) */
    const uint32_t*translucency;   /* NULL = shipped behaviour (DummyArray, cu:853...); else live plane (cu:337-341) */
    int            threads;        /* OpenMP threads, <=0: all */
} vrt_oracle_trace_args;

/* ref: trace_ray_function cu:317-374 + fill_struct/read_struct cu:468-516 + cu:953-956.
 * pos/epos: n*dim uint32 16.16; dir/edir: n*dim float or int16; eit/light: n uint32;
 * path: NULL or n*iterations*dim uint32, reverse order per ray (index iterations-1 = start).
 * Returns 0, or -1 on bad arguments. */
int vrt_oracle_trace(const vrt_oracle_trace_args *a, size_t n,
                     const uint32_t *pos, const void *dir,
                     uint32_t *epos, void *edir, uint32_t *eit, uint32_t *light, uint32_t *path);

/* ref: TraceRaysCu ctor cu:644-669 -- extra channel (0x7FFFFFFF - tr)/0x10000 and interleave. */
void vrt_oracle_fold_f32(int dim, size_t nvox, const float *const *diff, const uint32_t *translucency_cropped, float *out);
void vrt_oracle_fold_i16(int dim, size_t nvox, const int16_t *const *diff, const uint32_t *translucency_cropped, int16_t *out);

/* Scene prep ("next" row f1).  ref: image_util.cpp:501-643, convolution :239-298, stamps :421-425.
 * bounds: dim extents of ior/translucency; outputs are (bounds-2) per axis.
 * iorlog: prod(bounds) scratch/out; diff[d]: prod(bounds-2) each; tr_cropped: prod(bounds-2).
 * Returns 0, -1 on bad args, -2 ior <= 0 / log overflow, -3 "differention overflow". */
int vrt_oracle_prep_f32(int dim, const size_t *bounds, const float *ior, const uint32_t *translucency,
                        float *iorlog, float *const *diff, uint32_t *tr_cropped);
int vrt_oracle_prep_u32(int dim, const size_t *bounds, const uint32_t *ior, const uint32_t *translucency,
                        int32_t *iorlog, int16_t *const *diff, uint32_t *tr_cropped);

/* Ray pre-processing ("next" row f2).  ref: image_util.cpp:675-719 (range check, -0x8000, n = interp(ior),
 * dir *= n, -0x8000) and host interpolator<T> image_util.h:348-431.  In place.
 * Returns 0, or 1+index of the first out-of-range ray (the reference throws there). */
long vrt_oracle_normalise_f32(int dim, const size_t *bounds, const float *ior, size_t n, uint32_t *pos, float *dir);
long vrt_oracle_normalise_u32(int dim, const size_t *bounds, const uint32_t *ior, size_t n, uint32_t *pos, int16_t *dir,
                              long *overflow_ray);
float    vrt_oracle_interp_f32(int dim, const size_t *bounds, const float *img, const uint32_t *pos);
uint32_t vrt_oracle_interp_u32(int dim, const size_t *bounds, const uint32_t *img, const uint32_t *pos);
int32_t  vrt_oracle_interp_i32(int dim, const size_t *bounds, const int32_t *img, const uint32_t *pos);

#ifdef __cplusplus
}
#endif
#endif
