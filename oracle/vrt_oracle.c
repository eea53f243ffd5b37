/* TEST INFRASTRUCTURE ONLY -- see vrt_oracle.h.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off, so
 * that every fused multiply-add below is exactly where an fmaf() is written and nowhere else).
 *
 * Floating-point operation order.  The reference source leaves contraction to the compiler; what its two
 * builds actually execute was read from the compilers' output for the *unmodified* source:
 *   nvcc 12.9 -O2 (PTX of trace_rays_gpu<..,float,float,3>) and g++ 13.3 -O2 -mfma (trace_rays_cpu) both do
 *     lerp  : r = fma(lo, wl, hi * wr)           (x, then y, then z; weights are the 16-bit fraction and
 *                                                 0x10000 - fraction converted to float, cu:145-154)
 *     scale : g = r * 2^-48                       (cu:154)
 *     bend  : dir = fma(invscale, g, dir)         (cu:344-345)
 *     dot   : fma(dz, dz, fma(dx, dx, dy * dy))   (cu:346, tuple_math.h:267)
 *     step  : ((dir * invscale) * ilen), ilen = 0x42000000p0f / dot   (IEEE divide)  (cu:346-347)
 * and differ only in the final float->int rounding (vrt_oracle.h: ROUND_DEVICE / ROUND_HOST).
 */
#include "vrt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* float -> int32 conversions of the two reference builds                                           */

static inline int32_t cvt_device(float x) /* cvt.rni.s32.f32 */
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)rintf(x); /* default rounding mode: nearest, ties to even */
}

static inline int32_t cvt_host(float x) /* static_cast<int32_t>(std::round(x)) on x86-64 (cvttss2si) */
{
    float r = roundf(x);
    if (r != r || r >= 2147483648.0f || r < -2147483648.0f) return INT32_MIN;
    return (int32_t)r;
}

static inline int32_t cvt(float x, int mode) { return mode == VRT_ORACLE_ROUND_HOST ? cvt_host(x) : cvt_device(x); }

/* ------------------------------------------------------------------------------------------------ */
/* the marcher                                                                                      */

static inline float vol_at(const void *vol, int is_i16, size_t idx)
{
    return is_i16 ? (float)((const int16_t *)vol)[idx] : ((const float *)vol)[idx];
}

/* ref: interpolatef 3-D cu:130-155 (generic) / cu:167-188 (AVX) -- 4 channels. */
static inline void sample3(const vrt_oracle_trace_args *a, const uint32_t p[3], float g[4])
{
    const uint32_t by = a->bounds[1], bz = a->bounds[2];
    /* get_index is evaluated in uint32 (cu:113) and then used as an element offset */
    const uint32_t cell = ((p[0] >> 16) * by + (p[1] >> 16)) * bz + (p[2] >> 16);
    const size_t row[4] = {
        (size_t)cell + (size_t)((0u * by + 0u) * bz), (size_t)cell + (size_t)((0u * by + 1u) * bz),
        (size_t)cell + (size_t)((1u * by + 0u) * bz), (size_t)cell + (size_t)((1u * by + 1u) * bz) };
    float v[4][8];
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 8; ++k) v[r][k] = vol_at(a->volume, a->volume_is_i16, row[r] * 4 + (size_t)k);

    uint32_t mr = p[0] & 0xFFFF, ml = 0x10000 - mr;
    float fr = (float)mr, fl = (float)ml;
    for (int k = 0; k < 8; ++k) { v[0][k] = fmaf(v[0][k], fl, v[2][k] * fr); v[1][k] = fmaf(v[1][k], fl, v[3][k] * fr); }
    mr = p[1] & 0xFFFF; ml = 0x10000 - mr; fr = (float)mr; fl = (float)ml;
    for (int k = 0; k < 8; ++k) v[0][k] = fmaf(v[0][k], fl, v[1][k] * fr);
    mr = p[2] & 0xFFFF; ml = 0x10000 - mr; fr = (float)mr; fl = (float)ml;
    for (int k = 0; k < 4; ++k) g[k] = fmaf(v[0][k], fl, v[0][k + 4] * fr) * (1 / 0x1000000000000p0f);
}

/* ref: interpolatef 2-D cu:190-214 -- 3 channels, including its documented quirk: the second x-lerp
 * re-uses values[0] (already lerped) with values[3] (cu:207-208); values[1] is only used by the y-lerp:
 *     a = v0*wl + v2*wr;   b = a*wl + v3*wr;   c = b*wl_y + v1*wr_y;   g = c * 2^-32.
 * Which product of each sum is rounded on its own (the other is fused) was read from the two builds:
 *   DEVICE (nvcc PTX, all channels):   a = fma(wl,v0, v2*wr)   b = fma(wl,a, v3*wr)   c = fma(v1,wr_y, wl_y*b)
 *   HOST   (g++ 13.3, channels 0,1):   a = fma(v2,wr, v0*wl)   b = fma(v3,wr, wl*a)   c = fma(v1,wr_y, wl_y*b)
 *   HOST   (g++ 13.3, channel 2)   :   a = fma(wl,v0, v2*wr)   b = fma(wl,a, v3*wr)   c = fma(wl_y,b, v1*wr_y)
 * (g++ vectorises channels 0,1 and keeps channel 2 scalar, hence the split.) */
static inline void sample2(const vrt_oracle_trace_args *a, const uint32_t p[2], float g[3])
{
    const uint32_t by = a->bounds[1];
    const uint32_t cell = (p[0] >> 16) * by + (p[1] >> 16);
    float v[4][3];
    const size_t off[4] = { (size_t)cell, (size_t)cell + 1, (size_t)cell + by, (size_t)cell + by + 1 };
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 3; ++k) v[r][k] = vol_at(a->volume, a->volume_is_i16, off[r] * 3 + (size_t)k);
    float fr = (float)(p[0] & 0xFFFF), fl = 0x10000 - fr;
    float fry = (float)(p[1] & 0xFFFF), fly = 0x10000 - fry;
    for (int k = 0; k < 3; ++k)
    {
        float t;
        if (a->round_mode == VRT_ORACLE_ROUND_HOST && k < 2)
        {
            t = fmaf(v[2][k], fr, v[0][k] * fl);
            t = fmaf(v[3][k], fr, fl * t);
            t = fmaf(v[1][k], fry, fly * t);
        }
        else if (a->round_mode == VRT_ORACLE_ROUND_HOST)
        {
            t = fmaf(fl, v[0][k], v[2][k] * fr);
            t = fmaf(fl, t, v[3][k] * fr);
            t = fmaf(fly, t, v[1][k] * fry);
        }
        else
        {
            t = fmaf(fl, v[0][k], v[2][k] * fr);
            t = fmaf(fl, t, v[3][k] * fr);
            t = fmaf(v[1][k], fry, fly * t);
        }
        g[k] = t * (1 / 0x100000000p0f);
    }
}

static void trace_one(const vrt_oracle_trace_args *a, const uint32_t *pos_in, const void *dir_in, size_t ray,
                      uint32_t *epos, void *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    const int dim = a->dim;
    uint32_t pos[3] = {0, 0, 0};
    float dir[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d)
    {
        pos[d] = pos_in[ray * dim + d];
        dir[d] = a->dir_is_i16 ? (float)((const int16_t *)dir_in)[ray * dim + d] : ((const float *)dir_in)[ray * dim + d];
        dir[d] *= a->dir_is_i16 ? 256.0f : 65536.0f;                                   /* cu:330-331 */
    }
    uint32_t it = a->iterations;
    uint32_t brightness = 0xFFFFFFFFu;                                                 /* cu:332 */
    uint32_t *pth = path ? path + (size_t)a->iterations * ray * dim : NULL;            /* cu:392 */
    --it;
    if (pth) for (int d = 0; d < dim; ++d) pth[(size_t)it * dim + d] = pos[d];         /* cu:333 */

    uint16_t lim[3];
    for (int d = 0; d < dim; ++d) lim[d] = (uint16_t)(a->bounds[d] - 1);               /* bounds - 1 in uint16, cu:335 */

    for (;;)
    {
        if (!(it-- > 0)) break;                                                        /* cu:335 (short-circuit) */
        int inside = 1;
        for (int d = 0; d < dim; ++d) inside &= ((uint16_t)(pos[d] >> 16) < lim[d]);
        if (!inside) break;
        if (a->translucency)                                                           /* cu:337-341 */
        {
            uint32_t cell = dim == 3 ? ((pos[0] >> 16) * a->bounds[1] + (pos[1] >> 16)) * a->bounds[2] + (pos[2] >> 16)
                                     : (pos[0] >> 16) * a->bounds[1] + (pos[1] >> 16);
            uint32_t absorb = 0xFFFFFFFFu - a->translucency[cell];
            brightness -= brightness < absorb ? brightness : absorb;
            if (brightness < a->min_brightness) break;
        }
        float g[4];
        if (dim == 3) sample3(a, pos, g); else sample2(a, pos, g);                     /* cu:342 */
        if (g[dim] > 0) break;                                                         /* cu:343 */
        for (int d = 0; d < dim; ++d) dir[d] = fmaf(a->invscale[d], g[d], dir[d]);     /* cu:344-345 */
        float dot = dir[1] * dir[1];
        dot = fmaf(dir[0], dir[0], dot);
        if (dim == 3) dot = fmaf(dir[2], dir[2], dot);
        float ilen = 0x42000000p0f / dot;                                              /* cu:346 */
        for (int d = 0; d < dim; ++d)
            pos[d] += (uint32_t)cvt((dir[d] * a->invscale[d]) * ilen, a->round_mode);  /* cu:347 */
        if (pth) for (int d = 0; d < dim; ++d) pth[(size_t)it * dim + d] = pos[d];     /* cu:348 */
    }
    ++it;                                                                              /* cu:350 */
    const uint32_t it_out = it;
    if (pth) while (it-- > 0) for (int d = 0; d < dim; ++d) pth[(size_t)it * dim + d] = pos[d];   /* cu:352-358 */

    for (int d = 0; d < dim; ++d)
    {
        if (a->dir_is_i16) /* cu:359-363: /0x100, round, narrow int -> int16 */
            ((int16_t *)edir)[ray * dim + d] = (int16_t)cvt(dir[d] * (1.0f / 256.0f), a->round_mode);
        else               /* cu:364-368 */
            ((float *)edir)[ray * dim + d] = dir[d] * (1.0f / 65536.0f);
        epos[ray * dim + d] = pos[d];
    }
    light[ray] = a->translucency ? brightness : 0xFFFFFFFFu;                           /* cu:370-373, cu:485 */
    eit[ray] = a->iterations - it_out;                                                 /* cu:953-956 */
}

int vrt_oracle_trace(const vrt_oracle_trace_args *a, size_t n, const uint32_t *pos, const void *dir,
                     uint32_t *epos, void *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    if (!a || (a->dim != 2 && a->dim != 3) || !a->volume) return -1;
    if (path && a->iterations == 0) return -1;
    int threads = a->threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
    #pragma omp parallel for num_threads(threads) schedule(dynamic, 64) if (n > 0x100)
    for (size_t i = 0; i < n; ++i) trace_one(a, pos, dir, i, epos, edir, eit, light, path);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* TraceRaysCu constructor: extra channel + interleave (cu:654-669)                                  */

void vrt_oracle_fold_f32(int dim, size_t nvox, const float *const *diff, const uint32_t *tr, float *out)
{
    for (size_t i = 0; i < nvox; ++i)
    {
        for (int d = 0; d < dim; ++d) out[i * (dim + 1) + d] = diff[d][i];
        out[i * (dim + 1) + dim] = (float)(((int64_t)0x7FFFFFFF - (int64_t)tr[i]) / 0x10000);
    }
}

void vrt_oracle_fold_i16(int dim, size_t nvox, const int16_t *const *diff, const uint32_t *tr, int16_t *out)
{
    for (size_t i = 0; i < nvox; ++i)
    {
        for (int d = 0; d < dim; ++d) out[i * (dim + 1) + d] = diff[d][i];
        out[i * (dim + 1) + dim] = (int16_t)(((int64_t)0x7FFFFFFF - (int64_t)tr[i]) / 0x10000);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* scene prep (f1)                                                                                   */

/* ref: standart_3d_stamp / standart_2d_stamp image_util.cpp:421-425: derivative along the LAST axis. */
static const int STAMP3[27] = { -14,0,14, -47,0,47, -14,0,14,  -47,0,47, -162,0,162, -47,0,47,  -14,0,14, -47,0,47, -14,0,14 };
static const int STAMP2[9]  = { -47,0,47, -162,0,162, -47,0,47 };
#define STAMP_WEIGHT 812 /* sum |STAMP3|; the 2-D path also divides by the 3-D weight (image_util.cpp:479) */

typedef struct { int n; size_t off[18]; int val[18]; } reduced_stamp;

/* ref: stamp_t_struct ctor image_util.cpp:380-414 + convolution::operator() :271-278.  Stamp for axis `ax`
 * is the base stamp with dims ax and dim-1 swapped; non-zero taps in row-major order of the 3^dim stamp;
 * offsets are then re-expressed in the input volume's strides. */
static void make_stamp(int dim, int ax, const size_t *bounds, reduced_stamp *rs)
{
    const int *base = dim == 3 ? STAMP3 : STAMP2;
    const int cnt = dim == 3 ? 27 : 9;
    rs->n = 0;
    for (int j = 0; j < cnt; ++j)
    {
        int p[3] = {0, 0, 0};
        int r = j;
        for (int d = dim - 1; d >= 0; --d) { p[d] = r % 3; r /= 3; }
        int q[3] = { p[0], p[1], p[2] };
        int t = q[ax]; q[ax] = q[dim - 1]; q[dim - 1] = t;       /* swapped stamp: S_ax[p] = S[swap(p)] */
        int lin = 0;
        for (int d = 0; d < dim; ++d) lin = lin * 3 + q[d];
        int v = base[lin];
        if (v == 0) continue;
        size_t off = 0;
        for (int d = 0; d < dim; ++d) off = off * bounds[d] + (size_t)p[d];
        rs->off[rs->n] = off; rs->val[rs->n] = v; ++rs->n;
    }
}

static void crop(int dim, const size_t *bounds, const uint32_t *in, uint32_t *out) /* crop_matrix :300-319 with lower=1 */
{
    size_t ob[3] = {1, 1, 1}, ib[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { ob[3 - dim + d] = bounds[d] - 2; ib[3 - dim + d] = bounds[d]; }
    const size_t lo0 = dim == 3 ? 1 : 0; /* leading padded axis (2-D) is not cropped */
    size_t o = 0;
    for (size_t x = 0; x < (dim == 3 ? ob[0] : 1); ++x)
        for (size_t y = 0; y < ob[1]; ++y)
            for (size_t z = 0; z < ob[2]; ++z)
                out[o++] = in[((x + lo0) * ib[1] + (y + 1)) * ib[2] + (z + 1)];
}

static inline int32_t div_round_closest_i32(int32_t n, int32_t d) /* image_util.h:34-38 */
{
    return ((n < 0) ^ (d < 0)) ? ((n - d / 2) / d) : ((n + d / 2) / d);
}

int vrt_oracle_prep_f32(int dim, const size_t *bounds, const float *ior, const uint32_t *translucency,
                        float *iorlog, float *const *diff, uint32_t *tr_cropped)
{
    if (dim != 2 && dim != 3) return -1;
    size_t nin = 1, nout = 1, ob[3];
    for (int d = 0; d < dim; ++d) { if (bounds[d] < 3) return -1; nin *= bounds[d]; ob[d] = bounds[d] - 2; nout *= ob[d]; }
    crop(dim, bounds, translucency, tr_cropped);
    int bad = 0;
    #pragma omp parallel for reduction(|:bad)
    for (size_t i = 0; i < nin; ++i)
    {
        if (ior[i] <= 0) { bad = 1; continue; }
        /* image_util.cpp:611: `log(_ior[i]) * 0x420000` -- ::log(double), product in double, narrowed on store */
        iorlog[i] = (float)(log((double)ior[i]) * (double)0x420000);
    }
    if (bad) return -2;
    const float weight = (float)STAMP_WEIGHT * (float)0x100;     /* image_util.cpp:438, div = float(0x100) */
    for (int ax = 0; ax < dim; ++ax)
    {
        reduced_stamp rs; make_stamp(dim, ax, bounds, &rs);
        float *out = diff[ax];
        #pragma omp parallel for
        for (size_t o = 0; o < nout; ++o)
        {
            size_t r = o, base = 0, mul = 1;
            for (int d = dim - 1; d >= 0; --d) { base += (r % ob[d]) * mul; r /= ob[d]; mul *= bounds[d]; }
            float sum = 0;
            for (int j = 0; j < rs.n; ++j) sum += (float)rs.val[j] * iorlog[base + rs.off[j]];           /* :284-287, NOT contracted by g++ (checked) */
            out[o] = sum / weight;                                                                      /* :288-291 */
        }
    }
    return 0;
}

int vrt_oracle_prep_u32(int dim, const size_t *bounds, const uint32_t *ior, const uint32_t *translucency,
                        int32_t *iorlog, int16_t *const *diff, uint32_t *tr_cropped)
{
    if (dim != 2 && dim != 3) return -1;
    size_t nin = 1, nout = 1, ob[3];
    for (int d = 0; d < dim; ++d) { if (bounds[d] < 3) return -1; nin *= bounds[d]; ob[d] = bounds[d] - 2; nout *= ob[d]; }
    crop(dim, bounds, translucency, tr_cropped);
    int bad = 0;
    #pragma omp parallel for reduction(|:bad)
    for (size_t i = 0; i < nin; ++i)
    {
        double fior = (double)ior[i] / (double)0x10000;          /* image_util.cpp:532-543 */
        double tmp = log(fior) * (double)0x420000;
        if (tmp > (double)INT32_MAX || tmp < (double)INT32_MIN || tmp != tmp) { bad = 1; continue; }
        iorlog[i] = (int32_t)round(tmp);
    }
    if (bad) return -2;
    const int32_t weight = STAMP_WEIGHT * 0x100;
    int overflow = 0;
    for (int ax = 0; ax < dim; ++ax)
    {
        reduced_stamp rs; make_stamp(dim, ax, bounds, &rs);
        int16_t *out = diff[ax];
        #pragma omp parallel for reduction(|:overflow)
        for (size_t o = 0; o < nout; ++o)
        {
            size_t r = o, base = 0, mul = 1;
            for (int d = dim - 1; d >= 0; --d) { base += (r % ob[d]) * mul; r /= ob[d]; mul *= bounds[d]; }
            int32_t sum = 0;
            for (int j = 0; j < rs.n; ++j) sum = (int32_t)((uint32_t)sum + (uint32_t)(rs.val[j] * iorlog[base + rs.off[j]]));
            sum = div_round_closest_i32(sum, weight);
            out[o] = (int16_t)sum;
            if (out[o] != sum) overflow = 1;                     /* "differention overflow" :293-296 */
        }
    }
    return overflow ? -3 : 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* ray pre-processing (f2) and the host interpolator                                                 */

static inline size_t corner_index(int dim, const size_t *bounds, const uint32_t *pos, int k)
{
    /* interpolator::operator() image_util.h:372-380: bit (dim-1-d) of k selects +1 along axis d */
    size_t idx = 0;
    for (int d = 0; d < dim; ++d) idx = idx * bounds[d] + (size_t)(pos[d] >> 16) + (size_t)((k >> (dim - 1 - d)) & 1);
    return idx;
}

float vrt_oracle_interp_f32(int dim, const size_t *bounds, const float *img, const uint32_t *pos)
{
    float v[8];
    int cnt = 1 << dim;
    for (int k = 0; k < cnt; ++k) v[k] = img[corner_index(dim, bounds, pos, k)];
    for (int d = 0; d < dim; ++d)
    {
        float fr = (float)(pos[d] % 0x10000), fl = (float)(0x10000 - pos[d] % 0x10000);
        cnt >>= 1;
        for (int i = 0; i < cnt; ++i) v[i] = fmaf(v[i], fl, v[i + cnt] * fr);   /* image_util.h:421-424, contracted by g++ */
    }
    return v[0] * (float)(1. / pow(0x10000, dim));
}

uint32_t vrt_oracle_interp_u32(int dim, const size_t *bounds, const uint32_t *img, const uint32_t *pos)
{
    uint32_t v[8];
    int cnt = 1 << dim;
    for (int k = 0; k < cnt; ++k) v[k] = img[corner_index(dim, bounds, pos, k)];
    for (int d = 0; d < dim; ++d)
    {
        uint64_t mr = pos[d] % 0x10000, ml = 0x10000 - mr;
        cnt >>= 1;
        for (int i = 0; i < cnt; ++i) v[i] = (uint32_t)(((uint64_t)v[i] * ml + (uint64_t)v[i + cnt] * mr + 0x8000) / 0x10000); /* :384-387 */
    }
    return v[0];
}

int32_t vrt_oracle_interp_i32(int dim, const size_t *bounds, const int32_t *img, const uint32_t *pos)
{
    int32_t v[8];
    int cnt = 1 << dim;
    for (int k = 0; k < cnt; ++k) v[k] = img[corner_index(dim, bounds, pos, k)];
    for (int d = 0; d < dim; ++d)
    {
        uint64_t mr = pos[d] % 0x10000, ml = 0x10000 - mr;
        cnt >>= 1;
        for (int i = 0; i < cnt; ++i) v[i] = (int32_t)(((uint64_t)(int64_t)v[i] * ml + (uint64_t)(int64_t)v[i + cnt] * mr + 0x8000) / 0x10000);
    }
    return v[0];
}

static inline int in_range(int dim, const size_t *bounds, const uint32_t *p)
{
    for (int d = 0; d < dim; ++d)
        if ((size_t)p[d] < 0x10000 || (size_t)p[d] + 1 >= bounds[d] * 0x10000) return 0;   /* image_util.cpp:686 */
    return 1;
}

long vrt_oracle_normalise_f32(int dim, const size_t *bounds, const float *ior, size_t n, uint32_t *pos, float *dir)
{
    for (size_t i = 0; i < n; ++i) if (!in_range(dim, bounds, pos + i * dim)) return (long)i + 1;
    #pragma omp parallel for if (n > 0x100)
    for (size_t i = 0; i < n; ++i)
    {
        uint32_t *p = pos + i * dim;
        for (int d = 0; d < dim; ++d) p[d] -= 0x8000;
        float nn = vrt_oracle_interp_f32(dim, bounds, ior, p);
        for (int d = 0; d < dim; ++d) dir[i * dim + d] *= nn;
        for (int d = 0; d < dim; ++d) p[d] -= 0x8000;
    }
    return 0;
}

long vrt_oracle_normalise_u32(int dim, const size_t *bounds, const uint32_t *ior, size_t n, uint32_t *pos, int16_t *dir, long *overflow_ray)
{
    if (overflow_ray) *overflow_ray = 0;
    for (size_t i = 0; i < n; ++i) if (!in_range(dim, bounds, pos + i * dim)) return (long)i + 1;
    long ovf = 0;
    #pragma omp parallel for if (n > 0x100)
    for (size_t i = 0; i < n; ++i)
    {
        uint32_t *p = pos + i * dim;
        for (int d = 0; d < dim; ++d) p[d] -= 0x8000;
        int64_t nn = (int64_t)vrt_oracle_interp_u32(dim, bounds, ior, p);
        for (int d = 0; d < dim; ++d)
        {
            int64_t num = (int64_t)dir[i * dim + d] * nn, den = 0x10000;
            int64_t t = (num < 0) ? ((num - den / 2) / den) : ((num + den / 2) / den);         /* divRoundClosest */
            if (t > INT16_MAX || t < INT16_MIN)
            {
                #pragma omp critical
                { if (!ovf || (long)i + 1 < ovf) ovf = (long)i + 1; }                           /* "Normalize length failed" :703 */
            }
            dir[i * dim + d] = (int16_t)t;
        }
        for (int d = 0; d < dim; ++d) p[d] -= 0x8000;
    }
    if (overflow_ray) *overflow_ray = ovf;
    return 0;
}
