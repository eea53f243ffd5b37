// vrt_prep.cuh -- the two thin steps either side of the marcher, on the GPU ("next" rows f1, f2).
//
// f1  scene prep  (reference: RaytraceScene ctor image_util.cpp:501-643, convolution :239-298, stamps :421-425,
//                  TraceRaysCu ctor cuda_volume_raytracer.cu:654-669)
//       iorlog = log(n) * 0x420000; gradient = 3x3x3 {14,47,162} stencil / (812*256) per axis on the volume
//       shrunk by one voxel per face; extra channel = (0x7FFFFFFF - translucency)/0x10000; interleave.
// f2  ray pre-processing (reference: RaytraceScene::trace_rays image_util.cpp:675-719, interpolator image_util.h:348-431)
//
// Operation order follows what the reference's g++ build executes (checked bit for bit through
// oracle/vrt_oracle.c against oracle/_ref): the float stencil sum is a plain mul + add chain (not fused), the
// float interpolator is r = fma(lo, wl, hi*wr).  The only tolerated deviation is libm: log() of the device
// vs glibc may differ in the last double ulp, which survives the narrowing to float / int32 about once per
// 2^29 voxels.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace vrt {

struct PrepParams
{
    int      dim;
    uint32_t ib[3];   // uncropped bounds (padded at the FRONT with 1 for dim == 2 is NOT done: ib[0..dim-1])
    uint32_t ob[3];   // cropped bounds = ib - 2
    unsigned long long nin, nout;
};

// log(n) * 0x420000 -- float scene: image_util.cpp:611 (double log, double product, narrowed on store)
__global__ void iorlog_f32_kernel(const float *ior, float *iorlog, unsigned long long n, int *bad)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = ior[i];
    if (!(v > 0.0f)) { *bad = 1; iorlog[i] = 0.0f; return; }
    iorlog[i] = (float)(log((double)v) * (double)0x420000);
}

// int16 scene: image_util.cpp:532-543 (round to nearest, ties away)
__global__ void iorlog_u32_kernel(const uint32_t *ior, int32_t *iorlog, unsigned long long n, int *bad)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double fior = (double)ior[i] / (double)0x10000;
    double tmp = log(fior) * (double)0x420000;
    if (!(tmp <= 2147483647.0) || !(tmp >= -2147483648.0)) { *bad = 1; iorlog[i] = 0; return; }
    iorlog[i] = (int32_t)round(tmp);
}

// Output voxel of this thread from the launch geometry -- grid (ceil(oz/128), oy, ox), 128 threads along the contiguous
// axis -- so no 64-bit division is needed per voxel.  A 2-D volume is treated as one slab (ox = 1, no stencil extent in x).
// Returns false for the padding threads; o = linear output index, base = element offset of the stencil block's origin.
__device__ __forceinline__ bool prep_index(const PrepParams &p, unsigned long long &o, unsigned long long &base, unsigned long long &centre)
{
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, x = blockIdx.z;
    const uint32_t oy = p.dim == 3 ? p.ob[1] : p.ob[0], oz = p.dim == 3 ? p.ob[2] : p.ob[1];
    const uint32_t iy = p.dim == 3 ? p.ib[1] : p.ib[0], iz = p.dim == 3 ? p.ib[2] : p.ib[1];
    if (z >= oz) return false;
    o = ((unsigned long long)x * oy + y) * oz + z;
    base = ((unsigned long long)x * iy + y) * iz + z;
    centre = (p.dim == 3 ? (unsigned long long)iy * iz : 0ull) + iz + 1ull;        // crop_matrix :300-319, lower bound 1 per axis
    return true;
}

// The stamps (image_util.cpp:421-425) as compile-time tables.  The base stamp differentiates along the LAST axis:
// S[p0][p1][p2] = W[p0][p1] * D[p2]; the stamp for axis `ax` is the base stamp with axes ax and dim-1 swapped
// (stamp_t_struct ctor :380-414), and convolution::operator() (:271-291) visits its non-zero taps in row-major order.
__device__ __forceinline__ constexpr int stamp3_weight(int ax, int p0, int p1, int p2)
{
    constexpr int W[3][3] = {{14, 47, 14}, {47, 162, 47}, {14, 47, 14}};
    constexpr int D[3] = {-1, 0, 1};
    return ax == 2 ? W[p0][p1] * D[p2] : ax == 0 ? W[p2][p1] * D[p0] : W[p0][p2] * D[p1];
}
__device__ __forceinline__ constexpr int stamp2_weight(int ax, int p0, int p1)
{
    constexpr int W[3] = {47, 162, 47};
    constexpr int D[3] = {-1, 0, 1};
    return ax == 1 ? W[p0] * D[p1] : W[p1] * D[p0];
}

// Scene prep, one thread per cropped voxel: the 3^dim neighbourhood is read once into registers, the dim stencil sums run
// fully unrolled in the reference's tap order, then the extra channel, the interleave and the cropped translucency plane.
// T = float (float scene) or int32_t (int16 scene); arithmetic per type as in convolution::operator().
template <typename T, int DIM>
__device__ __forceinline__ void prep_sums(const T *iorlog, unsigned long long base, unsigned long long sy, unsigned long long sx, T sums[3])
{
    T v[DIM == 3 ? 3 : 1][3][3];
#pragma unroll
    for (int a = 0; a < (DIM == 3 ? 3 : 1); ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[a][b][c] = iorlog[base + a * sx + b * sy + c];
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax)
    {
        T sum = 0;
#pragma unroll
        for (int a = 0; a < (DIM == 3 ? 3 : 1); ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                {
                    const int w = DIM == 3 ? stamp3_weight(ax, a, b, c) : stamp2_weight(ax, b, c);
                    if (w != 0)
                    {
                        if constexpr (std::is_same<T, float>::value) sum = __fadd_rn(sum, __fmul_rn((float)w, v[a][b][c]));      // :284-287, not fused
                        else sum = (T)((uint32_t)sum + (uint32_t)(w * v[a][b][c]));                                               // int32, wraps like the reference's
                    }
                }
        sums[ax] = sum;
    }
}

template <int DIM>
__global__ void prep_f32_kernel(const PrepParams p, const float *iorlog, const uint32_t *translucency,
                                float *volume, uint32_t *tr_cropped)
{
    unsigned long long o, base, centre;
    if (!prep_index(p, o, base, centre)) return;
    const unsigned long long sy = DIM == 3 ? p.ib[2] : p.ib[1], sx = DIM == 3 ? (unsigned long long)p.ib[1] * p.ib[2] : 0ull;
    float sums[3];
    prep_sums<float, DIM>(iorlog, base, sy, sx, sums);
    const float weight = 812.0f * 256.0f;                                  // image_util.cpp:438
    float out[4];
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) out[ax] = __fdiv_rn(sums[ax], weight);       // :288-291
    const uint32_t tr = translucency[base + centre];
    tr_cropped[o] = tr;
    out[DIM] = (float)((long long)(0x7FFFFFFFll - (long long)tr) / 0x10000ll);    // cu:654-659
    float *dst = volume + o * (unsigned long long)(DIM + 1);
    if (DIM == 3) *reinterpret_cast<float4 *>(dst) = make_float4(out[0], out[1], out[2], out[3]);      // one 16-byte store per voxel
    else { dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; }
}

__device__ __forceinline__ int32_t div_round_closest(int32_t n, int32_t d) // image_util.h:34-38
{
    return ((n < 0) ^ (d < 0)) ? ((n - d / 2) / d) : ((n + d / 2) / d);
}

template <int DIM>
__global__ void prep_u32_kernel(const PrepParams p, const int32_t *iorlog, const uint32_t *translucency,
                                int16_t *volume, uint32_t *tr_cropped, int *overflow)
{
    unsigned long long o, base, centre;
    if (!prep_index(p, o, base, centre)) return;
    const unsigned long long sy = DIM == 3 ? p.ib[2] : p.ib[1], sx = DIM == 3 ? (unsigned long long)p.ib[1] * p.ib[2] : 0ull;
    int32_t sums[3];
    prep_sums<int32_t, DIM>(iorlog, base, sy, sx, sums);
    const int32_t weight = 812 * 256;
    int16_t out[4];
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax)
    {
        const int32_t v = div_round_closest(sums[ax], weight);
        out[ax] = (int16_t)v;
        if ((int32_t)out[ax] != v) *overflow = 1;                          // "differention overflow" :293-296
    }
    const uint32_t tr = translucency[base + centre];
    tr_cropped[o] = tr;
    out[DIM] = (int16_t)((long long)(0x7FFFFFFFll - (long long)tr) / 0x10000ll);
    int16_t *dst = volume + o * (unsigned long long)(DIM + 1);
    if (DIM == 3) *reinterpret_cast<short4 *>(dst) = make_short4(out[0], out[1], out[2], out[3]);
    else { dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; }
}

// TraceRaysCu ctor on planar inputs (cu:654-669): extra channel + interleave
template <typename T>
__global__ void fold_kernel(int dim, unsigned long long nvox, const T *d0, const T *d1, const T *d2, const uint32_t *tr, T *out)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvox) return;
    T *dst = out + i * (unsigned long long)(dim + 1);
    dst[0] = d0[i];
    dst[1] = d1[i];
    if (dim == 3) dst[2] = d2[i];
    dst[dim] = (T)((long long)(0x7FFFFFFFll - (long long)tr[i]) / 0x10000ll);
}

// ---------------------------------------------------------------------------------------------------
// f2: ray normalisation

struct NormParams
{
    int      dim;
    uint32_t ib[3];                  // uncropped bounds
    unsigned long long n;
};

__device__ __forceinline__ unsigned long long corner_index(const NormParams &p, const uint32_t *pos, int k)
{
    unsigned long long idx = 0;
    // A ray in the last half voxel of an axis passes the reference's range check (image_util.cpp:686) and then has
    // (pos >> 16) + 1 == bounds: the reference's host interpolator reads past the volume there.  The upper corner is clamped
    // to the last voxel instead (such a ray retires on the marcher's first bounds test, its direction scale is irrelevant).
    for (int d = 0; d < p.dim; ++d) idx = idx * p.ib[d] + min((pos[d] >> 16) + (unsigned)((k >> (p.dim - 1 - d)) & 1), p.ib[d] - 1u);
    return idx;
}

template <bool IOR_F32>
__global__ void normalise_kernel(const NormParams p, const void *ior, uint32_t *pos, void *dir, unsigned long long *first_bad,
                                 unsigned long long *first_overflow)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    uint32_t q[3];
    bool ok = true;
    for (int d = 0; d < p.dim; ++d)
    {
        q[d] = pos[i * p.dim + d];
        // image_util.cpp:686: lhs < 0x10000 || lhs + 1 >= rhs * 0x10000
        if ((unsigned long long)q[d] < 0x10000ull || (unsigned long long)q[d] + 1ull >= (unsigned long long)p.ib[d] * 0x10000ull) ok = false;
    }
    if (!ok) { atomicMin(first_bad, i + 1ull); return; }
    for (int d = 0; d < p.dim; ++d) q[d] -= 0x8000u;
    const int cnt = 1 << p.dim;
    if (IOR_F32)
    {
        float v[8];
        for (int k = 0; k < cnt; ++k) v[k] = ((const float *)ior)[corner_index(p, q, k)];
        int h = cnt;
        for (int d = 0; d < p.dim; ++d)
        {
            const float fr = (float)(q[d] & 0xFFFFu), fl = (float)(0x10000u - (q[d] & 0xFFFFu));
            h >>= 1;
            for (int k = 0; k < h; ++k) v[k] = __fmaf_rn(v[k], fl, __fmul_rn(v[k + h], fr));       // image_util.h:421-424
        }
        const float nn = __fmul_rn(v[0], p.dim == 3 ? 1.0f / 0x1000000000000p0f : 1.0f / 0x100000000p0f);
        float *dd = (float *)dir + i * p.dim;
        for (int d = 0; d < p.dim; ++d) dd[d] = __fmul_rn(dd[d], nn);                               // image_util.cpp:697
    }
    else
    {
        uint32_t v[8];
        for (int k = 0; k < cnt; ++k) v[k] = ((const uint32_t *)ior)[corner_index(p, q, k)];
        int h = cnt;
        for (int d = 0; d < p.dim; ++d)
        {
            const unsigned long long mr = q[d] & 0xFFFFu, ml = 0x10000ull - mr;
            h >>= 1;
            for (int k = 0; k < h; ++k) v[k] = (uint32_t)(((unsigned long long)v[k] * ml + (unsigned long long)v[k + h] * mr + 0x8000ull) / 0x10000ull);
        }
        const long long nn = (long long)v[0];
        short *dd = (short *)dir + i * p.dim;
        for (int d = 0; d < p.dim; ++d)
        {
            const long long num = (long long)dd[d] * nn, den = 0x10000;
            const long long t = (num < 0) ? ((num - den / 2) / den) : ((num + den / 2) / den);      // divRoundClosest, image_util.cpp:701
            if (t > 32767 || t < -32768) atomicMin(first_overflow, i + 1ull);                       // "Normalize length failed"
            dd[d] = (short)t;
        }
    }
    for (int d = 0; d < p.dim; ++d) pos[i * p.dim + d] = q[d] - 0x8000u;
}

// ---------------------------------------------------------------------------------------------------
// roofline helper: random sector gather (SURVEY.md section 8d asks for a measured L2 gather bandwidth)

template <int SECTOR_BYTES>
__global__ void gather_kernel(const uint4 *buf, unsigned long long nsectors, int rounds, unsigned long long *sink)
{
    unsigned long long x = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    unsigned acc = 0;
    for (int r = 0; r < rounds; ++r)
    {
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;                 // xorshift64*
            const unsigned long long s = ((x * 0x2545F4914F6CDD1Dull) >> 11) % nsectors;
            const uint4 *ptr = buf + s * (SECTOR_BYTES / 16);
            uint4 a;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(ptr));
            acc += a.x ^ a.w;
            if (SECTOR_BYTES == 32)
            {
                asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(ptr + 1));
                acc += a.y ^ a.z;
            }
        }
    }
    if (acc == 0x7FFFFFFFu) *sink = acc;
}

} // namespace vrt
