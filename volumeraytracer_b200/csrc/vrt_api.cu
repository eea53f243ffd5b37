// vrt_api.cu -- host side of the C ABI declared in include/vrt_b200.h.
//
// Replaces, for the marcher path only, what the reference does in cuda_volume_raytracer.cu ("cu:"):
//   TraceRaysCu ctor cu:644-720      -> vrt_scene_create*      (fold + interleave on the GPU, one upload)
//   trace_rays_cu_impl cu:774-972    -> vrt_trace / vrt_trace_device (no AoS pack, no 32 768-ray chunking with
//                                       cudaDeviceSynchronize between chunks, no per-call cudaMalloc/cudaFree of
//                                       fixed buffers: stream-ordered allocations, whole-batch launches, copies
//                                       pipelined against compute on two streams)
//   ~TraceRaysCu cu:974-989          -> vrt_scene_destroy
// There is deliberately no CPU path here (the reference falls back to trace_rays_cpu when num_rays <= 0x80 or no
// device exists, cu:804-810): without a CUDA device every call fails with VRT_ERR_CUDA.

#include "../../include/vrt_b200.h"
#include "vrt_march.cuh"
#include "vrt_prep.cuh"
#include "vrt_wave.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <mutex>
#include <string>
#include <vector>

using namespace vrt;

// ---------------------------------------------------------------------------------------------------
// errors

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define VRT_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
        {                                                                                                   \
            char buf__[512];                                                                                \
            snprintf(buf__, sizeof buf__, "%s in %s at line %d (%d)", cudaGetErrorString(e__), __FILE__, __LINE__, (int)e__); \
            cudaGetLastError();                                                                             \
            return fail(e__ == cudaErrorMemoryAllocation ? VRT_ERR_NOMEM : VRT_ERR_CUDA, buf__);            \
        }                                                                                                   \
    } while (0)

struct DeviceGuard
{
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------------
// scene

struct vrt_scene
{
    int       device = 0;
    int       dim = 3;
    uint64_t  bounds[3] = {1, 1, 1};   // cropped (gradient volume) extents
    uint64_t  nvox = 0;
    int       dtype = VRT_F32;     // element type at the API (the reference's DiffType)
    int       store = VRT_F32;     // element type of the device copy (an int16 scene may be staged as float, see apply_storage)
    void     *d_volume = nullptr;
    uint32_t *d_translucency = nullptr;
    bool      owns = true;
    bool      bricked = false;     // VRT_SCENE_LAYOUT_BRICK
    bool      paired = false;      // VRT_SCENE_LAYOUT_PAIR: d_volume holds {voxel, z neighbour} per cell (2 x nvox float4)
    double    flat_fraction = 0.0; // share of voxels with zero gradient and non-positive extra channel (set by apply_storage)
    bool      all_clear = false;   // no voxel of the scene can make a sample opaque: channel 3 carries the sign bit everywhere (set by apply_storage)
    cudaArray_t tex_array = nullptr;   // VRT_SCENE_LAYOUT_TEXTURE: block-linear copy + point-sampled texture object
    cudaTextureObject_t tex = 0;
    uint64_t  nb[3] = {1, 1, 1};   // bricks per axis
    // kept only by vrt_scene_create_from_ior, for vrt_normalise_rays_device (f2)
    void     *d_ior = nullptr;
    int       ior_dtype = VRT_F32;
    uint64_t  ior_bounds[3] = {1, 1, 1};
    bool      owns_ior = false;
    int       num_sms = 148;
    mutable uint32_t *d_wave_info = nullptr; // [kCtlWords] copy of the wavefront marcher's control block after its last launch (rounds, ...)
    int64_t last_wave_rounds_host() const
    {
        if (!d_wave_info) return 0;
        uint32_t h[vrt::kCtlWords] = {};
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        if (cudaMemcpy(h, d_wave_info, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
        return (int64_t)h[vrt::kCtlRounds];
    }
    unsigned long long *d_stats = nullptr;   // VRT_OPT_KERNEL 10: kStatSlots block counters (see kStat* in vrt_march.cuh), zeroed when the option is set
    // options
    std::atomic<int64_t> opt_kernel{0}, opt_block{128}, opt_refill{32}, opt_chunk{0}, opt_poll{128}, opt_max_ctas{0}, opt_region{0}, opt_rounds{12};
    std::atomic<int64_t> opt_wave{0}, opt_wave_margin{8}, opt_wave_check{16}, opt_wave_tail{20}, opt_wave_ctas{0}, opt_wave_refill{8}, opt_allclear{1};
};

static size_t elem_size(int dtype) { return dtype == VRT_I16 ? 2 : 4; }

// ---- per-call device buffers: a LIBRARY-OWNED stream-ordered pool per device ---------------------------------------------
// The reference cudaMalloc/cudaFree's its ray buffers on every call (cu:837-841,958-966).  Here they come from a pool that
// keeps freed blocks cached between calls (release threshold = unlimited) -- but it is the library's own pool, not the
// device's default one, so a host application's allocator (e.g. PyTorch's) is not affected by the attribute, and the cache
// is handed back to the driver (cudaMemPoolTrimTo 0) when the last scene on the device is destroyed.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[64] = {};
static int g_pool_users[64] = {};

static void pool_acquire(int device)
{
    if (device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    ++g_pool_users[device];
    if (g_pool[device]) return;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return; }   // pool_alloc falls back to the default pool
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    g_pool[device] = pool;
}

static void pool_release(int device)
{
    if (device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_users[device] > 0 && --g_pool_users[device] == 0 && g_pool[device])
    {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(g_pool[device], 0);
        cudaGetLastError();
    }
}

static cudaError_t pool_alloc(void **ptr, size_t bytes, int device, cudaStream_t st)
{
    cudaMemPool_t pool = device >= 0 && device < 64 ? g_pool[device] : nullptr;
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, st) : cudaMallocAsync(ptr, bytes, st);
}

static int check_geometry(int dim, const uint64_t *bounds, uint64_t *nvox)
{
    if (dim != 2 && dim != 3) return fail(VRT_ERR_INVALID, "Illegal dimension");                // cu:768-771
    if (!bounds) return fail(VRT_ERR_INVALID, "bounds is null");
    uint64_t n = 1;
    for (int d = 0; d < dim; ++d)
    {
        if (bounds[d] < 2 || bounds[d] > 0xFFFF) return fail(VRT_ERR_INVALID, "bounds must be in [2, 65535] per axis (uint16 in the reference, cu:321)");
        n *= bounds[d];
    }
    if (n >= (1ull << 32)) return fail(VRT_ERR_INVALID, "volume has >= 2^32 voxels (the reference indexes in uint32, cu:113)");
    *nvox = n;
    return VRT_OK;
}

static int new_scene(vrt_scene **out, int device, int dim, const uint64_t *bounds, int dtype)
{
    if (!out) return fail(VRT_ERR_INVALID, "out is null");
    *out = nullptr;
    if (dtype != VRT_F32 && dtype != VRT_I16) return fail(VRT_ERR_INVALID, "diff_dtype must be VRT_F32 or VRT_I16");
    uint64_t nvox = 0;
    int rc = check_geometry(dim, bounds, &nvox);
    if (rc) return rc;
    int count = 0;
    VRT_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(VRT_ERR_INVALID, "no such CUDA device");
    vrt_scene *s = new vrt_scene();
    s->device = device; s->dim = dim; s->dtype = dtype; s->store = dtype; s->nvox = nvox;
    for (int d = 0; d < dim; ++d) s->bounds[d] = bounds[d];
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->num_sms = prop.multiProcessorCount;
    pool_acquire(device);
    cudaGetLastError();
    *out = s;
    return VRT_OK;
}

static int alloc_scene_buffers(vrt_scene *s)
{
    VRT_CUDA(cudaMalloc(&s->d_volume, s->nvox * (s->dim + 1) * elem_size(s->dtype)));
    VRT_CUDA(cudaMalloc((void **)&s->d_translucency, s->nvox * sizeof(uint32_t)));
    s->owns = true;
    return VRT_OK;
}

// ---- staging: how the device copy of the volume is stored -------------------------------------------------------
// Every creator first builds the reference's layout (linear interleaved, API element type).  apply_storage() then
//   * widens an int16 scene to float on the device (exact: the marcher converts int16 -> float before its first
//     multiply anyway, cu:164/176), which takes the conversions out of the cell reload: +37 % on coherent bundles at
//     twice the memory; skipped with VRT_SCENE_KEEP_I16, for borrowed buffers, for 2-D and above 24 GiB;
//   * re-orders it into 2x2x2 bricks when VRT_SCENE_LAYOUT_BRICK is set.
// linearise() is the inverse, used by vrt_scene_download / vrt_scene_export_device.

__device__ __forceinline__ bool sign_clear(float v) { return (__float_as_uint(v) >> 31) == 0u; }
__device__ __forceinline__ bool sign_clear(short v) { return v >= 0; }
// count[0]: voxels that are empty space (zero gradient, non-positive channel 3); count[1]: voxels whose channel 3 does NOT carry the
// sign bit, i.e. voxels that can make a sample opaque (cu:343).  A scene without any lets the marcher drop its per-cell test (KVER 11).
template <typename T>
__global__ void count_flat_kernel(const T *vol, unsigned long long nvox, unsigned long long *count)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool flat = false, solid = false;
    if (i < nvox)
    {
        const T *v = vol + i * 4;
        flat = v[0] == T(0) && v[1] == T(0) && v[2] == T(0) && v[3] <= T(0);
        solid = sign_clear(v[3]);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, flat), m2 = __ballot_sync(0xFFFFFFFFu, solid);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
    if ((threadIdx.x & 31u) == 0 && m2) atomicAdd(count + 1, (unsigned long long)__popc(m2));
}

__global__ void widen_i16_kernel(const short *in, float *out, unsigned long long n)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void narrow_f32_kernel(const float *in, short *out, unsigned long long n)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (short)__float2int_rn(in[i]);
}

static int brick_convert(const vrt_scene *s, const void *src, void *dst, int to_brick, cudaStream_t st)
{
    const unsigned long long nslots = s->nb[0] * s->nb[1] * s->nb[2] * 8ull;
    const unsigned grid = (unsigned)((nslots + 255) / 256);
    if (s->store == VRT_F32)
        brick_convert_kernel<float4><<<grid, 256, 0, st>>>((const float4 *)src, (float4 *)dst, (uint32_t)s->bounds[0], (uint32_t)s->bounds[1], (uint32_t)s->bounds[2],
                                                          (uint32_t)s->nb[1], (uint32_t)s->nb[2], nslots, to_brick);
    else
        brick_convert_kernel<int2><<<grid, 256, 0, st>>>((const int2 *)src, (int2 *)dst, (uint32_t)s->bounds[0], (uint32_t)s->bounds[1], (uint32_t)s->bounds[2],
                                                        (uint32_t)s->nb[1], (uint32_t)s->nb[2], nslots, to_brick);
    ++g_launches;
    VRT_CUDA(cudaGetLastError());
    return VRT_OK;
}

static int apply_storage(vrt_scene *s, unsigned flags)
{
    const unsigned long long nelem = s->nvox * (unsigned long long)(s->dim + 1);
    if (s->dim == 3)   // how much of the volume is empty space?  (decides whether the default kernel takes the fast path)
    {
        unsigned long long *d_cnt = nullptr, h_cnt[2] = {0, 0};
        VRT_CUDA(cudaMalloc((void **)&d_cnt, 16));
        VRT_CUDA(cudaMemset(d_cnt, 0, 16));
        const unsigned grid = (unsigned)((s->nvox + 255) / 256);
        if (s->dtype == VRT_F32) count_flat_kernel<float><<<grid, 256>>>((const float *)s->d_volume, s->nvox, d_cnt);
        else                     count_flat_kernel<short><<<grid, 256>>>((const short *)s->d_volume, s->nvox, d_cnt);
        ++g_launches;
        cudaError_t e = cudaMemcpy(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost);
        cudaFree(d_cnt);
        VRT_CUDA(e);
        s->flat_fraction = (double)h_cnt[0] / (double)s->nvox;
        s->all_clear = h_cnt[1] == 0;
    }
    if (s->owns && s->dtype == VRT_I16 && s->dim == 3 && !(flags & VRT_SCENE_KEEP_I16) && nelem * 4ull <= (24ull << 30))
    {
        void *wide = nullptr;
        if (cudaMalloc(&wide, nelem * 4) == cudaSuccess)
        {
            widen_i16_kernel<<<(unsigned)((nelem + 255) / 256), 256>>>((const short *)s->d_volume, (float *)wide, nelem);
            ++g_launches;
            cudaError_t e = cudaGetLastError();
            e = e == cudaSuccess ? cudaDeviceSynchronize() : e;
            if (e != cudaSuccess) { cudaFree(wide); cudaGetLastError(); return fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
            cudaFree(s->d_volume);
            s->d_volume = wide;
            s->store = VRT_F32;
        }
        else cudaGetLastError();       // not enough memory for the wide copy: keep int16
    }
#ifndef VRT_STUDY
    if (flags & (VRT_SCENE_LAYOUT_TEXTURE | VRT_SCENE_LAYOUT_PAIR))
        return fail(VRT_ERR_UNSUPPORTED, "VRT_SCENE_LAYOUT_TEXTURE / _PAIR are layout-study variants: build the library with -DVRT_STUDY (make study)");
#endif
    if (flags & VRT_SCENE_LAYOUT_TEXTURE)
    {
        if (s->dim != 3 || s->store != VRT_F32 || (flags & VRT_SCENE_LAYOUT_BRICK))
            return fail(VRT_ERR_UNSUPPORTED, "VRT_SCENE_LAYOUT_TEXTURE needs a 3-D scene staged as float and excludes VRT_SCENE_LAYOUT_BRICK / _KEEP_I16 / _BORROW");
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
        VRT_CUDA(cudaMalloc3DArray(&s->tex_array, &desc, make_cudaExtent(s->bounds[2], s->bounds[1], s->bounds[0])));
        cudaMemcpy3DParms cp;
        memset(&cp, 0, sizeof cp);
        cp.srcPtr = make_cudaPitchedPtr(s->d_volume, s->bounds[2] * sizeof(float4), s->bounds[2], s->bounds[1]);
        cp.dstArray = s->tex_array;
        cp.extent = make_cudaExtent(s->bounds[2], s->bounds[1], s->bounds[0]);
        cp.kind = cudaMemcpyDeviceToDevice;
        VRT_CUDA(cudaMemcpy3D(&cp));
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray; rd.res.array.array = s->tex_array;
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        VRT_CUDA(cudaCreateTextureObject(&s->tex, &rd, &td, nullptr));
        return VRT_OK;
    }
    if (flags & VRT_SCENE_LAYOUT_PAIR)
    {
        if (s->dim != 3 || s->store != VRT_F32 || !s->owns || (flags & VRT_SCENE_LAYOUT_BRICK))
            return fail(VRT_ERR_UNSUPPORTED, "VRT_SCENE_LAYOUT_PAIR needs a 3-D scene staged as float and excludes VRT_SCENE_LAYOUT_BRICK / _TEXTURE / _KEEP_I16 / _BORROW");
        void *dst = nullptr;
        cudaError_t e = cudaMalloc(&dst, s->nvox * 32ull);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(VRT_ERR_NOMEM, "not enough device memory for the pair layout"); }
        pair_convert_kernel<<<(unsigned)((s->nvox + 255) / 256), 256>>>((const float4 *)s->d_volume, (float4 *)dst, s->nvox);
        ++g_launches;
        e = cudaGetLastError();
        e = e == cudaSuccess ? cudaDeviceSynchronize() : e;
        if (e != cudaSuccess) { cudaFree(dst); cudaGetLastError(); return fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
        cudaFree(s->d_volume);
        s->d_volume = dst;
        s->paired = true;
        return VRT_OK;
    }
    if (!(flags & VRT_SCENE_LAYOUT_BRICK)) return VRT_OK;
    if (s->dim != 3) return fail(VRT_ERR_UNSUPPORTED, "VRT_SCENE_LAYOUT_BRICK is 3-D only");
    if (!s->owns) return fail(VRT_ERR_UNSUPPORTED, "VRT_SCENE_LAYOUT_BRICK cannot be combined with VRT_SCENE_BORROW");
    for (int d = 0; d < 3; ++d) s->nb[d] = (s->bounds[d] + 1) / 2;
    const unsigned long long nslots = s->nb[0] * s->nb[1] * s->nb[2] * 8ull;
    if (nslots >= (1ull << 32)) return fail(VRT_ERR_INVALID, "bricked volume has >= 2^32 voxel slots");
    void *dst = nullptr;
    VRT_CUDA(cudaMalloc(&dst, nslots * 4 * elem_size(s->store)));
    int rc = brick_convert(s, s->d_volume, dst, 1, nullptr);
    if (rc == VRT_OK) { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) rc = fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
    if (rc) { cudaFree(dst); return rc; }
    cudaFree(s->d_volume);
    s->d_volume = dst;
    s->bricked = true;
    return VRT_OK;
}

// write the volume in the reference's layout and API element type into d_out (device memory, nvox*(dim+1) elements)
static int linearise(const vrt_scene *s, void *d_out, cudaStream_t st)
{
    const unsigned long long nelem = s->nvox * (unsigned long long)(s->dim + 1);
    const bool narrow = s->store != s->dtype;
    if (s->paired)      // first voxel of every pair: a strided copy
    {
        void *lin = d_out;
        if (narrow) VRT_CUDA(pool_alloc(&lin, nelem * 4, s->device, st));
        cudaError_t e = cudaMemcpy2DAsync(lin, 16, s->d_volume, 32, 16, s->nvox, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && narrow)
        {
            narrow_f32_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>((const float *)lin, (short *)d_out, nelem);
            ++g_launches;
            e = cudaGetLastError();
        }
        if (narrow) cudaFreeAsync(lin, st);
        VRT_CUDA(e);
        return VRT_OK;
    }
    if (!s->bricked && !narrow)
    {
        VRT_CUDA(cudaMemcpyAsync(d_out, s->d_volume, nelem * elem_size(s->dtype), cudaMemcpyDeviceToDevice, st));
        return VRT_OK;
    }
    const void *lin = s->d_volume;
    void *tmp = nullptr;
    if (s->bricked)
    {
        void *dst = d_out;
        if (narrow) { VRT_CUDA(pool_alloc(&tmp, nelem * elem_size(s->store), s->device, st)); dst = tmp; }
        int rc = brick_convert(s, s->d_volume, dst, 0, st);
        if (rc) { if (tmp) cudaFreeAsync(tmp, st); return rc; }
        lin = dst;
    }
    if (narrow)
    {
        narrow_f32_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>((const float *)lin, (short *)d_out, nelem);
        ++g_launches;
        cudaError_t e = cudaGetLastError();
        if (tmp) cudaFreeAsync(tmp, st);
        VRT_CUDA(e);
    }
    return VRT_OK;
}

extern "C" {

const char *vrt_last_error(void) { return g_last_error.c_str(); }
const char *vrt_version(void) { return "volumeraytracer_b200 0.1 (sm_100a)"; }
uint64_t vrt_launch_count(void) { return g_launches.load(); }
#ifdef VRT_CHECK
// negative control for the test: one deliberately failing check (thread 0 of one warp)
__global__ void chk_selftest_kernel() { VRT_CHK(threadIdx.x != 0); }
VRT_API int vrt_check_selftest(int device)
{
    DeviceGuard g(device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    chk_selftest_kernel<<<1, 32>>>();
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}
// bounds-checked build only: out-of-range accesses the kernels have counted on `device` since the library was loaded
VRT_API int vrt_check_violations(int device, uint64_t *count)
{
    if (!count) return fail(VRT_ERR_INVALID, "count is null");
    DeviceGuard g(device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    unsigned long long v = 0;
    VRT_CUDA(cudaDeviceSynchronize());
    VRT_CUDA(cudaMemcpyFromSymbol(&v, vrt::g_vrt_violations, sizeof v));
    *count = v;
    return VRT_OK;
}
#endif

int vrt_device_count(int *count)
{
    if (!count) return fail(VRT_ERR_INVALID, "count is null");
    *count = 0;
    VRT_CUDA(cudaGetDeviceCount(count));
    return VRT_OK;
}

int vrt_scene_destroy(vrt_scene *s)
{
    if (!s) return VRT_OK;
    DeviceGuard g(s->device);
    if (s->tex) cudaDestroyTextureObject(s->tex);
    if (s->tex_array) cudaFreeArray(s->tex_array);
    if (s->owns) { cudaFree(s->d_volume); cudaFree(s->d_translucency); }
    if (s->owns_ior) cudaFree(s->d_ior);
    if (s->d_stats) cudaFree(s->d_stats);
    if (s->d_wave_info) cudaFree(s->d_wave_info);
    cudaGetLastError();
    pool_release(s->device);
    delete s;
    return VRT_OK;
}

int vrt_scene_create(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                     const void *const *diff_planes, const uint32_t *translucency_cropped, unsigned flags)
{
    if (!diff_planes || !translucency_cropped) return fail(VRT_ERR_INVALID, "null input");
    for (int d = 0; d < dim && d < 3; ++d) if (!diff_planes[d]) return fail(VRT_ERR_INVALID, "null diff plane");
    vrt_scene *s = nullptr;
    int rc = new_scene(&s, device, dim, bounds, diff_dtype);
    if (rc) return rc;
    DeviceGuard g(device);
    rc = alloc_scene_buffers(s);
    if (rc) { vrt_scene_destroy(s); return rc; }
    const size_t es = elem_size(diff_dtype), plane_bytes = s->nvox * es;
    void *tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, plane_bytes * dim);
    if (e != cudaSuccess) { vrt_scene_destroy(s); return fail(VRT_ERR_NOMEM, cudaGetErrorString(e)); }
    for (int d = 0; d < dim; ++d)
        e = e == cudaSuccess ? cudaMemcpy((char *)tmp + plane_bytes * d, diff_planes[d], plane_bytes, cudaMemcpyHostToDevice) : e;
    e = e == cudaSuccess ? cudaMemcpy(s->d_translucency, translucency_cropped, s->nvox * 4, cudaMemcpyHostToDevice) : e;
    if (e == cudaSuccess)
    {
        const unsigned blocks = (unsigned)((s->nvox + 255) / 256);
        const char *t = (const char *)tmp;
        if (diff_dtype == VRT_F32)
            fold_kernel<float><<<blocks, 256>>>(dim, s->nvox, (const float *)t, (const float *)(t + plane_bytes), (const float *)(t + 2 * plane_bytes * (dim == 3)), s->d_translucency, (float *)s->d_volume);
        else
            fold_kernel<int16_t><<<blocks, 256>>>(dim, s->nvox, (const int16_t *)t, (const int16_t *)(t + plane_bytes), (const int16_t *)(t + 2 * plane_bytes * (dim == 3)), s->d_translucency, (int16_t *)s->d_volume);
        ++g_launches;
        e = cudaGetLastError();
        e = e == cudaSuccess ? cudaDeviceSynchronize() : e;
    }
    cudaFree(tmp);
    if (e != cudaSuccess) { vrt_scene_destroy(s); cudaGetLastError(); return fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
    rc = apply_storage(s, flags);
    if (rc) { vrt_scene_destroy(s); return rc; }
    *out = s;
    return VRT_OK;
}

int vrt_scene_create_interleaved(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                                 const void *volume_interleaved, const uint32_t *translucency_cropped, unsigned flags)
{
    if (!volume_interleaved || !translucency_cropped) return fail(VRT_ERR_INVALID, "null input");
    vrt_scene *s = nullptr;
    int rc = new_scene(&s, device, dim, bounds, diff_dtype);
    if (rc) return rc;
    DeviceGuard g(device);
    rc = alloc_scene_buffers(s);
    if (rc) { vrt_scene_destroy(s); return rc; }
    cudaError_t e = cudaMemcpy(s->d_volume, volume_interleaved, s->nvox * (dim + 1) * elem_size(diff_dtype), cudaMemcpyHostToDevice);
    e = e == cudaSuccess ? cudaMemcpy(s->d_translucency, translucency_cropped, s->nvox * 4, cudaMemcpyHostToDevice) : e;
    if (e != cudaSuccess) { vrt_scene_destroy(s); cudaGetLastError(); return fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
    rc = apply_storage(s, flags);
    if (rc) { vrt_scene_destroy(s); return rc; }
    *out = s;
    return VRT_OK;
}

int vrt_scene_create_device(vrt_scene **out, int device, int dim, const uint64_t *bounds, int diff_dtype,
                            const void *d_volume_interleaved, const uint32_t *d_translucency_cropped, unsigned flags)
{
    if (!d_volume_interleaved) return fail(VRT_ERR_INVALID, "null input");
    vrt_scene *s = nullptr;
    int rc = new_scene(&s, device, dim, bounds, diff_dtype);
    if (rc) return rc;
    DeviceGuard g(device);
    if (flags & VRT_SCENE_BORROW)
    {
        s->d_volume = const_cast<void *>(d_volume_interleaved);
        s->d_translucency = const_cast<uint32_t *>(d_translucency_cropped);
        s->owns = false;
    }
    else
    {
        rc = alloc_scene_buffers(s);
        if (rc) { vrt_scene_destroy(s); return rc; }
        cudaError_t e = cudaMemcpy(s->d_volume, d_volume_interleaved, s->nvox * (dim + 1) * elem_size(diff_dtype), cudaMemcpyDeviceToDevice);
        if (d_translucency_cropped)
            e = e == cudaSuccess ? cudaMemcpy(s->d_translucency, d_translucency_cropped, s->nvox * 4, cudaMemcpyDeviceToDevice) : e;
        else
            e = e == cudaSuccess ? cudaMemset(s->d_translucency, 0xFF, s->nvox * 4) : e;
        if (e != cudaSuccess) { vrt_scene_destroy(s); cudaGetLastError(); return fail(VRT_ERR_CUDA, cudaGetErrorString(e)); }
    }
    rc = apply_storage(s, flags);
    if (rc) { vrt_scene_destroy(s); return rc; }
    *out = s;
    return VRT_OK;
}

int vrt_scene_create_from_ior(vrt_scene **out, int device, int dim, const uint64_t *bounds, int ior_dtype,
                              const void *ior, const uint32_t *translucency, int ptrs_on_device, unsigned flags)
{
    if (!ior || !translucency || !bounds) return fail(VRT_ERR_INVALID, "null input");
    if (ior_dtype != VRT_F32 && ior_dtype != VRT_U32) return fail(VRT_ERR_INVALID, "ior_dtype must be VRT_F32 or VRT_U32");
    if (dim != 2 && dim != 3) return fail(VRT_ERR_INVALID, "Illegal dimension: " + std::to_string(dim));   // image_util.cpp:558
    uint64_t cb[3] = {1, 1, 1}, nin = 1;
    for (int d = 0; d < dim; ++d)
    {
        if (bounds[d] < 4) return fail(VRT_ERR_INVALID, "bounds must be >= 4 per axis");
        cb[d] = bounds[d] - 2; nin *= bounds[d];
    }
    if (nin >= (1ull << 31)) return fail(VRT_ERR_INVALID, "volume too large for the 32-bit stencil offsets");
    vrt_scene *s = nullptr;
    int rc = new_scene(&s, device, dim, cb, ior_dtype == VRT_F32 ? VRT_F32 : VRT_I16);
    if (rc) return rc;
    DeviceGuard g(device);
    rc = alloc_scene_buffers(s);
    if (rc) { vrt_scene_destroy(s); return rc; }
    s->ior_dtype = ior_dtype;
    for (int d = 0; d < dim; ++d) s->ior_bounds[d] = bounds[d];

    void *d_iorlog = nullptr; uint32_t *d_tr = nullptr; int *d_flag = nullptr;
    cudaError_t e = cudaMalloc(&s->d_ior, nin * 4);
    s->owns_ior = e == cudaSuccess;
    e = e == cudaSuccess ? cudaMalloc(&d_iorlog, nin * 4) : e;
    e = e == cudaSuccess ? cudaMalloc(&d_flag, 2 * sizeof(int)) : e;
    e = e == cudaSuccess ? cudaMemset(d_flag, 0, 2 * sizeof(int)) : e;
    const cudaMemcpyKind kind = ptrs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    e = e == cudaSuccess ? cudaMemcpy(s->d_ior, ior, nin * 4, kind) : e;
    if (ptrs_on_device) d_tr = const_cast<uint32_t *>(translucency);
    else
    {
        e = e == cudaSuccess ? cudaMalloc((void **)&d_tr, nin * 4) : e;
        e = e == cudaSuccess ? cudaMemcpy(d_tr, translucency, nin * 4, cudaMemcpyHostToDevice) : e;
    }
    int flags_h[2] = {0, 0};
    if (e == cudaSuccess)
    {
        PrepParams pp;
        pp.dim = dim; pp.nin = nin; pp.nout = s->nvox;
        for (int d = 0; d < 3; ++d) { pp.ib[d] = d < dim ? (uint32_t)bounds[d] : 1; pp.ob[d] = d < dim ? (uint32_t)cb[d] : 1; }
        const unsigned bin = (unsigned)((nin + 255) / 256);
        const dim3 bout = dim == 3 ? dim3((unsigned)((cb[2] + 127) / 128), (unsigned)cb[1], (unsigned)cb[0])
                                   : dim3((unsigned)((cb[1] + 127) / 128), (unsigned)cb[0], 1u);
        if (ior_dtype == VRT_F32)
        {
            iorlog_f32_kernel<<<bin, 256>>>((const float *)s->d_ior, (float *)d_iorlog, nin, d_flag);
            if (dim == 3) prep_f32_kernel<3><<<bout, 128>>>(pp, (const float *)d_iorlog, d_tr, (float *)s->d_volume, s->d_translucency);
            else          prep_f32_kernel<2><<<bout, 128>>>(pp, (const float *)d_iorlog, d_tr, (float *)s->d_volume, s->d_translucency);
        }
        else
        {
            iorlog_u32_kernel<<<bin, 256>>>((const uint32_t *)s->d_ior, (int32_t *)d_iorlog, nin, d_flag);
            if (dim == 3) prep_u32_kernel<3><<<bout, 128>>>(pp, (const int32_t *)d_iorlog, d_tr, (int16_t *)s->d_volume, s->d_translucency, d_flag + 1);
            else          prep_u32_kernel<2><<<bout, 128>>>(pp, (const int32_t *)d_iorlog, d_tr, (int16_t *)s->d_volume, s->d_translucency, d_flag + 1);
        }
        g_launches += 2;
        e = cudaGetLastError();
        e = e == cudaSuccess ? cudaMemcpy(flags_h, d_flag, sizeof flags_h, cudaMemcpyDeviceToHost) : e;
    }
    cudaFree(d_iorlog); cudaFree(d_flag);
    if (!ptrs_on_device) cudaFree(d_tr);
    if (e != cudaSuccess) { vrt_scene_destroy(s); cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? VRT_ERR_NOMEM : VRT_ERR_CUDA, cudaGetErrorString(e)); }
    if (flags_h[0]) { vrt_scene_destroy(s); return fail(VRT_ERR_INVALID, "refraction-index underflow"); }   // image_util.cpp:536-541,607-610
    if (flags_h[1]) { vrt_scene_destroy(s); return fail(VRT_ERR_INVALID, "differention overflow"); }        // image_util.cpp:293-296
    rc = apply_storage(s, flags);
    if (rc) { vrt_scene_destroy(s); return rc; }
    *out = s;
    return VRT_OK;
}

int vrt_scene_info(const vrt_scene *s, int *device, int *dim, uint64_t *bounds, int *diff_dtype,
                   void **d_volume_interleaved, uint32_t **d_translucency, uint64_t *volume_bytes)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    if (device) *device = s->device;
    if (dim) *dim = s->dim;
    if (bounds) for (int d = 0; d < s->dim; ++d) bounds[d] = s->bounds[d];
    if (diff_dtype) *diff_dtype = s->dtype;
    // the staging pointer is only handed out when it IS the reference's layout in the API element type (see the header)
    const bool plain = s->store == s->dtype && !s->bricked && !s->paired && !s->tex;
    if (d_volume_interleaved) *d_volume_interleaved = plain ? s->d_volume : nullptr;
    if (d_translucency) *d_translucency = s->d_translucency;
    if (volume_bytes) *volume_bytes = s->nvox * (s->dim + 1) * elem_size(s->dtype);
    return VRT_OK;
}

static uint64_t storage_bytes(const vrt_scene *s)
{
    if (s->paired) return s->nvox * 32ull;
    if (s->bricked) return s->nb[0] * s->nb[1] * s->nb[2] * 8ull * 4ull * elem_size(s->store);
    return s->nvox * (uint64_t)(s->dim + 1) * elem_size(s->store);
}

int vrt_scene_storage_info(const vrt_scene *s, int *storage_dtype, unsigned *layout_flags, uint64_t *bytes, void **d_storage)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    if (storage_dtype) *storage_dtype = s->store;
    if (layout_flags) *layout_flags = (s->bricked ? VRT_SCENE_LAYOUT_BRICK : 0u) | (s->paired ? VRT_SCENE_LAYOUT_PAIR : 0u) | (s->tex ? VRT_SCENE_LAYOUT_TEXTURE : 0u) |
                                      (s->dtype == VRT_I16 && s->store == VRT_I16 && s->dim == 3 ? VRT_SCENE_KEEP_I16 : 0u) | (s->owns ? 0u : VRT_SCENE_BORROW);
    if (bytes) *bytes = storage_bytes(s);
    if (d_storage) *d_storage = s->d_volume;
    return VRT_OK;
}

int vrt_scene_download(const vrt_scene *s, void *host_volume, uint32_t *host_translucency)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    DeviceGuard g(s->device);
    const size_t bytes = s->nvox * (s->dim + 1) * elem_size(s->dtype);
    if (host_volume && (s->bricked || s->paired || s->store != s->dtype))      // hand out the reference's layout and element type
    {
        void *tmp = nullptr;
        VRT_CUDA(cudaMalloc(&tmp, bytes));
        int rc = linearise(s, tmp, nullptr);
        cudaError_t e = rc == VRT_OK ? cudaMemcpy(host_volume, tmp, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
        cudaFree(tmp);
        if (rc) return rc;
        VRT_CUDA(e);
    }
    else if (host_volume) VRT_CUDA(cudaMemcpy(host_volume, s->d_volume, bytes, cudaMemcpyDeviceToHost));
    if (host_translucency && s->d_translucency) VRT_CUDA(cudaMemcpy(host_translucency, s->d_translucency, s->nvox * 4, cudaMemcpyDeviceToHost));
    return VRT_OK;
}

int vrt_scene_export_device(const vrt_scene *s, void *d_volume_out, uint32_t *d_translucency_out, void *cuda_stream)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d_volume_out) { int rc = linearise(s, d_volume_out, st); if (rc) return rc; }
    if (d_translucency_out && s->d_translucency) VRT_CUDA(cudaMemcpyAsync(d_translucency_out, s->d_translucency, s->nvox * 4, cudaMemcpyDeviceToDevice, st));
    return VRT_OK;
}

int vrt_scene_set_option(vrt_scene *s, int key, int64_t v)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    switch (key)
    {
    case VRT_OPT_KERNEL:
        if (v == 10)       // instrumented copy of the default kernel: allocate / zero the block counters
        {
            DeviceGuard g(s->device);
            if (!s->d_stats) VRT_CUDA(cudaMalloc((void **)&s->d_stats, kStatSlots * sizeof(unsigned long long)));
            VRT_CUDA(cudaMemset(s->d_stats, 0, kStatSlots * sizeof(unsigned long long)));
        }
        else if (v < 0 || v > 6 || v == 4 || v == 5) return fail(VRT_ERR_INVALID, "kernel must be 0, 1, 2, 3, 6 or 10 (4/5/7/8/9 are selected implicitly)");
        s->opt_kernel = v; break;
    case VRT_OPT_BLOCK_THREADS:  if (v < 32 || v > 256 || v % 32) return fail(VRT_ERR_INVALID, "block threads must be a multiple of 32 in [32,256]"); s->opt_block = v; break;
    case VRT_OPT_REFILL:         if (v < 0 || v > 32) return fail(VRT_ERR_INVALID, "refill threshold must be 0..32"); s->opt_refill = v; break;
    case VRT_OPT_CHUNK_RAYS:     if (v < 0) return fail(VRT_ERR_INVALID, "chunk must be >= 0"); s->opt_chunk = v; break;
    case VRT_OPT_STEPS_PER_POLL: if (v < 1 || v > 4096) return fail(VRT_ERR_INVALID, "steps per poll must be 1..4096"); s->opt_poll = v; break;
    case VRT_OPT_MAX_CTAS_PER_SM: if (v < 0 || v > 32) return fail(VRT_ERR_INVALID, "max CTAs per SM must be 0..32"); s->opt_max_ctas = v; break;
    case VRT_OPT_REGION_LOG2:    if (v != 0 && v != -1 && (v < 5 || v > 9)) return fail(VRT_ERR_INVALID, "region log2 must be -1, 0 or 5..9"); s->opt_region = v; break;
    case VRT_OPT_REGION_ROUNDS:  if (v < 1 || v > 256) return fail(VRT_ERR_INVALID, "region rounds must be 1..256"); s->opt_rounds = v; break;   // accepted, unused
    case VRT_OPT_WAVE_LOG2:      if (v != 0 && v != -1 && (v < 3 || v > 8)) return fail(VRT_ERR_INVALID, "wavefront brick log2 must be -1, 0 or 3..8"); s->opt_wave = v; break;
    case VRT_OPT_WAVE_MARGIN:    if (v < 0 || v > 64) return fail(VRT_ERR_INVALID, "wavefront margin must be 0..64"); s->opt_wave_margin = v; break;
    case VRT_OPT_WAVE_CHECK:     if (v < 1 || v > 4096) return fail(VRT_ERR_INVALID, "wavefront steps per check must be 1..4096"); s->opt_wave_check = v; break;
    case VRT_OPT_WAVE_TAIL_PERMILLE: if (v < 0 || v > 1000) return fail(VRT_ERR_INVALID, "wavefront tail must be 0..1000 permille"); s->opt_wave_tail = v; break;
    case VRT_OPT_WAVE_CTAS_PER_SM: if (v < 0 || v > 8) return fail(VRT_ERR_INVALID, "wavefront CTAs per SM must be 0..8"); s->opt_wave_ctas = v; break;
    case VRT_OPT_WAVE_REFILL:    if (v < 1 || v > 32) return fail(VRT_ERR_INVALID, "wavefront refill threshold must be 1..32"); s->opt_wave_refill = v; break;
    case VRT_OPT_ALL_CLEAR_KERNEL: if (v < 0 || v > 1) return fail(VRT_ERR_INVALID, "must be 0 or 1"); s->opt_allclear = v; break;
    default: return fail(VRT_ERR_INVALID, "unknown option");
    }
    return VRT_OK;
}

int vrt_scene_get_option(const vrt_scene *s, int key, int64_t *v)
{
    if (!s || !v) return fail(VRT_ERR_INVALID, "null argument");
    switch (key)
    {
    case VRT_OPT_KERNEL: *v = s->opt_kernel; break;
    case VRT_OPT_BLOCK_THREADS: *v = s->opt_block; break;
    case VRT_OPT_REFILL: *v = s->opt_refill; break;
    case VRT_OPT_CHUNK_RAYS: *v = s->opt_chunk; break;
    case VRT_OPT_STEPS_PER_POLL: *v = s->opt_poll; break;
    case VRT_OPT_MAX_CTAS_PER_SM: *v = s->opt_max_ctas; break;
    case VRT_OPT_REGION_LOG2: *v = s->opt_region; break;
    case VRT_OPT_REGION_ROUNDS: *v = s->opt_rounds; break;
    case VRT_OPT_WAVE_LOG2: *v = s->opt_wave; break;
    case VRT_OPT_WAVE_MARGIN: *v = s->opt_wave_margin; break;
    case VRT_OPT_WAVE_CHECK: *v = s->opt_wave_check; break;
    case VRT_OPT_WAVE_TAIL_PERMILLE: *v = s->opt_wave_tail; break;
    case VRT_OPT_WAVE_CTAS_PER_SM: *v = s->opt_wave_ctas; break;
    case VRT_OPT_WAVE_REFILL: *v = s->opt_wave_refill; break;
    case VRT_OPT_ALL_CLEAR_KERNEL: *v = s->opt_allclear; break;
    case VRT_INFO_ALL_CLEAR: *v = s->all_clear ? 1 : 0; break;
    case VRT_INFO_WAVE_ROUNDS: *v = s->last_wave_rounds_host(); break;
    case VRT_INFO_EMPTY_PERMILLE: *v = (int64_t)(s->flat_fraction * 1000.0 + 0.5); break;
    case VRT_INFO_NUM_SMS: *v = s->num_sms; break;
    case VRT_INFO_STAT_BASE + 0: case VRT_INFO_STAT_BASE + 1: case VRT_INFO_STAT_BASE + 2: case VRT_INFO_STAT_BASE + 3:
    case VRT_INFO_STAT_BASE + 4: case VRT_INFO_STAT_BASE + 5: case VRT_INFO_STAT_BASE + 6: case VRT_INFO_STAT_BASE + 7:
    case VRT_INFO_STAT_BASE + 8: case VRT_INFO_STAT_BASE + 9: case VRT_INFO_STAT_BASE + 10: case VRT_INFO_STAT_BASE + 11:
    {
        *v = 0;
        if (!s->d_stats) break;
        DeviceGuard g(s->device);
        unsigned long long h = 0;
        VRT_CUDA(cudaDeviceSynchronize());
        VRT_CUDA(cudaMemcpy(&h, s->d_stats + (key - VRT_INFO_STAT_BASE), sizeof h, cudaMemcpyDeviceToHost));
        *v = (int64_t)h;
        break;
    }
    default: return fail(VRT_ERR_INVALID, "unknown option");
    }
    return VRT_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------
// multi-GPU replication of a staged scene (reference: the per-device host upload loop cu:676-686)

// a scene on `device` with src's geometry, storage description and options, without buffers
static int clone_empty(const vrt_scene *src, int device, vrt_scene **out)
{
    vrt_scene *s = nullptr;
    int rc = new_scene(&s, device, src->dim, src->bounds, src->dtype);
    if (rc) return rc;
    s->store = src->store; s->bricked = src->bricked; s->flat_fraction = src->flat_fraction; s->all_clear = src->all_clear;
    for (int d = 0; d < 3; ++d) { s->nb[d] = src->nb[d]; s->ior_bounds[d] = src->ior_bounds[d]; }
    s->ior_dtype = src->ior_dtype;
    s->opt_kernel = src->opt_kernel.load(); s->opt_block = src->opt_block.load(); s->opt_refill = src->opt_refill.load();
    s->opt_chunk = src->opt_chunk.load(); s->opt_poll = src->opt_poll.load(); s->opt_max_ctas = src->opt_max_ctas.load();
    s->opt_region = src->opt_region.load(); s->opt_rounds = src->opt_rounds.load();
    s->opt_wave = src->opt_wave.load(); s->opt_wave_margin = src->opt_wave_margin.load(); s->opt_wave_check = src->opt_wave_check.load();
    s->opt_wave_tail = src->opt_wave_tail.load(); s->opt_wave_ctas = src->opt_wave_ctas.load();
    s->opt_wave_refill = src->opt_wave_refill.load(); s->opt_allclear = src->opt_allclear.load();
    *out = s;
    return VRT_OK;
}

struct ReplSeg { const char *from; char *to; size_t bytes; };

static uint64_t ior_bytes(const vrt_scene *s)
{
    uint64_t n = 1;
    for (int d = 0; d < s->dim; ++d) n *= s->ior_bounds[d];
    return n * 4;
}

// allocate the buffers of a replica and list the (source, destination, bytes) segments to fill
static int alloc_replica(const vrt_scene *src, vrt_scene *dst, const vrt_scene *from, std::vector<ReplSeg> &segs)
{
    DeviceGuard g(dst->device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    VRT_CUDA(cudaMalloc(&dst->d_volume, storage_bytes(src)));
    dst->owns = true;
    segs.push_back({(const char *)from->d_volume, (char *)dst->d_volume, (size_t)storage_bytes(src)});
    if (src->d_translucency)
    {
        VRT_CUDA(cudaMalloc((void **)&dst->d_translucency, src->nvox * 4));
        segs.push_back({(const char *)from->d_translucency, (char *)dst->d_translucency, (size_t)(src->nvox * 4)});
    }
    if (src->d_ior)
    {
        VRT_CUDA(cudaMalloc(&dst->d_ior, ior_bytes(src)));
        dst->owns_ior = true;
        segs.push_back({(const char *)from->d_ior, (char *)dst->d_ior, (size_t)ior_bytes(src)});
    }
    return VRT_OK;
}

static void enable_peer(int a, int b)
{
    if (a == b) return;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can)
    {
        DeviceGuard g(a);
        cudaDeviceEnablePeerAccess(b, 0);     // cudaErrorPeerAccessAlreadyEnabled is fine
    }
    cudaGetLastError();                        // without peer access cudaMemcpyPeerAsync stages through the host: slower, still correct
}

// ---- NCCL, loaded at run time -------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };
struct NcclApi
{
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;    // optional: the rendezvous of vrt_scene_broadcast
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    bool ok = false;
};
NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("VRT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names)
        {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);   // a process that already has NCCL mapped (PyTorch) gets that copy
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = (int (*)(NcclId *))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (int (*)(void *))dlsym(api.handle, "ncclCommDestroy");
        api.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.handle, "ncclBroadcast");
        api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
        api.GetErrorString = (const char *(*)(int))dlsym(api.handle, "ncclGetErrorString");
        api.GetVersion = (int (*)(int *))dlsym(api.handle, "ncclGetVersion");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.GetErrorString;
    });
    return api;
}
} // namespace

struct vrt_comm
{
    int device = 0, rank = 0, world = 1;
    void *comm = nullptr;
    cudaStream_t stream = nullptr;
    char *d_hdr = nullptr;       // 256-byte device staging for the scene header
};

#define VRT_NCCL(call)                                                                                      \
    do {                                                                                                    \
        int r__ = (call);                                                                                   \
        if (r__ != 0) return fail(VRT_ERR_CUDA, std::string("NCCL: ") + nccl().GetErrorString(r__) + " (" #call ")"); \
    } while (0)

// what vrt_scene_broadcast sends ahead of the payload
struct SceneHeader
{
    uint32_t magic;
    int32_t  dim, dtype, store, bricked, has_tr, has_ior, ior_dtype;
    uint64_t bounds[3], nb[3], ior_bounds[3];
    double   flat_fraction;
    int64_t  opt[8];
    int32_t  all_clear, pad_;
};
static_assert(sizeof(SceneHeader) <= 256, "header must fit the staging buffer");

extern "C" {

int vrt_scene_replicate(const vrt_scene *src, int n, const int *devices, vrt_scene **out, double *seconds)
{
    if (seconds) *seconds = 0.0;
    if (!src || n < 0 || (n && (!devices || !out))) return fail(VRT_ERR_INVALID, "null argument");
    for (int i = 0; i < n; ++i) out[i] = nullptr;
    if (n == 0) return VRT_OK;
    if (src->tex || src->paired) return fail(VRT_ERR_UNSUPPORTED, "study layouts are not replicated");
    int count = 0;
    VRT_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < n; ++i) if (devices[i] < 0 || devices[i] >= count) return fail(VRT_ERR_INVALID, "no such CUDA device");

    // the chain: src -> devices[0] -> devices[1] -> ...; hop j copies from chain[j] to chain[j+1]
    std::vector<const vrt_scene *> chain(1, src);
    std::vector<std::vector<ReplSeg>> segs(n);
    int rc = VRT_OK;
    for (int i = 0; i < n && rc == VRT_OK; ++i)
    {
        rc = clone_empty(src, devices[i], &out[i]);
        if (rc == VRT_OK) rc = alloc_replica(src, out[i], chain.back(), segs[i]);
        if (rc == VRT_OK) chain.push_back(out[i]);
    }
    auto cleanup = [&]() { for (int i = 0; i < n; ++i) { vrt_scene_destroy(out[i]); out[i] = nullptr; } };
    if (rc) { const std::string msg = g_last_error; cleanup(); return fail(rc, msg); }
    for (int j = 0; j < n; ++j) { enable_peer(chain[j]->device, chain[j + 1]->device); enable_peer(chain[j + 1]->device, chain[j]->device); }

    constexpr size_t kSlice = 64ull << 20;
    std::vector<cudaStream_t> st(n, nullptr);
    std::vector<std::vector<cudaEvent_t>> ev(n);
    cudaError_t e = cudaSuccess;
    size_t nslices = 0;
    for (const ReplSeg &g : segs[0]) nslices += (g.bytes + kSlice - 1) / kSlice;
    for (int j = 0; j < n && e == cudaSuccess; ++j)
    {
        DeviceGuard g(chain[j]->device);                       // the sender pushes
        e = cudaStreamCreateWithFlags(&st[j], cudaStreamNonBlocking);
        if (j + 1 < n) { ev[j].resize(nslices, nullptr); for (size_t k = 0; k < nslices && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[j][k], cudaEventDisableTiming); }
    }
    { DeviceGuard g(src->device); if (e == cudaSuccess) e = cudaDeviceSynchronize(); }
    const auto t0 = std::chrono::steady_clock::now();
    size_t k = 0;
    for (size_t sg = 0; sg < segs[0].size() && e == cudaSuccess; ++sg)
        for (size_t off = 0; off < segs[0][sg].bytes && e == cudaSuccess; off += kSlice, ++k)
        {
            const size_t len = std::min(kSlice, segs[0][sg].bytes - off);
            for (int j = 0; j < n && e == cudaSuccess; ++j)    // slice k travels down the chain; hop j waits for hop j-1's copy of the same slice
            {
                DeviceGuard g(chain[j]->device);
                if (j > 0) e = cudaStreamWaitEvent(st[j], ev[j - 1][k], 0);
                if (e == cudaSuccess) e = cudaMemcpyPeerAsync(segs[j][sg].to + off, chain[j + 1]->device, segs[j][sg].from + off, chain[j]->device, len, st[j]);
                if (e == cudaSuccess && j + 1 < n) e = cudaEventRecord(ev[j][k], st[j]);
            }
        }
    for (int j = 0; j < n; ++j) if (st[j]) { cudaError_t e2 = cudaStreamSynchronize(st[j]); if (e == cudaSuccess) e = e2; }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int j = 0; j < n; ++j)
    {
        for (cudaEvent_t x : ev[j]) if (x) cudaEventDestroy(x);
        if (st[j]) cudaStreamDestroy(st[j]);
    }
    if (e != cudaSuccess) { cleanup(); cudaGetLastError(); return fail(VRT_ERR_CUDA, std::string("replication failed: ") + cudaGetErrorString(e)); }
    g_launches += 0;
    if (seconds) *seconds = dt;
    return VRT_OK;
}

int vrt_comm_unique_id(void *id)
{
    if (!id) return fail(VRT_ERR_INVALID, "id is null");
    if (!nccl().ok) return fail(VRT_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded");
    NcclId u;
    VRT_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return VRT_OK;
}

int vrt_comm_create(vrt_comm **out, int device, int rank, int world, const void *id)
{
    if (!out || !id) return fail(VRT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(VRT_ERR_INVALID, "bad rank / world");
    if (!nccl().ok) return fail(VRT_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded");
    DeviceGuard g(device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    vrt_comm *c = new vrt_comm();
    c->device = device; c->rank = rank; c->world = world;
    NcclId u;
    memcpy(&u, id, sizeof u);
    int r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; return fail(VRT_ERR_CUDA, std::string("NCCL: ") + nccl().GetErrorString(r) + " (ncclCommInitRank)"); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    e = e == cudaSuccess ? cudaMalloc((void **)&c->d_hdr, 512) : e;      // 256-byte scene header + a word for the rendezvous
    e = e == cudaSuccess ? cudaMemset(c->d_hdr, 0, 512) : e;
    // warm-up: NCCL connects its channels lazily on the first collective; do that here, not inside the first scene broadcast
    void *warm = nullptr;
    e = e == cudaSuccess ? cudaMalloc(&warm, 8u << 20) : e;
    if (e == cudaSuccess)
    {
        r = nccl().Broadcast(c->d_hdr, c->d_hdr, 4, 0, 0, c->comm, c->stream);
        if (r == 0) r = nccl().Broadcast(warm, warm, 8u << 20, 0, 0, c->comm, c->stream);
        if (r == 0 && nccl().AllReduce) r = nccl().AllReduce(c->d_hdr + 256, c->d_hdr + 256, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, c->stream);
        e = cudaStreamSynchronize(c->stream);
    }
    if (warm) cudaFree(warm);
    if (e != cudaSuccess || r != 0)
    {
        const std::string msg = r != 0 ? std::string("NCCL: ") + nccl().GetErrorString(r) : std::string(cudaGetErrorString(e));
        vrt_comm_destroy(c);
        cudaGetLastError();
        return fail(VRT_ERR_CUDA, msg);
    }
    *out = c;
    return VRT_OK;
}

int vrt_comm_destroy(vrt_comm *c)
{
    if (!c) return VRT_OK;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    if (c->d_hdr) cudaFree(c->d_hdr);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return VRT_OK;
}

int vrt_scene_broadcast(vrt_comm *c, int root, vrt_scene *src, vrt_scene **out, double *seconds)
{
    if (seconds) *seconds = 0.0;
    if (!c || !out) return fail(VRT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (root < 0 || root >= c->world) return fail(VRT_ERR_INVALID, "bad root");
    const bool is_root = c->rank == root;
    if (is_root && !src) return fail(VRT_ERR_INVALID, "the root rank needs a scene");
    if (is_root && (src->tex || src->paired)) return fail(VRT_ERR_UNSUPPORTED, "study layouts are not replicated");
    if (is_root && src->device != c->device) return fail(VRT_ERR_INVALID, "the scene is not on the communicator's device");
    DeviceGuard g(c->device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    SceneHeader h;
    memset(&h, 0, sizeof h);
    if (is_root)
    {
        h.magic = 0x56525442u; h.dim = src->dim; h.dtype = src->dtype; h.store = src->store; h.bricked = src->bricked;
        h.has_tr = src->d_translucency != nullptr; h.has_ior = src->d_ior != nullptr; h.ior_dtype = src->ior_dtype;
        for (int d = 0; d < 3; ++d) { h.bounds[d] = src->bounds[d]; h.nb[d] = src->nb[d]; h.ior_bounds[d] = src->ior_bounds[d]; }
        h.flat_fraction = src->flat_fraction; h.all_clear = src->all_clear ? 1 : 0; h.pad_ = 0;
        h.opt[0] = src->opt_kernel; h.opt[1] = src->opt_block; h.opt[2] = src->opt_refill; h.opt[3] = src->opt_chunk;
        h.opt[4] = src->opt_poll; h.opt[5] = src->opt_max_ctas; h.opt[6] = src->opt_region; h.opt[7] = src->opt_rounds;
        VRT_CUDA(cudaMemcpyAsync(c->d_hdr, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    }
    VRT_NCCL(nccl().Broadcast(c->d_hdr, c->d_hdr, 256, 0, root, c->comm, c->stream));
    if (!is_root) VRT_CUDA(cudaMemcpyAsync(&h, c->d_hdr, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(cudaStreamSynchronize(c->stream));
    if (h.magic != 0x56525442u) return fail(VRT_ERR_INVALID, "scene header did not arrive (root has no scene?)");

    vrt_scene *s = src;
    if (!is_root)
    {
        vrt_scene proto;                       // geometry carrier for clone_empty / alloc_replica
        proto.dim = h.dim; proto.dtype = h.dtype; proto.store = h.store; proto.bricked = h.bricked != 0; proto.flat_fraction = h.flat_fraction; proto.all_clear = h.all_clear != 0;
        proto.ior_dtype = h.ior_dtype; proto.nvox = 1;
        for (int d = 0; d < 3; ++d) { proto.bounds[d] = h.bounds[d]; proto.nb[d] = h.nb[d]; proto.ior_bounds[d] = h.ior_bounds[d]; }
        for (int d = 0; d < h.dim; ++d) proto.nvox *= h.bounds[d];
        proto.opt_kernel = h.opt[0]; proto.opt_block = h.opt[1]; proto.opt_refill = h.opt[2]; proto.opt_chunk = h.opt[3];
        proto.opt_poll = h.opt[4]; proto.opt_max_ctas = h.opt[5]; proto.opt_region = h.opt[6]; proto.opt_rounds = h.opt[7];
        proto.d_translucency = h.has_tr ? (uint32_t *)1 : nullptr;      // only tested for presence
        proto.d_ior = h.has_ior ? (void *)1 : nullptr;
        int rc = clone_empty(&proto, c->device, &s);
        std::vector<ReplSeg> unused;
        if (rc == VRT_OK) rc = alloc_replica(&proto, s, &proto, unused);
        proto.d_translucency = nullptr; proto.d_ior = nullptr;
        if (rc) { const std::string msg = g_last_error; vrt_scene_destroy(s); return fail(rc, msg); }
    }
    // Rendezvous: the receiving ranks have just allocated their replica (a cudaMalloc of the whole scene takes tens of milliseconds).  Without
    // it the root's broadcast kernel starts at once and waits for them on the device, so the root's `seconds` -- and the maximum over the ranks
    // -- would be the receivers' allocation time plus the transfer (seen as 181 instead of ~600 GB/s on 2 GPUs).  A 4-byte all-reduce: every
    // rank has to arrive.
    int r = 0;
    cudaError_t e = cudaSuccess;
    if (nccl().AllReduce)
    {
        r = nccl().AllReduce(c->d_hdr + 256, c->d_hdr + 256, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, c->stream);
        e = r == 0 ? cudaStreamSynchronize(c->stream) : e;
    }
    cudaEvent_t a = nullptr, b = nullptr;
    e = e == cudaSuccess && r == 0 ? cudaEventCreate(&a) : e;
    e = e == cudaSuccess && r == 0 ? cudaEventCreate(&b) : e;
    e = e == cudaSuccess && r == 0 ? cudaEventRecord(a, c->stream) : e;
    if (e == cudaSuccess && r == 0)
    {
        r = nccl().Broadcast(s->d_volume, s->d_volume, (size_t)storage_bytes(s), 0, root, c->comm, c->stream);       // in place on every rank
        if (r == 0 && h.has_tr) r = nccl().Broadcast(s->d_translucency, s->d_translucency, (size_t)(s->nvox * 4), 0, root, c->comm, c->stream);
        if (r == 0 && h.has_ior) r = nccl().Broadcast(s->d_ior, s->d_ior, (size_t)ior_bytes(s), 0, root, c->comm, c->stream);
        e = cudaEventRecord(b, c->stream);
        e = e == cudaSuccess ? cudaStreamSynchronize(c->stream) : e;
    }
    float ms = 0.0f;
    if (e == cudaSuccess && r == 0) cudaEventElapsedTime(&ms, a, b);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (e != cudaSuccess || r != 0)
    {
        const std::string msg = r != 0 ? std::string("NCCL: ") + nccl().GetErrorString(r) + " (ncclBroadcast)" : std::string(cudaGetErrorString(e));
        if (!is_root) vrt_scene_destroy(s);
        cudaGetLastError();
        return fail(VRT_ERR_CUDA, msg);
    }
    if (seconds) *seconds = ms * 1e-3;
    *out = s;
    return VRT_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------
// launch

template <typename VoxT, bool DIR_I16, bool LIVE, bool PATH, int KVER>
static cudaError_t launch3(const vrt_scene *s, const MarchParams &p, int block, cudaStream_t st)
{
    auto kern = march3_kernel<VoxT, DIR_I16, LIVE, PATH, KVER>;
    static std::atomic<unsigned long long> carved{0};  // per device: the marcher uses no shared memory, give the unified array to L1
    const unsigned long long bit = 1ull << (s->device & 63);
    if (!(carved.fetch_or(bit) & bit))
    {
        // every CTA reserves 1 KB of shared memory: the smallest carve-out (MaxL1) admits 8 CTAs per SM, the 9-CTA instantiation needs the next one
        int carve = MarchBounds<VoxT, LIVE, PATH, KVER>::kNine ? 8 : (int)cudaSharedmemCarveoutMaxL1;
        if (const char *e = std::getenv("VRT_CARVEOUT_PCT")) carve = std::atoi(e);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    unsigned grid;
    if (p.refill == 0) grid = (unsigned)((p.n + block - 1) / block);
    else
    {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        const int cap = (int)s->opt_max_ctas.load();
        if (cap > 0 && cap < per_sm) per_sm = cap;
        if (MarchBounds<VoxT, LIVE, PATH, KVER>::kNine && cap == 0 && per_sm == 9)
        {
            // a batch of equally long rays (config 5: every ray runs the cap) takes ceil(n / resident rays) waves: with few waves one CTA
            // fewer per SM can mean one partial wave fewer (a 2 M-ray shard: 14 waves of 8 CTAs against 13 of 9, each 9/8/1.03 as long)
            auto cost = [&](int c) {
                const unsigned long long per_wave = (unsigned long long)s->num_sms * c * block;
                return (double)((p.n + per_wave - 1) / per_wave) * c / (c == 9 ? 1.03 : 1.0);
            };
            if (cost(8) < cost(9)) per_sm = 8;
        }
        const unsigned long long want = (p.n + block - 1) / block;
        grid = (unsigned)std::min<unsigned long long>((unsigned long long)per_sm * s->num_sms, want);
    }
    kern<<<grid, block, 0, st>>>(p);
    ++g_launches;
    return cudaGetLastError();
}

template <typename VoxT, bool DIR_I16, bool LIVE>
static cudaError_t launch3_k(const vrt_scene *s, const MarchParams &p, bool path, int kver, int block, cudaStream_t st)
{
#ifdef VRT_STUDY
    if (path && kver == 7) return launch3<VoxT, DIR_I16, LIVE, true, 7>(s, p, block, st);
#endif
    if (path) return kver == 8 ? launch3<VoxT, DIR_I16, LIVE, true, 8>(s, p, block, st)
                               : launch3<VoxT, DIR_I16, LIVE, true, 2>(s, p, block, st);   // polyline output is store-bound: one variant per rounding
    switch (kver)
    {
    case 1: return launch3<VoxT, DIR_I16, LIVE, false, 1>(s, p, block, st);
    case 2: return launch3<VoxT, DIR_I16, LIVE, false, 2>(s, p, block, st);
    case 4: return launch3<VoxT, DIR_I16, LIVE, false, 4>(s, p, block, st);
    case 6: return launch3<VoxT, DIR_I16, LIVE, false, 6>(s, p, block, st);
    case 8: return launch3<VoxT, DIR_I16, LIVE, false, 8>(s, p, block, st);
#ifdef VRT_STUDY
    case 5: return launch3<VoxT, DIR_I16, LIVE, false, 5>(s, p, block, st);
    case 7: return launch3<VoxT, DIR_I16, LIVE, false, 7>(s, p, block, st);
#endif
    case 9: return launch3<VoxT, DIR_I16, LIVE, false, 9>(s, p, block, st);
    case 11: return launch3<VoxT, DIR_I16, LIVE, false, 11>(s, p, block, st);
    case 10: return launch3<float, false, false, false, 10>(s, p, block, st);      // enqueue_march only selects it for this instantiation
    default: return launch3<VoxT, DIR_I16, LIVE, false, 3>(s, p, block, st);
    }
}

template <typename VoxT, bool DIR_I16, bool LIVE>
static cudaError_t launch2_k(const MarchParams &p, bool path, bool hostr, int block, cudaStream_t st)
{
    const unsigned grid = (unsigned)((p.n + block - 1) / block);
    if (hostr)
    {
        if (path) march2_kernel<VoxT, DIR_I16, LIVE, true, true><<<grid, block, 0, st>>>(p);
        else      march2_kernel<VoxT, DIR_I16, LIVE, false, true><<<grid, block, 0, st>>>(p);
    }
    else
    {
        if (path) march2_kernel<VoxT, DIR_I16, LIVE, true, false><<<grid, block, 0, st>>>(p);
        else      march2_kernel<VoxT, DIR_I16, LIVE, false, false><<<grid, block, 0, st>>>(p);
    }
    ++g_launches;
    return cudaGetLastError();
}

template <typename VoxT>
static cudaError_t launch_vox(const vrt_scene *s, const MarchParams &p, bool dir_i16, bool live, bool path, int kver, int block, cudaStream_t st)
{
    if (s->dim == 3)
    {
        if (dir_i16) return live ? launch3_k<VoxT, true, true>(s, p, path, kver, block, st) : launch3_k<VoxT, true, false>(s, p, path, kver, block, st);
        return live ? launch3_k<VoxT, false, true>(s, p, path, kver, block, st) : launch3_k<VoxT, false, false>(s, p, path, kver, block, st);
    }
    const bool hostr = kver == 8;
    if (dir_i16) return live ? launch2_k<VoxT, true, true>(p, path, hostr, block, st) : launch2_k<VoxT, true, false>(p, path, hostr, block, st);
    return live ? launch2_k<VoxT, false, true>(p, path, hostr, block, st) : launch2_k<VoxT, false, false>(p, path, hostr, block, st);
}

// ---- wavefront mode (vrt_wave.cuh): ONE cooperative launch does bucket passes + marching, round by round -------------------
// rays per brick from which the all-clear wavefront kernel runs 4 instead of 3 CTAs per SM (VRT_WAVE_DENSE_RAYS_PER_BRICK overrides: tuning)
static int64_t wave_dense_threshold()
{
    static const int64_t v = [] { const char *e = std::getenv("VRT_WAVE_DENSE_RAYS_PER_BRICK"); return e ? (int64_t)std::atoll(e) : (int64_t)768; }();
    return v;
}
template <typename VoxT, bool DIR_I16, bool LIVE, bool ALLCLEAR = false>
static cudaError_t launch_wave(const vrt_scene *s, WaveParams &wp, cudaStream_t st)
{
    auto kern = march3_wave_kernel<VoxT, DIR_I16, LIVE, ALLCLEAR>;
    static std::atomic<unsigned long long> carved{0};  // per device: next to no shared memory in use, give the unified array to L1
    const unsigned long long bit = 1ull << (s->device & 63);
    if (!(carved.fetch_or(bit) & bit)) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWaveThreads, 0);
    if (e != cudaSuccess) return e;
    const int cap = (int)s->opt_wave_ctas.load();
    if (cap > 0 && cap < per_sm) per_sm = cap;
    // The rays in flight (resident threads) spread over [threads / rays per brick] bricks, whose boxes L2 has to hold: a fourth CTA per SM
    // pays when bricks are dense (config 4, 8 M rays over 4096 bricks = 2048 rays per brick: 95.5 -> 101.2 G ray-steps/s) and costs when
    // they are sparse (1 M rays, 256 per brick: 59.6 -> 52.8; profiles/r02_c4_sweep_allclear.log)
    else if (cap == 0 && per_sm > 3 && wp.m.n / std::max<uint64_t>(wp.K, 1) < (uint64_t)wave_dense_threshold()) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    const unsigned grid = (unsigned)(per_sm * s->num_sms);          // cooperative: every CTA must be resident
    void *args[] = {(void *)&wp};
    ++g_launches;
    return cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(kWaveThreads), args, 0, st);
}

static bool wave_geometry(const vrt_scene *s, int k, uint32_t nb[3], uint64_t *K)
{
    const uint64_t e = 1ull << k;
    uint64_t tot = 1;
    for (int d = 0; d < 3; ++d) { nb[d] = (uint32_t)((s->bounds[d] + e - 1) / e); tot *= nb[d]; }
    *K = tot;
    return tot < (1ull << 24);           // histogram / cursors / work items stay small next to the ray state
}

static int enqueue_march_wave(const vrt_scene *s, const MarchParams &mp, bool di16, bool live, cudaStream_t st, int k)
{
    const uint64_t n = mp.n;
    uint32_t nb[3]; uint64_t K = 0;
    if (!wave_geometry(s, k, nb, &K)) return fail(VRT_ERR_INVALID, "too many bricks: raise VRT_OPT_WAVE_LOG2");
    if (n >= (1ull << 31)) return fail(VRT_ERR_INVALID, "wavefront mode takes at most 2^31-1 rays per call");
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device);
    if (!coop) return fail(VRT_ERR_UNSUPPORTED, "device does not support cooperative launches");
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_pos = 0, o_dir = o_pos + al(n * 12), o_it = o_dir + al(n * 12), o_light = o_it + al(n * 4), o_key = o_light + al(n * 4);
    const size_t o_o0 = o_key + al(n * 4), o_o1 = o_o0 + al(n * 4), o_h0 = o_o1 + al(n * 4), o_h1 = o_h0 + al(K * 4), o_off = o_h1 + al(K * 4);
    const size_t o_part = o_off + al((K + 1) * 4), o_ctl = o_part + al(4096 * 4), total = o_ctl + 256;
    char *ws = nullptr;
    VRT_CUDA(pool_alloc((void **)&ws, total, s->device, st));
    WaveParams wp;
    wp.m = mp;
    wp.m.counter = nullptr; wp.m.refill = 0;
    wp.st_pos = (uint32_t *)(ws + o_pos); wp.st_dir = (float *)(ws + o_dir); wp.st_it = (uint32_t *)(ws + o_it); wp.st_light = (uint32_t *)(ws + o_light);
    wp.key_of_ray = (uint32_t *)(ws + o_key);
    wp.order[0] = (uint32_t *)(ws + o_o0); wp.order[1] = (uint32_t *)(ws + o_o1);
    wp.hist[0] = (uint32_t *)(ws + o_h0); wp.hist[1] = (uint32_t *)(ws + o_h1);
    wp.bin_off = (uint32_t *)(ws + o_off);
    wp.partial = (uint32_t *)(ws + o_part);
    wp.ctl = (uint32_t *)(ws + o_ctl);
    wp.log2_brick = k; wp.margin = (uint32_t)s->opt_wave_margin.load();
    wp.nby = nb[1]; wp.nbz = nb[2]; wp.K = (uint32_t)K;
    wp.tail_rays = (uint32_t)std::min<uint64_t>(n, std::max<uint64_t>(n * (uint64_t)s->opt_wave_tail.load() / 1000, 1024));
    wp.steps_per_check = (int)s->opt_wave_check.load();
    wp.refill = (uint32_t)s->opt_wave_refill.load();
    wp.max_rounds = mp.iterations + 8u < mp.iterations ? 0xFFFFFFFFu : mp.iterations + 8u;
    cudaError_t err = cudaMemsetAsync(wp.ctl, 0, 256, st);
    if (err == cudaSuccess)
    {
        // a scene without any possibly opaque voxel (VRT_INFO_ALL_CLEAR; shipped translucency behaviour): the variant that keeps no channel 3
        // in its cell cache -- 64 registers, 4 resident CTAs per SM instead of 3 (same bits: vrt_wave.cuh)
        const bool allclear = s->all_clear && s->opt_allclear.load() != 0 && !live;
        if (allclear && s->store == VRT_F32)
            err = di16 ? launch_wave<float, true, false, true>(s, wp, st) : launch_wave<float, false, false, true>(s, wp, st);
        else if (allclear)
            err = di16 ? launch_wave<int16_t, true, false, true>(s, wp, st) : launch_wave<int16_t, false, false, true>(s, wp, st);
        else if (s->store == VRT_F32)
            err = di16 ? (live ? launch_wave<float, true, true>(s, wp, st) : launch_wave<float, true, false>(s, wp, st))
                       : (live ? launch_wave<float, false, true>(s, wp, st) : launch_wave<float, false, false>(s, wp, st));
        else
            err = di16 ? (live ? launch_wave<int16_t, true, true>(s, wp, st) : launch_wave<int16_t, true, false>(s, wp, st))
                       : (live ? launch_wave<int16_t, false, true>(s, wp, st) : launch_wave<int16_t, false, false>(s, wp, st));
    }
    if (err == cudaSuccess)
    {
        if (!s->d_wave_info && cudaMalloc((void **)&s->d_wave_info, kCtlWords * 4) != cudaSuccess) { cudaGetLastError(); s->d_wave_info = nullptr; }
        if (s->d_wave_info) cudaMemcpyAsync(s->d_wave_info, wp.ctl, kCtlWords * 4, cudaMemcpyDeviceToDevice, st);
    }
    cudaFreeAsync(ws, st);
    VRT_CUDA(err);
    return VRT_OK;
}

static int validate_trace(const vrt_scene *s, uint64_t n, const void *pos, const void *dir, int dir_dtype, const float *invscale,
                          uint32_t iterations, unsigned flags, const void *epos, const void *edir, const void *eit, const void *light, const void *path)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    if (dir_dtype != VRT_F32 && dir_dtype != VRT_I16) return fail(VRT_ERR_INVALID, "dir_dtype must be VRT_F32 or VRT_I16");
    if (!invscale) return fail(VRT_ERR_INVALID, "invscale is null");
    if (n && (!pos || !dir || !epos || !edir || !eit || !light)) return fail(VRT_ERR_INVALID, "null ray buffer");
    if ((flags & VRT_TRACE_PATHS) && n && !path) return fail(VRT_ERR_INVALID, "VRT_TRACE_PATHS needs a path buffer");
    // iterations == 0 makes the reference's counter wrap (cu:333: path[--iterations]) and march for 2^32-1 steps
    if (iterations == 0) return fail(VRT_ERR_INVALID, "iterations must be >= 1");
    if ((flags & VRT_TRACE_LIVE_TRANSLUCENCY) && !s->d_translucency) return fail(VRT_ERR_INVALID, "scene has no translucency plane");
    if ((flags & VRT_TRACE_PATHS) && (s->bricked || s->tex)) return fail(VRT_ERR_UNSUPPORTED, "path output is not implemented for VRT_SCENE_LAYOUT_BRICK / _TEXTURE scenes");
    if (n >= (1ull << 40)) return fail(VRT_ERR_INVALID, "too many rays");
    return VRT_OK;
}

// which marcher enqueue_march uses: the single launch (default), the wavefront marcher, or -- gate != null -- whichever the
// device-side probe flag selects (gate_want: the flag value this launch runs for)
struct MarchMode
{
    int wave_log2 = 0;
    const uint32_t *gate = nullptr;
    uint32_t gate_want = 0;
};

// enqueue one marcher launch on `st`; `scratch` is a 16-byte device scratch (zeroed here: [0] = refill counter, [1] = cap flag,
// read back by vrt_trace) or null for static mode without a cap flag
static int enqueue_march(const vrt_scene *s, uint64_t n, const uint32_t *d_pos, const void *d_dir, int dir_dtype, const float *invscale,
                         uint32_t minb, uint32_t iterations, unsigned flags, uint32_t *d_epos, void *d_edir, uint32_t *d_eit,
                         uint32_t *d_light, uint32_t *d_path, unsigned long long *scratch, cudaStream_t st, const MarchMode &mode)
{
    unsigned long long *counter = scratch && s->opt_refill.load() > 0 && s->dim == 3 ? scratch : nullptr;
    if (n == 0) return VRT_OK;
    MarchParams p;
    p.volume = s->d_volume; p.translucency = s->d_translucency;
    p.vol_bytes = storage_bytes(s); p.nvox = s->nvox;
    p.by = (uint32_t)s->bounds[1]; p.bz = (uint32_t)s->bounds[2];
    p.limx = (uint32_t)((s->bounds[0] - 1) & 0xFFFF); p.limy = (uint32_t)((s->bounds[1] - 1) & 0xFFFF); p.limz = (uint32_t)((s->bounds[2] - 1) & 0xFFFF);
    p.limx16 = p.limx << 16; p.limy16 = p.limy << 16; p.limz16 = p.limz << 16;
    p.invx = invscale[0]; p.invy = invscale[1]; p.invz = s->dim == 3 ? invscale[2] : 0.0f;
    p.iterations = iterations; p.min_brightness = minb; p.n = n;
    p.pos = d_pos; p.dir = d_dir; p.epos = d_epos; p.edir = d_edir; p.eit = d_eit; p.light = d_light; p.path = d_path;
    p.steps_per_poll = (int)s->opt_poll.load();
    {
        const unsigned long long vb = s->paired ? 32ull : (s->dim + 1) * elem_size(s->store);
        p.row1 = (unsigned long long)p.bz * vb; p.row2 = (unsigned long long)(uint32_t)(p.by * p.bz) * vb; p.row3 = (unsigned long long)(uint32_t)((p.by + 1u) * p.bz) * vb;
    }
    p.pair = s->paired ? 1 : 0; p.brick = s->bricked ? 1 : 0;
    {
        // |dir|^2 range of the fast loop (see div_is_fast_in): [max(2^-95, m^2 * 2^17 * 1.01), 2^97), m = max |invscale|; an unusable m
        // (NaN, inf, huge) leaves an empty range and every step goes through the generic code
        const float m = std::fmax(std::fabs(p.invx), std::fmax(std::fabs(p.invy), std::fabs(p.invz)));
        float lo = m * m * 131072.0f * 1.01f;
        uint32_t lo_bits = 0x70000000u;
        if (lo == lo && lo < 0x1p97f) { if (lo < 0x1p-95f) lo = 0x1p-95f; memcpy(&lo_bits, &lo, 4); }
        p.dot_lo = lo_bits; p.dot_span = 0x70000000u - lo_bits;
    }
    p.one[0] = p.one[1] = 1.0f; p.zero = 0;
    p.refill = counter ? (int)s->opt_refill.load() : 0;
    p.counter = counter;
    p.cap_flag = scratch ? (uint32_t *)(scratch + 1) : nullptr;
    p.mode_flag = mode.gate; p.mode_want = mode.gate_want;
    int kver = (int)s->opt_kernel.load();
    if (kver == 0) kver = 3;   // 6 (empty-space fast path) stays opt-in: it wins on coherent bundles through mostly empty volumes
                               // (config 1: 25x, config 2: +7 %) and loses where flat and curved cells mix inside a warp
    if (s->bricked) kver = 4;
    if (s->tex) kver = 5;
    if (s->paired) kver = 7;
    p.tex = s->tex;
    p.nby = (uint32_t)s->nb[1]; p.nbz = (uint32_t)s->nb[2];
    const int block = (int)s->opt_block.load();
    const bool live = flags & VRT_TRACE_LIVE_TRANSLUCENCY, path = flags & VRT_TRACE_PATHS, di16 = dir_dtype == VRT_I16;
    if (kver == 10)    // the instrumented copy exists for the default instantiation only; anything else runs the default kernel
    {
        const bool fits = s->dim == 3 && s->store == VRT_F32 && !di16 && !live && !path && !s->bricked && !s->tex && !s->paired &&
                          p.invx == 1.0f && p.invy == 1.0f && p.invz == 1.0f && s->d_stats;
        if (!fits) kver = 3;
    }
    p.stats = kver == 10 ? s->d_stats : nullptr;
    if (kver == 3 && !path && p.invx == 1.0f && p.invy == 1.0f && p.invz == 1.0f) kver = 9;   // unit invscale: two multiplies fewer per step, same bits
    if (kver == 9 && s->all_clear && s->opt_allclear.load() != 0) kver = 11;                  // ... and no voxel that could make a sample opaque: no per-cell test
    if (kver == 11 && s->store == VRT_F32 && !(flags & VRT_TRACE_LIVE_TRANSLUCENCY) && s->opt_block.load() > 128) kver = 9;   // that instantiation is compiled for 128-thread CTAs (MarchBounds)
    const bool hostr = flags & VRT_TRACE_ROUND_HOST;
    if (hostr)
    {
        if (s->bricked || s->tex || s->paired) return fail(VRT_ERR_UNSUPPORTED, "VRT_TRACE_ROUND_HOST needs the linear layout");
        kver = 8;
    }
    if (mode.wave_log2 > 0 && s->dim == 3 && !path && !s->tex && !s->bricked && !s->paired && !hostr)
    {
        // the gated pair shares one scratch: the single-launch marcher (enqueued first) has zeroed it; a stand-alone wavefront launch does it here
        if (scratch && !mode.gate) VRT_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(unsigned long long), st));
        return enqueue_march_wave(s, p, di16, live, st, mode.wave_log2);       // in-place calls are fine: phase 0 has read every start buffer before the first result is written
    }
    if (scratch) VRT_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(unsigned long long), st));
    cudaError_t e = s->store == VRT_F32 ? launch_vox<float>(s, p, di16, live, path, kver, block, st)
                                        : launch_vox<int16_t>(s, p, di16, live, path, kver, block, st);
    VRT_CUDA(e);
    return VRT_OK;
}

// Host-side coherence probe for vrt_trace: are neighbouring rays of the batch neighbours in space?  Samples up to 4096 pairs
// (i, i+1): a pair is incoherent when its start positions are more than 4 voxels apart or its directions differ by more than
// ~25 degrees.  Coherent bundles (a camera, a parallel beam) keep the single-launch marcher; mostly incoherent batches over a
// volume that does not fit L2 go to the wavefront marcher (same results, DRAM traffic replaced by L2 hits).
static bool batch_is_incoherent(uint64_t n, int dim, const uint32_t *pos, const void *dir, int dir_dtype)
{
    if (n < 2) return false;
    const uint64_t samples = std::min<uint64_t>(n - 1, 4096);
    uint64_t bad = 0;
    for (uint64_t j = 0; j < samples; ++j)
    {
        const uint64_t i = ((j * 0x9E3779B97F4A7C15ull) >> 11) % (n - 1);      // scattered, not strided: a stride can alias with the row length of a ray grid
        bool far = false;
        double dot = 0, na = 0, nb = 0;
        for (int d = 0; d < dim; ++d)
        {
            const int32_t dp = (int32_t)(pos[(i + 1) * dim + d] - pos[i * dim + d]);
            if (dp > (4 << 16) || dp < -(4 << 16)) far = true;
            const double a = dir_dtype == VRT_I16 ? (double)((const int16_t *)dir)[i * dim + d] : (double)((const float *)dir)[i * dim + d];
            const double b = dir_dtype == VRT_I16 ? (double)((const int16_t *)dir)[(i + 1) * dim + d] : (double)((const float *)dir)[(i + 1) * dim + d];
            dot += a * b; na += a * a; nb += b * b;
        }
        if (far || !(dot * dot >= 0.81 * na * nb && dot >= 0)) ++bad;
    }
    return bad * 2 > samples;
}

// Per host thread and device: the two pipeline streams of vrt_trace and a pinned/device staging pair for small
// batches.  Created on first use, reused by every later call of that thread (stream creation and pinned allocation
// cost more than a small trace), released when the thread exits.
// Read-back helper of vrt_trace: a thread that issues the D2H copies of finished chunks while the calling thread is still staging
// later chunks.  With PAGEABLE host buffers both directions block the thread that issues them for the whole copy (the runtime stages
// through its own pinned buffers), so one thread pays H2D + D2H -- for config 5 that is 0.94 GB per pass, about as long as the march
// itself, and at 8 ranks per host it WAS the critical path of the end-to-end call.  Two threads pay max(H2D, D2H).
class Drainer
{
public:
    Drainer() : _stop(false), _busy(0) { _th = std::thread([this] { loop(); }); }
    ~Drainer()
    {
        { std::lock_guard<std::mutex> lk(_mu); _stop = true; }
        _cv.notify_all();
        if (_th.joinable()) _th.join();
    }
    void push(std::function<void()> job)
    {
        { std::lock_guard<std::mutex> lk(_mu); _q.push_back(std::move(job)); ++_busy; }
        _cv.notify_one();
    }
    // blocks until at most `limit` jobs are queued or running
    void wait_below(size_t limit)
    {
        std::unique_lock<std::mutex> lk(_mu);
        _done.wait(lk, [&] { return _busy <= limit; });
    }
private:
    void loop()
    {
        for (;;)
        {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(_mu);
                _cv.wait(lk, [&] { return _stop || !_q.empty(); });
                if (_q.empty()) return;
                job = std::move(_q.front());
                _q.pop_front();
            }
            job();
            { std::lock_guard<std::mutex> lk(_mu); --_busy; }
            _done.notify_all();
        }
    }
    std::thread _th;
    std::mutex _mu;
    std::condition_variable _cv, _done;
    std::deque<std::function<void()>> _q;
    bool _stop;
    size_t _busy;
};

struct ThreadCtx
{
    static constexpr size_t kSmallBytes = 1u << 20;
    std::unique_ptr<Drainer> drainer;      // created on the first large pipelined call of this thread
    int device = -1;
    static constexpr size_t kEvents = 128;
    cudaStream_t st[2] = {nullptr, nullptr}, in = nullptr, out = nullptr;
    cudaEvent_t ev[kEvents] = {};
    char *h_stage = nullptr;   // pinned
    char *d_stage = nullptr;
    bool ensure(int dev)
    {
        if (device == dev) return true;
        release();
        if (cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
        if (cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); release(); return false; }
        if (cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); release(); return false; }
        if (cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); release(); return false; }
        for (size_t i = 0; i < kEvents; ++i)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); release(); return false; }
        if (cudaMallocHost((void **)&h_stage, kSmallBytes) != cudaSuccess || cudaMalloc((void **)&d_stage, kSmallBytes) != cudaSuccess)
        { cudaGetLastError(); release(); return false; }
        device = dev;
        return true;
    }
    void release()
    {
        for (int i = 0; i < 2; ++i) if (st[i]) { cudaStreamDestroy(st[i]); st[i] = nullptr; }
        if (in) { cudaStreamDestroy(in); in = nullptr; }
        if (out) { cudaStreamDestroy(out); out = nullptr; }
        for (size_t i = 0; i < kEvents; ++i) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
        if (h_stage) { cudaFreeHost(h_stage); h_stage = nullptr; }
        if (d_stage) { cudaFree(d_stage); d_stage = nullptr; }
        device = -1;
        cudaGetLastError();
    }
    ~ThreadCtx() { release(); }   // at process exit the runtime may already be unloading: the calls then fail harmlessly
};
static thread_local ThreadCtx t_ctx;
static thread_local int t_cap_hit = -1;     // vrt_trace_cap_hit(): 1 / 0 / -1 (unknown)

// Is this batch a candidate for the wavefront marcher (the probe -- on the host for vrt_trace, on the device for vrt_trace_device --
// then decides)?  Large 3-D batches over a volume that does not fit L2, linear layout, no path output, device rounding.  Returns the
// brick log2 to use (5: 32^3-voxel bricks, the fastest on config 4; larger when the volume has too many of them), 0 = not a candidate.
// VRT_OPT_WAVE_LOG2 (and its round-1 alias VRT_OPT_REGION_LOG2: both name the brick edge): > 0 forced, 0 automatic, < 0 never
static int wave_request(const vrt_scene *s)
{
    const int w = (int)s->opt_wave.load(), r = (int)s->opt_region.load();
    return w != 0 ? w : r;
}

static int auto_wave_log2(const vrt_scene *s, uint64_t n, unsigned flags)
{
    if (s->dim != 3 || (flags & (VRT_TRACE_PATHS | VRT_TRACE_ROUND_HOST)) || s->bricked || s->tex || s->paired) return 0;
    if (n < (1u << 18) || n >= (1ull << 31)) return 0;
    if (s->nvox * 4 * elem_size(s->store) <= (96ull << 20)) return 0;
    uint32_t nb[3]; uint64_t K;
    for (int k = 5; k <= 8; ++k) if (wave_geometry(s, k, nb, &K)) return k;
    return 0;
}

extern "C" {

int vrt_trace_device(vrt_scene *s, uint64_t n, const uint32_t *d_pos, const void *d_dir, int dir_dtype, const float *invscale,
                     uint32_t minb, uint32_t iterations, unsigned flags, uint32_t *d_epos, void *d_edir, uint32_t *d_eit,
                     uint32_t *d_light, uint32_t *d_path, void *cuda_stream)
{
    int rc = validate_trace(s, n, d_pos, d_dir, dir_dtype, invscale, iterations, flags, d_epos, d_edir, d_eit, d_light, d_path);
    if (rc) return rc;
    DeviceGuard g(s->device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (n)
    {
        // buffers on another GPU would be reached over NVLink once peer access is on (vrt_scene_replicate enables it) -- slowly, and
        // with a stream that belongs to the wrong device: refuse instead
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, d_pos) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != s->device)
            return fail(VRT_ERR_INVALID, "ray buffers are on device " + std::to_string(at.device) + ", the scene is on device " + std::to_string(s->device));
        cudaGetLastError();
    }
    unsigned long long *scratch = nullptr;       // [0] refill counter, [1] cap flag, [2] probe flag
    VRT_CUDA(pool_alloc((void **)&scratch, 4 * sizeof(unsigned long long), s->device, st));
    MarchMode mode;
    const int wave = wave_request(s);
    if (wave > 0) mode.wave_log2 = wave;                                     // wavefront marcher on request
    const int auto_k = wave == 0 ? auto_wave_log2(s, n, flags) : 0;
    if (auto_k > 0)
    {
        // The batch lives in device memory and this call must not synchronise: a one-CTA probe kernel looks at the rays and writes a
        // flag; the single-launch marcher and the wavefront marcher are BOTH enqueued, gated on that flag, and one of them returns
        // at its first instruction.
        uint32_t *flag = (uint32_t *)(scratch + 2);
        coherence_probe_kernel<<<1, 256, 0, st>>>(d_pos, d_dir, dir_dtype == VRT_I16 ? 1 : 0, n, s->dim, flag);
        ++g_launches;
        VRT_CUDA(cudaGetLastError());
        mode.gate = flag; mode.gate_want = 0;
        rc = enqueue_march(s, n, d_pos, d_dir, dir_dtype, invscale, minb, iterations, flags, d_epos, d_edir, d_eit, d_light, d_path, scratch, st, mode);
        if (rc == VRT_OK)
        {
            mode.wave_log2 = auto_k; mode.gate_want = 1;
            rc = enqueue_march(s, n, d_pos, d_dir, dir_dtype, invscale, minb, iterations, flags, d_epos, d_edir, d_eit, d_light, d_path, scratch, st, mode);
        }
    }
    else
        rc = enqueue_march(s, n, d_pos, d_dir, dir_dtype, invscale, minb, iterations, flags, d_epos, d_edir, d_eit, d_light, d_path, scratch, st, mode);
    cudaFreeAsync(scratch, st);
    return rc;
}

int vrt_trace(vrt_scene *s, uint64_t n, const uint32_t *pos, const void *dir, int dir_dtype, const float *invscale,
              uint32_t minb, uint32_t iterations, unsigned flags, uint32_t *epos, void *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    int rc = validate_trace(s, n, pos, dir, dir_dtype, invscale, iterations, flags, epos, edir, eit, light, path);
    if (rc) return rc;
    if (n == 0) return VRT_OK;
    DeviceGuard g(s->device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    const int dim = s->dim;
    const size_t ds = elem_size(dir_dtype);
    const bool want_path = flags & VRT_TRACE_PATHS;
    if (!t_ctx.ensure(s->device)) return fail(VRT_ERR_CUDA, "could not create the per-thread streams / staging buffers");
    // Incoherent batches: the wavefront marcher on request (VRT_OPT_WAVE_LOG2 3..8) or -- default -- when the host-side coherence probe
    // says that most neighbouring rays of the batch are not neighbours in space.
    MarchMode mode;
    {
        const int wave = wave_request(s);
        if (wave > 0) mode.wave_log2 = wave;
        else if (wave == 0)
        {
            const int k = auto_wave_log2(s, n, flags);
            if (k > 0 && batch_is_incoherent(n, dim, pos, dir, dir_dtype)) mode.wave_log2 = k;
        }
    }
    const int region = mode.wave_log2;                                  // wavefront mode: fewer, larger chunks (below)
    t_cap_hit = -1;

    // Small batches (latency path): one packed H2D, one launch, one packed D2H through pinned staging -- 2 copies instead
    // of 6, no allocation.  Static ray-to-thread mapping (the grid covers the batch at once, nothing to refill).
    {
        const size_t b_pos = (size_t)n * dim * 4, b_dir = ((size_t)n * dim * ds + 3) & ~(size_t)3, b_u32 = (size_t)n * 4;
        if (!want_path && s->opt_chunk.load() == 0 && b_pos + b_dir + 2 * b_u32 <= ThreadCtx::kSmallBytes)
        {
            cudaStream_t q = t_ctx.st[0];
            char *h = t_ctx.h_stage, *d = t_ctx.d_stage;
            memcpy(h, pos, b_pos);
            memcpy(h + b_pos, dir, (size_t)n * dim * ds);
            VRT_CUDA(cudaMemcpyAsync(d, h, b_pos + b_dir, cudaMemcpyHostToDevice, q));
            uint32_t *d_pos = (uint32_t *)d; void *d_dir = d + b_pos;
            uint32_t *d_eit = (uint32_t *)(d + b_pos + b_dir), *d_light = (uint32_t *)(d + b_pos + b_dir + b_u32);
            rc = enqueue_march(s, n, d_pos, d_dir, dir_dtype, invscale, minb, iterations, flags, d_pos, d_dir, d_eit, d_light, nullptr, nullptr, q, mode);
            if (rc) return rc;
            VRT_CUDA(cudaMemcpyAsync(h, d, b_pos + b_dir + 2 * b_u32, cudaMemcpyDeviceToHost, q));
            VRT_CUDA(cudaStreamSynchronize(q));
            memcpy(epos, h, b_pos);
            memcpy(edir, h + b_pos, (size_t)n * dim * ds);
            memcpy(eit, h + b_pos + b_dir, b_u32);
            memcpy(light, h + b_pos + b_dir + b_u32, b_u32);
            int cap = 0;
            for (uint64_t i = 0; i < n; ++i) cap |= eit[i] == iterations;
            t_cap_hit = cap;
            return VRT_OK;
        }
    }

    // rays per pipelined chunk: copies of chunk i+1 / i-1 overlap the march of chunk i on the other stream
    uint64_t chunk = (uint64_t)s->opt_chunk.load();
    uint64_t chunk_wave = (uint64_t)s->num_sms * 1024;                      // rays resident at once (set below when the chunk size is chosen here)
    // at least 2^17 rays per chunk (one wave of the persistent grid), at most 16 chunks: config 2 (1 M rays) through pageable buffers
    // runs 8 chunks at 224 G ray-steps/s instead of 2 chunks at 190
    if (chunk == 0 && n > (1u << 17))
    {
        // ... and a whole number of WAVES of the persistent grid (148 SMs x 1024 resident threads = 151 552 rays): on a workload whose
        // rays all run equally long (config 5) a chunk of 131 072 rays leaves 13 % of the grid idle for a whole ray-time, which is
        // what a 2 M-ray shard of the strong-scaling run paid in every one of its 16 chunks (e2e 2174 -> see DESIGN.md section 7)
        uint64_t per_sm = 1024;                                             // 64 registers per thread
        if (s->all_clear && s->opt_allclear.load() != 0 && s->store == VRT_F32 && dim == 3 && !(flags & (VRT_TRACE_LIVE_TRANSLUCENCY | VRT_TRACE_PATHS | VRT_TRACE_ROUND_HOST)) &&
            invscale[0] == 1.0f && invscale[1] == 1.0f && invscale[2] == 1.0f && s->opt_block.load() <= 128 && (s->opt_kernel.load() == 0 || s->opt_kernel.load() == 3) &&
            !s->bricked && !s->tex && !s->paired)
            per_sm = 1152;                                                  // the all-clear kernel: 9 CTAs of 128 threads (MarchBounds)
        const int64_t cap = s->opt_max_ctas.load(), blk = s->opt_block.load();
        if (cap > 0) per_sm = std::min<uint64_t>(per_sm, (uint64_t)(cap * blk));
        const uint64_t wave = std::max<uint64_t>(1, (uint64_t)s->num_sms * per_sm);
        chunk_wave = wave;
        const uint64_t m = std::max<uint64_t>(1, ((n + 15) / 16 + wave / 2) / wave);
        chunk = std::max<uint64_t>(1u << 17, m * wave);
    }
    if (chunk == 0) chunk = n;
    if (region > 0 && s->opt_chunk.load() == 0) chunk = std::max<uint64_t>(chunk, std::min<uint64_t>(n, 4u << 20));   // the wavefront marcher wants many rays per brick
    if (want_path)
    {
        const uint64_t per_ray = (uint64_t)iterations * dim * 4;
        const uint64_t cap = std::max<uint64_t>(1, (1ull << 31) / std::max<uint64_t>(per_ray, 1));     // <= 2 GiB of polyline per chunk
        chunk = std::min(chunk, cap);
    }
    // Four streams: `in` carries every H2D copy, st[0]/st[1] the marches of even/odd chunks, `out` every D2H copy, tied
    // together by events.  All chunks are issued first (bounded by kWindowBytes / kWindowChunks of device memory), the D2H
    // copies afterwards, oldest first.  With pinned host buffers everything is asynchronous and the copies of chunk i+1 /
    // i-1 hide under the march of chunk i.  With PAGEABLE buffers (the std::vectors behind the reference API) the runtime
    // makes a H2D copy wait for the stream it is issued on and blocks the host for the whole of a D2H copy: separate copy
    // streams keep those waits off the marching streams, so the staging of later chunks and the read-back of earlier ones
    // still overlap the march (config 5 through pageable numpy arrays: 171 -> see DESIGN.md section 5).
    constexpr size_t kWindowBytes = 4ull << 30;
    constexpr size_t kWindowChunks = ThreadCtx::kEvents / 2;
    struct Pending { char *buf; uint64_t off, m; size_t o_dir, o_eit, o_light, o_path, b_path, bytes, o_flag; uint32_t *flag_slot; cudaEvent_t marched; };
    std::deque<Pending> pending;
    size_t pending_bytes = 0;
    int result = VRT_OK;
    auto note = [&](cudaError_t e) {
        if (e != cudaSuccess && result == VRT_OK) { result = fail(e == cudaErrorMemoryAllocation ? VRT_ERR_NOMEM : VRT_ERR_CUDA, cudaGetErrorString(e)); cudaGetLastError(); }
    };
    // the D2H copies of one chunk; runs on the calling thread or on the read-back helper (errors go to `drain_error`, merged below)
    std::atomic<int> drain_error{(int)cudaSuccess};
    const cudaStream_t out_stream = t_ctx.out;
    const int dev_id = s->device;
    auto drain_chunk = [=, &drain_error](const Pending &c, bool on_helper) {
        if (on_helper) cudaSetDevice(dev_id);
        cudaStream_t q = out_stream;
        const size_t b_pos = (size_t)c.m * dim * 4, b_dir = (size_t)c.m * dim * ds, b_u32 = (size_t)c.m * 4;
        auto chk = [&](cudaError_t e) { if (e != cudaSuccess) { int ok = (int)cudaSuccess; drain_error.compare_exchange_strong(ok, (int)e); cudaGetLastError(); } };
        chk(cudaStreamWaitEvent(q, c.marched, 0));
        if (drain_error.load() == (int)cudaSuccess)
        {
            chk(cudaMemcpyAsync(epos + c.off * dim, c.buf, b_pos, cudaMemcpyDeviceToHost, q));
            chk(cudaMemcpyAsync((char *)edir + c.off * dim * ds, c.buf + c.o_dir, b_dir, cudaMemcpyDeviceToHost, q));
            chk(cudaMemcpyAsync(eit + c.off, c.buf + c.o_eit, b_u32, cudaMemcpyDeviceToHost, q));
            chk(cudaMemcpyAsync(light + c.off, c.buf + c.o_light, b_u32, cudaMemcpyDeviceToHost, q));
            if (want_path) chk(cudaMemcpyAsync(path + c.off * iterations * dim, c.buf + c.o_path, c.b_path, cudaMemcpyDeviceToHost, q));
            if (c.flag_slot) chk(cudaMemcpyAsync(c.flag_slot, c.buf + c.o_flag, 4, cudaMemcpyDeviceToHost, q));   // pinned: asynchronous
        }
        cudaFreeAsync(c.buf, q);
    };
    // large batches hand the read-back to the helper thread as soon as a chunk's march is enqueued; small ones keep the old order
    // (issue everything, then read back oldest first) on the calling thread
    const bool use_helper = n >= 4 * chunk && !want_path && std::getenv("VRT_NO_READBACK_THREAD") == nullptr;
    if (use_helper && !t_ctx.drainer) t_ctx.drainer.reset(new Drainer());
    Drainer *const helper = use_helper ? t_ctx.drainer.get() : nullptr;
    auto drain_front = [&]() {
        Pending c = pending.front();
        pending.pop_front();
        pending_bytes -= c.bytes;
        drain_chunk(c, false);
    };
    uint32_t *const flag_slots = (uint32_t *)t_ctx.h_stage;
    constexpr uint64_t kFlagSlots = ThreadCtx::kSmallBytes / 4;

    // Chunk schedule.  The first chunk's H2D copy and the last chunk's D2H copy cannot overlap any marching, and a chunk can only be
    // staged while its predecessors march: when the regular chunk is several waves of the persistent grid, the batch starts with 1, 2,
    // 4, ... waves (staging pageable memory is ~2.5x as fast as marching the same rays: each chunk is staged before its predecessor
    // has finished) and ends with ..., 4, 2, 1 waves (the read-back, ~1.9x as fast as the march, follows right behind it), so that head
    // and tail of the copy pipeline are one wave each instead of one regular chunk (config 5 on one GPU: 6 waves).
    std::vector<uint64_t> sizes;
    {
        const uint64_t wave = std::max<uint64_t>(1u << 17, chunk_wave);
        std::vector<uint64_t> ramp;
        if (s->opt_chunk.load() == 0 && !want_path && region == 0 && s->dim == 3 && chunk >= 2 * wave)
            for (uint64_t w = wave; w < chunk; w *= 2) ramp.push_back(w);
        uint64_t ramp_sum = 0;
        for (uint64_t w : ramp) ramp_sum += w;
        if (ramp.empty() || n < 2 * ramp_sum + 2 * chunk) ramp.clear(), ramp_sum = 0;
        for (uint64_t w : ramp) sizes.push_back(w);
        uint64_t middle = n - 2 * ramp_sum;
        // without a ramp (chunks of one wave: a shard of a multi-GPU run) the partial chunk goes FIRST: its launch leaves most of the
        // machine to the next chunk, which starts at once on the other stream, whereas a partial last chunk runs alone for a whole ray-time
        if (ramp.empty() && chunk < n && n % chunk != 0) { sizes.push_back(n % chunk); middle -= n % chunk; }
        while (middle > 0)
        {
            const uint64_t m = (middle >= chunk + chunk / 2 || ramp.empty()) ? std::min(chunk, middle) : middle;   // the remainder joins the last regular chunk
            sizes.push_back(m);
            middle -= m;
        }
        for (size_t i = ramp.size(); i-- > 0;) sizes.push_back(ramp[i]);
    }
    uint64_t index = 0;
    for (uint64_t off = 0, m = 0; index < sizes.size() && result == VRT_OK; off += m, ++index)
    {
        m = sizes[index];
        cudaStream_t q = t_ctx.st[index & 1];
        // one stream-ordered allocation per chunk: [pos | dir | eit | light | counter | path]
        Pending c;
        const size_t b_pos = (size_t)m * dim * 4, b_dir = (size_t)m * dim * ds, b_u32 = (size_t)m * 4;
        c.off = off; c.m = m;
        c.o_dir = (b_pos + 255) & ~(size_t)255; c.o_eit = (c.o_dir + b_dir + 255) & ~(size_t)255; c.o_light = (c.o_eit + b_u32 + 255) & ~(size_t)255;
        const size_t o_cnt = (c.o_light + b_u32 + 255) & ~(size_t)255;
        c.o_path = o_cnt + 256;
        c.o_flag = o_cnt + 8;
        c.flag_slot = index < kFlagSlots ? flag_slots + index : nullptr;
        if (c.flag_slot) *c.flag_slot = 0;
        c.b_path = want_path ? (size_t)m * iterations * dim * 4 : 0;
        c.bytes = c.o_path + c.b_path;
        if (helper) helper->wait_below(std::min<size_t>(kWindowChunks - 1, std::max<size_t>(1, kWindowBytes / std::max<size_t>(c.bytes, 1)) - 1));
        while (!pending.empty() && (pending_bytes + c.bytes > kWindowBytes || pending.size() >= kWindowChunks)) drain_front();
        if (result != VRT_OK) break;
        c.buf = nullptr;
        note(pool_alloc((void **)&c.buf, c.bytes, s->device, t_ctx.in));
        if (result != VRT_OK) break;
        cudaEvent_t staged = t_ctx.ev[(index * 2) % ThreadCtx::kEvents];
        c.marched = t_ctx.ev[(index * 2 + 1) % ThreadCtx::kEvents];
        note(cudaMemcpyAsync(c.buf, pos + off * dim, b_pos, cudaMemcpyHostToDevice, t_ctx.in));
        note(cudaMemcpyAsync(c.buf + c.o_dir, (const char *)dir + off * dim * ds, b_dir, cudaMemcpyHostToDevice, t_ctx.in));
        note(cudaEventRecord(staged, t_ctx.in));
        note(cudaStreamWaitEvent(q, staged, 0));
        if (result == VRT_OK)
        {
            // results overwrite the start buffers on the device ("written back in place")
            uint32_t *d_pos = (uint32_t *)c.buf; void *d_dir = c.buf + c.o_dir;
            int rc2 = enqueue_march(s, m, d_pos, d_dir, dir_dtype, invscale, minb, iterations, flags, d_pos, d_dir, (uint32_t *)(c.buf + c.o_eit),
                                    (uint32_t *)(c.buf + c.o_light), want_path ? (uint32_t *)(c.buf + c.o_path) : nullptr,
                                    (unsigned long long *)(c.buf + o_cnt), q, mode);
            if (rc2 && result == VRT_OK) result = rc2;
        }
        note(cudaEventRecord(c.marched, q));
        if (helper) helper->push([drain_chunk, c] { drain_chunk(c, true); });
        else { pending.push_back(c); pending_bytes += c.bytes; }
    }
    while (!pending.empty()) drain_front();
    if (helper) helper->wait_below(0);
    if (drain_error.load() != (int)cudaSuccess) note((cudaError_t)drain_error.load());
    cudaStream_t all[4] = {t_ctx.in, t_ctx.st[0], t_ctx.st[1], t_ctx.out};
    for (cudaStream_t q : all) note(cudaStreamSynchronize(q));
    if (result == VRT_OK && index <= kFlagSlots)
    {
        uint32_t any = 0;
        for (uint64_t i = 0; i < index; ++i) any |= flag_slots[i];
        t_cap_hit = any ? 1 : 0;
    }
    return result;
}

int vrt_trace_cap_hit(void) { return t_cap_hit; }

int vrt_normalise_rays_device(vrt_scene *s, uint64_t n, uint32_t *d_pos, void *d_dir, int dir_dtype, int64_t *first_bad_ray, void *cuda_stream)
{
    if (!s) return fail(VRT_ERR_INVALID, "scene is null");
    if (!s->d_ior) return fail(VRT_ERR_UNSUPPORTED, "scene was not created from an ior volume");
    if ((s->ior_dtype == VRT_F32) != (dir_dtype == VRT_F32))
        return fail(VRT_ERR_UNSUPPORTED, "the reference instantiates float scene/float dirs and int16 scene/int16 dirs only (image_util.cpp:914-955)");
    if (first_bad_ray) *first_bad_ray = 0;
    if (n == 0) return VRT_OK;
    if (!d_pos || !d_dir) return fail(VRT_ERR_INVALID, "null ray buffer");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    unsigned long long *d_flags = nullptr;
    VRT_CUDA(pool_alloc((void **)&d_flags, 16, s->device, st));
    VRT_CUDA(cudaMemsetAsync(d_flags, 0xFF, 16, st));
    NormParams np;
    np.dim = s->dim; np.n = n;
    for (int d = 0; d < 3; ++d) np.ib[d] = d < s->dim ? (uint32_t)s->ior_bounds[d] : 1;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (s->ior_dtype == VRT_F32) normalise_kernel<true><<<grid, 256, 0, st>>>(np, s->d_ior, d_pos, d_dir, d_flags, d_flags + 1);
    else                         normalise_kernel<false><<<grid, 256, 0, st>>>(np, s->d_ior, d_pos, d_dir, d_flags, d_flags + 1);
    ++g_launches;
    VRT_CUDA(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    VRT_CUDA(cudaMemcpyAsync(h, d_flags, 16, cudaMemcpyDeviceToHost, st));
    VRT_CUDA(cudaStreamSynchronize(st));
    cudaFreeAsync(d_flags, st);
    if (h[0] != ~0ull)
    {
        if (first_bad_ray) *first_bad_ray = (int64_t)h[0];
        return fail(VRT_ERR_INVALID, "ray " + std::to_string(h[0] - 1) + " is not in 0 to bounds");          // image_util.cpp:686-691
    }
    if (h[1] != ~0ull) return fail(VRT_ERR_INVALID, "Normalize length failed (ray " + std::to_string(h[1] - 1) + ")"); // image_util.cpp:703
    return VRT_OK;
}

int vrt_selftest_division(int device, uint64_t *mismatches)
{
    if (!mismatches) return fail(VRT_ERR_INVALID, "mismatches is null");
    DeviceGuard g(device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    unsigned long long *d_bad = nullptr, h_bad = 0;
    VRT_CUDA(cudaMalloc((void **)&d_bad, 8));
    VRT_CUDA(cudaMemset(d_bad, 0, 8));
    // every bit pattern div_is_fast() accepts: [0x10000000, 0x70000000), plus a margin on both sides that it must reject
    div_selftest_kernel<0><<<148 * 8, 256>>>(0x0F000000u, 0x62000000u, d_bad);
    // rni_small(): every float with |v| < 2^22 (bit patterns below 0x4A800000, both signs)
    rni_selftest_kernel<<<148 * 8, 256>>>(0x00000000u, 0x4A800000u, d_bad);
    rni_selftest_kernel<<<148 * 8, 256>>>(0x80000000u, 0x4A800000u, d_bad);
    g_launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    VRT_CUDA(e);
    *mismatches = h_bad;
    return VRT_OK;
}

int vrt_measure_gather_bandwidth(int device, uint64_t bytes, int sector_bytes, int iters, double *gb_per_s)
{
    if (!gb_per_s || (sector_bytes != 16 && sector_bytes != 32) || bytes < 4096 || iters < 1) return fail(VRT_ERR_INVALID, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(VRT_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    VRT_CUDA(cudaGetDeviceProperties(&prop, device));
    uint4 *buf = nullptr; unsigned long long *sink = nullptr;
    VRT_CUDA(cudaMalloc((void **)&buf, bytes));
    VRT_CUDA(cudaMalloc((void **)&sink, 8));
    VRT_CUDA(cudaMemset(buf, 1, bytes));
    const unsigned long long nsectors = bytes / sector_bytes;
    const int block = 256, grid = prop.multiProcessorCount * 8, rounds = 64;
    cudaEvent_t a, b;
    VRT_CUDA(cudaEventCreate(&a)); VRT_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it)
    {
        VRT_CUDA(cudaEventRecord(a));
        if (sector_bytes == 32) gather_kernel<32><<<grid, block>>>(buf, nsectors, rounds, sink);
        else                    gather_kernel<16><<<grid, block>>>(buf, nsectors, rounds, sink);
        ++g_launches;
        VRT_CUDA(cudaEventRecord(b));
        VRT_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (it >= 2) best = std::min(best, ms);
    }
    const double moved = (double)grid * block * rounds * 8.0 * sector_bytes;
    *gb_per_s = moved / (best * 1e-3) / 1e9;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    return VRT_OK;
}

} // extern "C"
