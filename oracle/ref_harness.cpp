// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product.
//
// C-ABI wrapper around the UNMODIFIED reference implementation, compiled from the sources
// where they lie under $(REF_SRC) (= /root/reference/src) by oracle/Makefile into
// oracle/_ref/libvrt_ref.so.  No reference source is copied: this translation unit
// textually #includes the reference's cuda_volume_raytracer.cu (as plain C++ with -DNCUDA,
// exactly as the reference Makefile:57-58 builds its CPU object) so that, in one TU, we get
//   * TraceRaysCu<float|diff_t>            (the shipped boundary class, cu:637-989)
//   * trace_rays_cpu<...>                  (cu:376-394) instantiated with a LIVE
//                                          translucency_t* / brightness_t, which the shipped
//                                          call sites compile out (cu:853-938)
// and links image_util.o/util.o/io_util.o/serialize.o for RaytraceScene<> (scene prep and the
// ray pre/post-processing above the boundary, image_util.cpp:501-772).
//
// It is used (a) to pin the C restatement in oracle/vrt_oracle.c, (b) to generate the golden
// fixtures under tests/golden/, and (c) as the "reference" CPU baseline of bench.py.
// Built with -fno-access-control so that private members (gradient volume etc.) can be read.

// TU 1 of 2: includes the reference .cu (TraceRaysCu<> + trace_rays_cpu<>).  The scene-level wrappers
// live in ref_harness_scene.cpp because tuple_io.h and io_util.h both declare `print`.

#ifdef VRTREF_HEADER_ONLY   // libvrt_dropin.so: only the boundary class, declared by the reference's header and DEFINED by our shim
#include <vector>
#include <memory>
#include <string>
#include <stdexcept>
#include <cstdint>
#include <cstddef>
#include <iostream>
#include <omp.h>
#include "cuda_volume_raytracer.h"
#else
#ifndef VRTREF_CUDA      // VRTREF_CUDA: the same harness compiled by nvcc for sm_100 = the reference's own CUDA trace
#define NCUDA 1
#endif
#include "cuda_volume_raytracer.cu"   // found via -I$(REF_SRC)
#endif

#include <cstring>
#include <string>
#include <memory>

extern "C" void vrtref_set_error(const char *msg);
#define VRTREF_TRY try {
#define VRTREF_CATCH                                                                 \
    } catch (std::exception const & e) { vrtref_set_error(e.what()); return -1; }   \
      catch (...) { vrtref_set_error("unknown exception"); return -1; }             \
    return 0;

static thread_local std::string g_err;
extern "C" void vrtref_set_error(const char *msg) { g_err = msg; }

namespace {
template <typename DiffType>
struct TracerBox
{
    std::vector<translucency_t> translucency;   // TraceRaysCu keeps a const& (h:65)
    std::vector<std::vector<DiffType> > diff;
    std::unique_ptr<TraceRaysCu<DiffType> > tracer;
};

template <typename DiffType>
int tracer_new(void **out, const size_t *bounds, int dim, const DiffType *const *diff, const uint32_t *tr)
{
    VRTREF_TRY
    auto *box = new TracerBox<DiffType>();
    std::vector<size_t> b(bounds, bounds + dim);
    size_t n = 1; for (size_t v : b) n *= v;
    box->translucency.assign(tr, tr + n);
    box->diff.resize(dim);
    for (int d = 0; d < dim; ++d) box->diff[d].assign(diff[d], diff[d] + n);
    box->tracer.reset(new TraceRaysCu<DiffType>(b, box->diff, box->translucency));
    *out = box;
    VRTREF_CATCH
}

template <typename DiffType, typename DirType>
int tracer_trace(void *h, size_t n, const uint32_t *pos, const DirType *dir, const float *invscale,
                 uint32_t minb, uint32_t iterations, int trace_path, int max_cpu,
                 uint32_t *epos, DirType *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    VRTREF_TRY
    auto *box = static_cast<TracerBox<DiffType>*>(h);
    size_t dim = box->tracer->_output_sizes.size();
    std::vector<pos_t> sp(pos, pos + n * dim);
    std::vector<DirType> sd(dir, dir + n * dim);
    std::vector<float> isc(invscale, invscale + dim);
    // pre-sized by the caller, as image_util.cpp:738-741 does
    std::vector<pos_t> ep(n * dim); std::vector<DirType> ed(n * dim); std::vector<uint32_t> ei(n); std::vector<brightness_t> rl(n); std::vector<pos_t> pa;
    Options opt; opt._loglevel = 0; if (max_cpu > 0) opt._max_cpu = max_cpu;
    box->tracer->trace_rays_cu(sp, sd, ep, ed, ei, rl, pa, isc, minb, iterations, trace_path != 0, opt);
    std::memcpy(epos, ep.data(), ep.size() * sizeof(pos_t));
    std::memcpy(edir, ed.data(), ed.size() * sizeof(DirType));
    std::memcpy(eit, ei.data(), ei.size() * sizeof(uint32_t));
    std::memcpy(light, rl.data(), rl.size() * sizeof(uint32_t));
    if (trace_path && path) std::memcpy(path, pa.data(), pa.size() * sizeof(pos_t));
    VRTREF_CATCH
}

#ifndef VRTREF_HEADER_ONLY
// The LIVE-translucency instantiation of the reference marcher (cu:337-341,370-373): the template is
// given a real translucency_t* and brightness_t instead of DummyArray/DummyObject.  Mirrors what
// trace_rays_cu_impl does around the call (fill_struct cu:468-488, read_struct cu:490-516,
// end_iteration = iterations - it cu:953-956) using the reference's own helpers.
template <typename DiffType, typename DirType, uint8_t dim>
int trace_live_dim(DiffType *vol, const uint32_t *tr, const size_t *bounds, const float *invscale,
               size_t n, const uint32_t *pos, const DirType *dir, uint32_t iterations, uint32_t minb,
               int trace_path, int threads,
               uint32_t *epos, DirType *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    VRTREF_TRY
    std::vector<pos_t> sp(pos, pos + n * dim);
    std::vector<DirType> sd(dir, dir + n * dim);
    std::vector<raydata_t<dim, DirType> > ray_data;
    fill_struct<dim>(sp, sd, iterations, ray_data);
    cuda_tuple<float, dim> isc = make_struct<float, dim>()(invscale);
    std::vector<uint16_t> b16(bounds, bounds + dim);
    cuda_tuple<uint16_t, dim> osz = make_struct<uint16_t, dim>()(b16.data());
    translucency_t *trp = const_cast<translucency_t*>(tr);
    brightness_t mb = minb;
    size_t chunk = 0x8000;
    for (size_t i = 0; i < n; i += chunk)
    {
        size_t m = std::min(chunk, n - i);
        if (!trp && trace_path)
            trace_rays_cpu(vol, DummyArray(), osz, isc, ray_data.data() + i, reinterpret_cast<cuda_tuple<pos_t,dim>*>(path) + i * iterations, iterations, DummyObject(), m, (size_t)threads);
        else if (!trp)
            trace_rays_cpu(vol, DummyArray(), osz, isc, ray_data.data() + i, DummyArray(), iterations, DummyObject(), m, (size_t)threads);
        else if (trace_path)
            trace_rays_cpu(vol, trp, osz, isc, ray_data.data() + i, reinterpret_cast<cuda_tuple<pos_t,dim>*>(path) + i * iterations, iterations, mb, m, (size_t)threads);
        else
            trace_rays_cpu(vol, trp, osz, isc, ray_data.data() + i, DummyArray(), iterations, mb, m, (size_t)threads);
    }
    std::vector<pos_t> ep(n * dim); std::vector<DirType> ed(n * dim); std::vector<uint32_t> ei(n); std::vector<brightness_t> rl(n);
    read_struct<dim>(ep, ed, rl, ei, ray_data);
    for (size_t i = 0; i < n; ++i) ei[i] = iterations - ei[i];
    std::memcpy(epos, ep.data(), ep.size() * sizeof(pos_t));
    std::memcpy(edir, ed.data(), ed.size() * sizeof(DirType));
    std::memcpy(eit, ei.data(), ei.size() * sizeof(uint32_t));
    std::memcpy(light, rl.data(), rl.size() * sizeof(uint32_t));
    VRTREF_CATCH
}

#endif
} // namespace

extern "C" {

const char *vrtref_last_error() { return g_err.c_str(); }
int vrtref_omp_max_threads() { return omp_get_max_threads(); }
int vrtref_is_cuda_build()
{
#ifdef VRTREF_CUDA
    return 1;
#else
    return 0;
#endif
}

// ---- TraceRaysCu<> (boundary level, h:61-115) ----
int vrtref_tracer_new_f32(void **out, const size_t *bounds, int dim, const float *const *diff, const uint32_t *tr) { return tracer_new<float>(out, bounds, dim, diff, tr); }
int vrtref_tracer_new_i16(void **out, const size_t *bounds, int dim, const int16_t *const *diff, const uint32_t *tr) { return tracer_new<diff_t>(out, bounds, dim, diff, tr); }
void vrtref_tracer_delete_f32(void *h) { delete static_cast<TracerBox<float>*>(h); }
void vrtref_tracer_delete_i16(void *h) { delete static_cast<TracerBox<diff_t>*>(h); }
#if !defined(VRTREF_CUDA) && !defined(VRTREF_HEADER_ONLY)   // private member access needs -fno-access-control (host build only)
void vrtref_tracer_interleaved_f32(void *h, float *out)   { auto *b = static_cast<TracerBox<float>*>(h);  size_t n = b->diff[0].size() * (b->diff.size() + 1); std::memcpy(out, b->tracer->_diff_interleaved.get(), n * sizeof(float)); }
void vrtref_tracer_interleaved_i16(void *h, int16_t *out) { auto *b = static_cast<TracerBox<diff_t>*>(h); size_t n = b->diff[0].size() * (b->diff.size() + 1); std::memcpy(out, b->tracer->_diff_interleaved.get(), n * sizeof(int16_t)); }
#endif

#define VRTREF_TRACER_TRACE(NAME, DIFF, DIR)                                                                                     \
int NAME(void *h, size_t n, const uint32_t *pos, const DIR *dir, const float *invscale, uint32_t minb, uint32_t iterations,      \
         int trace_path, int max_cpu, uint32_t *epos, DIR *edir, uint32_t *eit, uint32_t *light, uint32_t *path)                 \
{ return tracer_trace<DIFF, DIR>(h, n, pos, dir, invscale, minb, iterations, trace_path, max_cpu, epos, edir, eit, light, path); }
VRTREF_TRACER_TRACE(vrtref_tracer_trace_f32_f32, float, float)
VRTREF_TRACER_TRACE(vrtref_tracer_trace_f32_i16, float, int16_t)
VRTREF_TRACER_TRACE(vrtref_tracer_trace_i16_f32, diff_t, float)
VRTREF_TRACER_TRACE(vrtref_tracer_trace_i16_i16, diff_t, int16_t)

#ifndef VRTREF_HEADER_ONLY
// ---- live translucency / minimum brightness: trace_rays_cpu on an interleaved volume ----
#define VRTREF_TRACE_LIVE(NAME, DIFF, DIR)                                                                                        \
int NAME(const DIFF *vol, const uint32_t *tr, const size_t *bounds, int dim, const float *invscale, size_t n, const uint32_t *pos, \
         const DIR *dir, uint32_t iterations, uint32_t minb, int trace_path, int threads,                                          \
         uint32_t *epos, DIR *edir, uint32_t *eit, uint32_t *light, uint32_t *path)                                                \
{                                                                                                                                  \
    if (dim == 3) return trace_live_dim<DIFF, DIR, 3>(const_cast<DIFF*>(vol), tr, bounds, invscale, n, pos, dir, iterations, minb, trace_path, threads, epos, edir, eit, light, path); \
    if (dim == 2) return trace_live_dim<DIFF, DIR, 2>(const_cast<DIFF*>(vol), tr, bounds, invscale, n, pos, dir, iterations, minb, trace_path, threads, epos, edir, eit, light, path); \
    vrtref_set_error("Illegal dimension"); return -1;                                                                                        \
}
// tr == NULL: the SHIPPED instantiation (DummyArray translucency, DummyObject brightness: cu:916-941) run directly on a
// caller-provided interleaved volume, i.e. trace_rays_cu_impl without the ctor's extra host copy of the volume.
VRTREF_TRACE_LIVE(vrtref_trace_live_f32_f32, float, float)
VRTREF_TRACE_LIVE(vrtref_trace_live_f32_i16, float, int16_t)
VRTREF_TRACE_LIVE(vrtref_trace_live_i16_f32, diff_t, float)
VRTREF_TRACE_LIVE(vrtref_trace_live_i16_i16, diff_t, int16_t)

#endif

} // extern "C"
