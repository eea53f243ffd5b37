// TEST INFRASTRUCTURE ONLY -- see ref_harness.cpp.  TU 2 of 2: C-ABI wrappers around the reference's
// RaytraceScene<> (scene prep image_util.cpp:501-643, ray pre/post-processing + trace image_util.cpp:645-772)
// and host interpolator<T> (image_util.h:348-431).  Links against the unmodified image_util.o.
// Built with -fno-access-control so the gradient volume the constructor produced can be read out.

#include <memory>
#include <vector>
#include <stdexcept>
#include <cstdlib>
#include "cuda_volume_raytracer.h"   // found via -I$(REF_SRC)
#include "image_util.h"
#include <fstream>

#include <cstring>
#include <string>
#include <memory>

extern "C" void vrtref_set_error(const char *msg);
#define VRTREF_TRY try {
#define VRTREF_CATCH                                                                 \
    } catch (std::exception const & e) { vrtref_set_error(e.what()); return -1; }   \
      catch (...) { vrtref_set_error("unknown exception"); return -1; }             \
    return 0;

namespace {
template <typename IorType, typename IorLogType, typename DiffType>
struct SceneBox
{
    std::unique_ptr<RaytraceScene<IorType, IorLogType, DiffType> > scene;
};

template <typename IorType, typename IorLogType, typename DiffType>
int scene_new(void **out, const size_t *bounds, int dim, const IorType *ior, const uint32_t *tr, int loglevel)
{
    VRTREF_TRY
    std::vector<size_t> b(bounds, bounds + dim);
    size_t n = 1; for (size_t v : b) n *= v;
    std::vector<IorType> iorv(ior, ior + n);
    std::vector<translucency_t> trv(tr, tr + n);
    Options opt; opt._loglevel = loglevel;
    auto *box = new SceneBox<IorType, IorLogType, DiffType>();
    box->scene.reset(new RaytraceScene<IorType, IorLogType, DiffType>(b, iorv, trv, opt));
    *out = box;
    VRTREF_CATCH
}

template <typename IorType, typename IorLogType, typename DiffType, typename DirType>
int scene_trace(void *h, size_t n, const uint32_t *pos, const DirType *dir, const float *invscale,
                uint32_t minb, uint32_t iterations, int trace_path, int max_cpu,
                uint32_t *epos, DirType *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{
    VRTREF_TRY
    auto *box = static_cast<SceneBox<IorType, IorLogType, DiffType>*>(h);
    size_t dim = box->scene->_bound_vec.size();
    std::vector<pos_t> sp(pos, pos + n * dim);
    std::vector<DirType> sd(dir, dir + n * dim);
    std::vector<float> isc(invscale, invscale + dim);
    std::vector<pos_t> ep; std::vector<DirType> ed; std::vector<uint32_t> ei; std::vector<brightness_t> rl; std::vector<pos_t> pa;
    Options opt; opt._loglevel = 0; opt._minimum_gpu = 0x80; if (max_cpu > 0) opt._max_cpu = max_cpu;
    box->scene->trace_rays(RayTraceRayInstanceRef<DirType>(sp, sd, isc, minb, iterations, trace_path != 0, true),
                           ep, ed, ei, rl, pa, opt);
    std::memcpy(epos, ep.data(), ep.size() * sizeof(pos_t));
    std::memcpy(edir, ed.data(), ed.size() * sizeof(DirType));
    std::memcpy(eit, ei.data(), ei.size() * sizeof(uint32_t));
    std::memcpy(light, rl.data(), rl.size() * sizeof(uint32_t));
    if (trace_path && path) std::memcpy(path, pa.data(), pa.size() * sizeof(pos_t));
    VRTREF_CATCH
}

} // namespace

extern "C" {

// ---- RaytraceScene<> (API level: scene prep + normalise + trace + coordinate shifts) ----
int vrtref_scene_new_f32(void **out, const size_t *bounds, int dim, const float *ior, const uint32_t *tr, int loglevel)
{ return scene_new<float, float, float>(out, bounds, dim, ior, tr, loglevel); }
int vrtref_scene_new_u32(void **out, const size_t *bounds, int dim, const uint32_t *ior, const uint32_t *tr, int loglevel)
{ return scene_new<ior_t, iorlog_t, diff_t>(out, bounds, dim, ior, tr, loglevel); }
void vrtref_scene_delete_f32(void *h) { delete static_cast<SceneBox<float, float, float>*>(h); }
void vrtref_scene_delete_u32(void *h) { delete static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h); }

int vrtref_scene_trace_f32(void *h, size_t n, const uint32_t *pos, const float *dir, const float *invscale, uint32_t minb, uint32_t iterations, int trace_path, int max_cpu,
                           uint32_t *epos, float *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{ return scene_trace<float, float, float, float>(h, n, pos, dir, invscale, minb, iterations, trace_path, max_cpu, epos, edir, eit, light, path); }
int vrtref_scene_trace_u32(void *h, size_t n, const uint32_t *pos, const int16_t *dir, const float *invscale, uint32_t minb, uint32_t iterations, int trace_path, int max_cpu,
                           uint32_t *epos, int16_t *edir, uint32_t *eit, uint32_t *light, uint32_t *path)
{ return scene_trace<ior_t, iorlog_t, diff_t, dir_t>(h, n, pos, dir, invscale, minb, iterations, trace_path, max_cpu, epos, edir, eit, light, path); }

// read-outs of what the scene constructor produced (needs -fno-access-control)
void vrtref_scene_diff_bounds_f32(void *h, size_t *out) { auto &s = *static_cast<SceneBox<float, float, float>*>(h)->scene; std::copy(s._diff_bound_vec.begin(), s._diff_bound_vec.end(), out); }
void vrtref_scene_diff_bounds_u32(void *h, size_t *out) { auto &s = *static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h)->scene; std::copy(s._diff_bound_vec.begin(), s._diff_bound_vec.end(), out); }
// interleaved [d0,d1,(d2,)extra] gradient volume as built by the TraceRaysCu ctor (cu:654-669)
void vrtref_scene_interleaved_f32(void *h, float *out)
{ auto &s = *static_cast<SceneBox<float, float, float>*>(h)->scene; size_t n = s._diff[0].size() * (s._diff.size() + 1); std::memcpy(out, s._calculation_object->_diff_interleaved.get(), n * sizeof(float)); }
void vrtref_scene_interleaved_u32(void *h, int16_t *out)
{ auto &s = *static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h)->scene; size_t n = s._diff[0].size() * (s._diff.size() + 1); std::memcpy(out, s._calculation_object->_diff_interleaved.get(), n * sizeof(int16_t)); }
void vrtref_scene_diff_f32(void *h, int axis, float *out)   { auto &s = *static_cast<SceneBox<float, float, float>*>(h)->scene; std::memcpy(out, s._diff[axis].data(), s._diff[axis].size() * sizeof(float)); }
void vrtref_scene_diff_u32(void *h, int axis, int16_t *out) { auto &s = *static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h)->scene; std::memcpy(out, s._diff[axis].data(), s._diff[axis].size() * sizeof(int16_t)); }
void vrtref_scene_iorlog_f32(void *h, float *out)   { auto &s = *static_cast<SceneBox<float, float, float>*>(h)->scene; std::memcpy(out, s._ior_log.data(), s._ior_log.size() * sizeof(float)); }
void vrtref_scene_iorlog_u32(void *h, int32_t *out) { auto &s = *static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h)->scene; std::memcpy(out, s._ior_log.data(), s._ior_log.size() * sizeof(int32_t)); }
void vrtref_scene_translucency_cropped_f32(void *h, uint32_t *out) { auto &s = *static_cast<SceneBox<float, float, float>*>(h)->scene; std::memcpy(out, s._translucency_cropped.data(), s._translucency_cropped.size() * 4); }
void vrtref_scene_translucency_cropped_u32(void *h, uint32_t *out) { auto &s = *static_cast<SceneBox<ior_t, iorlog_t, diff_t>*>(h)->scene; std::memcpy(out, s._translucency_cropped.data(), s._translucency_cropped.size() * 4); }

// the reference's own serializer (serialize.h, image_util.cpp:35-144): what it writes, our reader must read and vice versa
int vrtref_write_scene_instance_u32(const char *path, const size_t *bounds, int dim, const uint32_t *ior, const uint32_t *tr)
{
    VRTREF_TRY
    RayTraceSceneInstance<ior_t> inst;
    inst._bound_vec.assign(bounds, bounds + dim);
    size_t n = 1; for (int d = 0; d < dim; ++d) n *= bounds[d];
    inst._ior.assign(ior, ior + n); inst._translucency.assign(tr, tr + n);
    std::ofstream out(path, std::ios::binary);
    SERIALIZE::write_value(out, inst);
    VRTREF_CATCH
}
// reads a combined instance with the reference's reader and traces it with the reference CPU path
int vrtref_replay_instance_u32(const char *path, size_t *n_rays, uint32_t *epos, int16_t *edir, uint32_t *eit, size_t cap)
{
    VRTREF_TRY
    RaytraceInstance<ior_t, dir_t> inst;
    std::ifstream in(path, std::ios::binary);
    SERIALIZE::read_value(in, inst);
    std::vector<pos_t> ep; std::vector<dir_t> ed; std::vector<uint32_t> ei; std::vector<brightness_t> rl; std::vector<pos_t> pa;
    Options opt;
    trace_rays<ior_t, iorlog_t, diff_t, dir_t>(inst, ep, ed, ei, rl, pa, opt);
    *n_rays = ei.size();
    if (ei.size() > cap) throw std::runtime_error("output buffers too small");
    std::memcpy(epos, ep.data(), ep.size() * 4); std::memcpy(edir, ed.data(), ed.size() * 2); std::memcpy(eit, ei.data(), ei.size() * 4);
    VRTREF_CATCH
}

// host interpolator<T> (image_util.h:348-431) -- pins axis order / fraction semantics (image_util_test.h:4-35)
int vrtref_interpolate_f32(const float *img, const size_t *bounds, int dim, const uint32_t *pos, size_t n, float *out)
{
    VRTREF_TRY
    std::vector<size_t> b(bounds, bounds + dim); size_t m = 1; for (size_t v : b) m *= v;
    std::vector<float> im(img, img + m);
    interpolator<float> interp(im, b);
    for (size_t i = 0; i < n; ++i) out[i] = interp(pos + i * dim);
    VRTREF_CATCH
}
int vrtref_interpolate_u32(const uint32_t *img, const size_t *bounds, int dim, const uint32_t *pos, size_t n, uint32_t *out)
{
    VRTREF_TRY
    std::vector<size_t> b(bounds, bounds + dim); size_t m = 1; for (size_t v : b) m *= v;
    std::vector<uint32_t> im(img, img + m);
    interpolator<uint32_t> interp(im, b);
    for (size_t i = 0; i < n; ++i) out[i] = interp(pos + i * dim);
    VRTREF_CATCH
}
int vrtref_interpolate_i32(const int32_t *img, const size_t *bounds, int dim, const uint32_t *pos, size_t n, int32_t *out)
{
    VRTREF_TRY
    std::vector<size_t> b(bounds, bounds + dim); size_t m = 1; for (size_t v : b) m *= v;
    std::vector<int32_t> im(img, img + m);
    interpolator<int32_t> interp(im, b);
    for (size_t i = 0; i < n; ++i) out[i] = interp(pos + i * dim);
    VRTREF_CATCH
}

} // extern "C"
