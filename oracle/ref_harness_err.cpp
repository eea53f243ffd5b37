// TEST INFRASTRUCTURE ONLY -- error slot for libvrt_dropin.so (ref_harness_scene.cpp without ref_harness.cpp).
#include <string>
static thread_local std::string g_err;
extern "C" void vrtref_set_error(const char *msg) { g_err = msg; }
extern "C" const char *vrtref_last_error() { return g_err.c_str(); }
extern "C" int vrtref_omp_max_threads() { return 1; }
