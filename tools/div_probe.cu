// Exhaustive probe (not part of the product): is a SHORTER sequence than the compiler's in-range division (MUFU.RCP + 5 FFMA, div_fast() in
// csrc/vrt_march.cuh) still equal to div.rn.f32 for the marcher's fixed numerator 0x42000000p0f over the whole fast range of |dir|^2?
//   S1: r = rcp(d); q = N * r; e = fma(-d, q, N); q = fma(r, e, q)                      (MUFU + 3)
//   S2: S1 with a second correction step                                                (MUFU + 5, for comparison)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/div_probe.cu -o tools/div_probe ; run on the GPU box
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float rcp_approx(float d) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d)); return r; }
__device__ __forceinline__ float s1(float d)
{
    const float N = 0x42000000p0f;
    const float r = rcp_approx(d);
    const float q = __fmul_rn(N, r);
    const float e = __fmaf_rn(-d, q, N);
    return __fmaf_rn(r, e, q);
}
__device__ __forceinline__ float s2(float d)
{
    const float N = 0x42000000p0f;
    const float r = rcp_approx(d);
    float q = __fmul_rn(N, r);
    float e = __fmaf_rn(-d, q, N);
    q = __fmaf_rn(r, e, q);
    e = __fmaf_rn(-d, q, N);
    return __fmaf_rn(r, e, q);
}
__global__ void probe(uint32_t first, uint32_t count, unsigned long long *bad, uint32_t *examples)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const float d = __uint_as_float(first + (uint32_t)i);
        const uint32_t want = __float_as_uint(__fdiv_rn(0x42000000p0f, d));
        if (__float_as_uint(s1(d)) != want) { const unsigned long long k = atomicAdd(&bad[0], 1ull); if (k < 16) examples[k] = first + (uint32_t)i; }
        if (__float_as_uint(s2(d)) != want) atomicAdd(&bad[1], 1ull);
    }
}
int main()
{
    unsigned long long *bad; uint32_t *ex;
    cudaMallocManaged(&bad, 16); cudaMallocManaged(&ex, 64);
    bad[0] = bad[1] = 0;
    const uint32_t first = 0x48000000u, count = 0x28000000u;      // [2^17, 2^97): div_is_fast_unit()
    probe<<<148 * 16, 256>>>(first, count, bad, ex);
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"range\": \"[2^17, 2^97)\", \"values\": %u, \"s1_mismatches\": %llu, \"s2_mismatches\": %llu, \"cuda\": \"%s\", \"s1_examples\": [", count, bad[0], bad[1], cudaGetErrorString(e));
    for (unsigned long long k = 0; k < bad[0] && k < 16; ++k) printf("%s\"0x%08x\"", k ? ", " : "", ex[k]);
    printf("]}\n");
    return 0;
}
