// pipe_bench.cu -- what does ONE warp-instruction of each kind cost a B200 scheduler, and do the pipes overlap?
//
// The marcher's roofline counts warp-instructions against SMs x 4 schedulers x clock, i.e. it assumes that every
// instruction costs one issue cycle.  This micro-benchmark measures that assumption for the instruction kinds of the step loop, at the
// marcher's occupancy (8 resident warps per scheduler, 1024 threads per SM): every warp runs an unrolled body of register-only
// instructions on 16 independent chains (8 for packed ops), and the result is printed as scheduler cycles per warp-instruction.
// Mixed bodies interleave two kinds 1:1; if the two pipes overlapped, a mix would cost max(a, b) per pair, if they share the issue
// path it costs a + b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu      (tools/sass check: cuobjdump -sass)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { FFMA, FADD, FMUL, FFMA2, FMUL2, LOP3, IADD3, ISETP_SEL, PRMT, SHF, IMAD, I2F, MUFU, NKIND };
static const char *kNames[NKIND] = {"FFMA", "FADD", "FMUL", "FFMA2", "FMUL2", "LOP3", "IADD3", "ISETP+SEL", "PRMT", "SHF", "IMAD", "I2F.U16", "MUFU.RCP"};
static const int kInstr[NKIND]   = {1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1};      // SASS instructions per op
static const bool kPacked[NKIND] = {false, false, false, true, true, false, false, false, false, false, false, false, false};

struct State
{
    float a[16];
    unsigned long long p[8];
    uint32_t w[16];
};

template <int K>
__device__ __forceinline__ void op(State &s, int i, float x, float y, uint32_t u, unsigned long long xy)
{
    if (K == FFMA)  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s.a[i]) : "f"(x), "f"(y));
    if (K == FADD)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s.a[i]) : "f"(y));
    if (K == FMUL)  asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s.a[i]) : "f"(x));
    if (K == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(s.p[i & 7]) : "l"(xy));
    if (K == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(s.p[i & 7]) : "l"(xy));
    if (K == LOP3)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(s.w[i]) : "r"(u), "r"(s.w[(i + 1) & 15]));
    if (K == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(s.w[i]) : "r"(s.w[(i + 1) & 15]));
    if (K == ISETP_SEL) asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %2, %0, q; }" : "+r"(s.w[i]) : "r"(u), "r"(s.w[(i + 1) & 15]));
    if (K == PRMT)  asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(s.w[i]) : "r"(s.w[(i + 1) & 15]));
    if (K == SHF)   asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(s.w[i]) : "r"(s.w[(i + 1) & 15]), "r"(u));
    if (K == IMAD)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s.w[i]) : "r"(u), "r"(s.w[(i + 1) & 15]));
    if (K == I2F)   asm volatile("{ .reg .u16 h; mov.b32 {h, _}, %1; cvt.rn.f32.u16 %0, h; }" : "=f"(s.a[i]) : "r"(s.w[i]));
    if (K == MUFU)  asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s.a[i]));
}

// body: 16 ops of kind A, or 8 of A interleaved with 8 of B
template <int A, int B>
__global__ void __launch_bounds__(256, 4) bench(float x, float y, uint32_t u, int iters, float *sink)
{
    State s;
    const unsigned long long xy = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
#pragma unroll
    for (int i = 0; i < 16; ++i) { s.a[i] = x * (float)(threadIdx.x + i + 1); s.w[i] = u + threadIdx.x * 17u + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) s.p[i] = xy + (unsigned long long)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            if (B < 0)
            {
#pragma unroll
                for (int i = 0; i < 16; ++i) op<A>(s, i, x, y, u, xy);
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 8; ++i) { op<A>(s, i, x, y, u, xy); op<(B < 0 ? 0 : B)>(s, i + 8, x, y, u, xy); }
            }
        }
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += s.a[i] + (float)s.w[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) t += (float)(s.p[i] & 0xFFFF);
    if (t == 12345.678f) sink[0] = t;
}

static int g_sms = 0;
static double g_mhz = 0;
static float *g_sink = nullptr;

template <int A, int B>
static double run()
{
    const int iters = 8000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    bench<A, B><<<g_sms * 4, 256>>>(1.0001f, 0.5f, 0x12345u, 50, g_sink);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    bench<A, B><<<g_sms * 4, 256>>>(1.0001f, 0.5f, 0x12345u, iters, g_sink);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double cycles = ms * 1e-3 * g_mhz * 1e6;                  // per SM = per scheduler (all four run the same)
    const double ops_per_warp = (double)iters * 4.0 * 16.0;         // 16 ops per body, 4 bodies per iteration
    const double per_op = cycles / (8.0 * ops_per_warp);            // 8 warps per scheduler
    if (B < 0)
        printf("{\"kind\": \"%s\", \"scheduler_cycles_per_warp_instruction\": %.3f, \"sass_instructions_per_op\": %d}\n", kNames[A], per_op / kInstr[A], kInstr[A]);
    else
        printf("{\"mix\": \"%s + %s (1:1)\", \"scheduler_cycles_per_pair\": %.3f}\n", kNames[A], kNames[B < 0 ? 0 : B], per_op * 2.0);
    return per_op;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    g_sms = pr.multiProcessorCount; g_mhz = khz / 1000.0;
    cudaMalloc(&g_sink, 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"note\": \"8 warps per scheduler, 16 independent chains per thread (8 for packed); cycles at the nominal max clock\"}\n",
           pr.name, g_sms, g_mhz);
    run<FFMA, -1>(); run<FADD, -1>(); run<FMUL, -1>(); run<FFMA2, -1>(); run<FMUL2, -1>(); run<LOP3, -1>(); run<IADD3, -1>(); run<ISETP_SEL, -1>();
    run<PRMT, -1>(); run<SHF, -1>(); run<IMAD, -1>(); run<I2F, -1>(); run<MUFU, -1>();
    run<FFMA, LOP3>(); run<FFMA, IADD3>(); run<FFMA2, LOP3>(); run<FFMA2, IADD3>(); run<FFMA2, FFMA>(); run<FFMA, I2F>(); run<FFMA2, I2F>(); run<FFMA, MUFU>();
    run<LOP3, I2F>(); run<FFMA, IMAD>(); run<FFMA2, FADD>(); run<IADD3, LOP3>();
    return 0;
}
