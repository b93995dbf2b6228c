// Integer-pipe calibration for the MSM/NTT rooflines: throughput of
// IMAD, IMAD.WIDE (64-bit accumulate), the carry-chained IMAD.WIDE.X pairs, and of the field /
// curve primitives built from them.  Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>
#include "../renegade_b200/csrc/ec.cuh"
using namespace b200;

constexpr int ITERS = 4096;

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = r[i] * a + b;
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t x = (uint32_t)r[i] ^ a;
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(x), "r"(b));
        }
    uint64_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide_x(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t e[8];
    for (int i = 0; i < 8; ++i) e[i] = threadIdx.x + i;
#ifdef __CUDA_ARCH__
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        ptx::wmad_cc(e[0], e[1], a, e[7]);
        ptx::wmadc_cc(e[2], e[3], b, e[1]);
        ptx::wmadc_cc(e[4], e[5], a, e[3]);
        ptx::wmadc_cc(e[6], e[7], b, e[5]);
        ptx::wmad_cc(e[1], e[2], a, e[0]);
        ptx::wmadc_cc(e[3], e[4], b, e[2]);
        ptx::wmadc_cc(e[5], e[6], a, e[4]);
        ptx::wmadc_cc(e[7], e[0], b, e[6]);
    }
#endif
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (r[i] + a + b) ^ (r[i] >> 3);
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 pipe: the rate that bounds a DFMA-based (52-bit limb) multiplier
__global__ void k_dfma(double* out, double a, double b) {
    double r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 and integer pipes side by side: do DFMA and IMAD.WIDE issue concurrently?
__global__ void k_dfma_imad(double* out, double a, double b, uint32_t x, uint32_t y) {
    double r[4];
    uint64_t q[4];
    for (int i = 0; i < 4; ++i) { r[i] = threadIdx.x + i; q[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = fma(r[i], a, b);
            uint32_t z = (uint32_t)q[i] ^ x;
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[i]) : "r"(z), "r"(y));
        }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += r[i] + (double)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <bool WIDE>
__global__ void k_femul(fe* out) {
    fe a, b;
    for (int i = 0; i < 8; ++i) { a.l[i] = threadIdx.x * 7 + i; b.l[i] = blockIdx.x + 3 * i; }
    a.l[7] &= 0x0fffffff; b.l[7] &= 0x0fffffff;
    for (int it = 0; it < ITERS / 8; ++it) {
        if (WIDE) { a = fe_mul<FqCfg>(a, b); b = fe_mul<FqCfg>(b, a); }
        else { a = fe_mul_chain<FqCfg>(a, b); b = fe_mul_chain<FqCfg>(b, a); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = fe_add<FqCfg>(a, b);
}
__global__ void k_madd(g1_xyzz* out) {
    g1_affine p;
    p.x = fe_one<FqCfg>(); p.y = fe_from_u32<FqCfg>(2);
    g1_xyzz acc = g1_dbl_affine(p);
    for (int i = 0; i < (int)(threadIdx.x & 3); ++i) acc = g1_dbl(acc);
    for (int it = 0; it < ITERS / 16; ++it) acc = g1_add_mixed(acc, p);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 128);
    const double nthreads = (double)blocks * threads;
    float t1 = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    float t2 = time_ms([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)buf, 3, 5); });
    float t3 = time_ms([&] { k_imad_wide_x<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    float t4 = time_ms([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    float t5 = time_ms([&] { k_femul<true><<<blocks, threads>>>((fe*)buf); });
    float t6 = time_ms([&] { k_femul<false><<<blocks, threads>>>((fe*)buf); });
    float t7 = time_ms([&] { k_madd<<<blocks, 128>>>((g1_xyzz*)buf); });
    float t8 = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9); });
    float t9 = time_ms([&] { k_dfma_imad<<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 3, 5); });
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"sms\": %d, \"clock_khz\": %d, \"imad_Gops\": %.1f, \"imad_wide_Gops\": %.1f, \"imad_wide_x_Gops\": %.1f, "
           "\"iadd3_lop_Gops\": %.1f, \"fq_mul_wide_G\": %.3f, \"fq_mul_wordserial_G\": %.3f, \"xyzz_madd_G\": %.4f, "
           "\"dfma_Gops\": %.1f, \"dfma_plus_imad_wide_pairs_Gops\": %.1f}\n",
           sms, clk, nthreads * ITERS * 8 / t1 / 1e6, nthreads * ITERS * 8 / t2 / 1e6, nthreads * ITERS * 8 / t3 / 1e6,
           nthreads * ITERS * 8 * 2 / t4 / 1e6, nthreads * (ITERS / 8) * 2 / t5 / 1e6, nthreads * (ITERS / 8) * 2 / t6 / 1e6,
           (double)blocks * 128 * (ITERS / 16) / t7 / 1e6, nthreads * ITERS * 8 / t8 / 1e6, nthreads * ITERS * 4 / t9 / 1e6);
    return 0;
}
