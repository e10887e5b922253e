// Microbenchmark (measurement aid, not product code): peak issue rates on B200 of the
// instructions the Fr arithmetic is made of, and of the Montgomery multiply itself.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../halo2_aes_b200/csrc \
//        -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "fr.cuh"
using namespace h2;

#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_imad_wide(uint32_t *out, int iters, uint32_t m) {
    uint64_t acc[CHAINS];
    uint32_t a = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = (uint64_t)(uint32_t)acc[c] * (a + c) + (acc[c] >> 7) + m;
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_imad32(uint32_t *out, int iters, uint32_t m) {
    uint32_t acc[CHAINS];
    uint32_t a = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = acc[c] * (a + c) + m;
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double m) {
    double acc[CHAINS];
    double a = 1.0 + 1e-9 * threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, m);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// CHAINS independent Montgomery multiply chains per thread; INL: inline copy vs the out-of-line call
__device__ __forceinline__ fr_t mul_inline(const fr_t &a, const fr_t &b) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fr_mul_ptx.inc"
    fr_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
    fr_t o = {{r0, r1, r2, r3, r4, r5, r6, r7}};
    return o;
}
template <int CHAINS, bool INL>
__global__ void __launch_bounds__(128) k_frmul(fr_t *out, int iters) {
    fr_t x[CHAINS], y = fr_r2();
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = fr_one();
        x[c].v[0] ^= threadIdx.x + 977 * c;
    }
    y.v[1] ^= blockIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = INL ? mul_inline(x[c], y) : fr_mul_call(x[c], y);
    }
    fr_t s = x[0];
#pragma unroll
    for (int c = 1; c < CHAINS; ++c) s = fr_add(s, x[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void *buf; CK(cudaMalloc(&buf, 64 << 20));
    const int iters = 4096;
    printf("SMs %d\n", sms);
    {
        int grid = sms * 8; double ops;
        float ms = time_ms([&] { k_imad_wide<8><<<grid, 256>>>((uint32_t *)buf, iters, 3); });
        ops = (double)grid * 256 * iters * 8;
        printf("IMAD.WIDE (8 chains)   %8.3f ms  %7.1f Gop/s  %.1f lanes/clk/SM\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
        ms = time_ms([&] { k_imad32<8><<<grid, 256>>>((uint32_t *)buf, iters, 3); });
        printf("IMAD 32   (8 chains)   %8.3f ms  %7.1f Gop/s  %.1f lanes/clk/SM\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
        ms = time_ms([&] { k_dfma<8><<<grid, 256>>>((double *)buf, iters, 0.5); });
        printf("DFMA      (8 chains)   %8.3f ms  %7.1f Gop/s  %.1f lanes/clk/SM\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
    }
    const int it2 = 512;
#define RUN(CH, INL, OCC)                                                                        \
    {                                                                                            \
        int grid = sms * OCC;                                                                    \
        float ms = time_ms([&] { k_frmul<CH, INL><<<grid, 128>>>((fr_t *)buf, it2); });          \
        double muls = (double)grid * 128 * it2 * CH;                                             \
        printf("fr_mul %s chains=%d ctas/SM=%2d  %8.3f ms  %7.2f Gmul/s\n", INL ? "inline" : "call  ", CH, OCC, ms, muls / ms / 1e6); \
    }
    RUN(1, false, 4) RUN(1, false, 8) RUN(1, false, 16)
    RUN(2, false, 4) RUN(2, false, 8)
    RUN(1, true, 4) RUN(1, true, 8) RUN(1, true, 16)
    RUN(2, true, 4) RUN(2, true, 8)
    RUN(4, true, 4)
    return 0;
}
