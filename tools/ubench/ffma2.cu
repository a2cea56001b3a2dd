// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) and VIADDMNMX.S16x2 on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 4096
#define UNROLL 16
__global__ void k_ffma(float *out, float a, float b)
{
    float x[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) x[j] = fmaf(x[j], a, b);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < UNROLL; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b)
{
    unsigned long long x[UNROLL];
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int j = 0; j < UNROLL; j++) { float v = threadIdx.x + j; asm("mov.b64 %0, {%1, %1};" : "=l"(x[j]) : "f"(v)); }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[j]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < UNROLL; j++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[j])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_viaddmax(uint32_t *out, uint32_t a)
{
    uint32_t x[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) x[j] = __viaddmax_s16x2(x[j], a, 0u);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < UNROLL; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int sms = p.multiProcessorCount, clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8, threads = 256;       // 64 warps / SM
    float *out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    const double ninst = (double)blocks * (threads / 32) * ITERS * UNROLL;      // warp instructions
    float t1 = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t2 = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t3 = timeit([&] { k_viaddmax<<<blocks, threads>>>((uint32_t *)out, 0x00010001u); });
    const double cyc = clk_khz * 1e3;
    printf("SMs %d clock %.0f MHz\n", sms, clk_khz / 1e3);
    printf("FFMA      : %.3f ms  %.3f warp-inst/clk/SM  (%.1f TFLOP/s)\n", t1, ninst / (t1 * 1e-3) / cyc / sms, ninst * 64 / (t1 * 1e-3) / 1e12);
    printf("FFMA2     : %.3f ms  %.3f warp-inst/clk/SM  (%.1f TFLOP/s)\n", t2, ninst / (t2 * 1e-3) / cyc / sms, ninst * 128 / (t2 * 1e-3) / 1e12);
    printf("VIADDMNMX : %.3f ms  %.3f warp-inst/clk/SM  (%.1f T s16-lane-ops/s)\n", t3, ninst / (t3 * 1e-3) / cyc / sms, ninst * 64 / (t3 * 1e-3) / 1e12);
    return 0;
}
