// Microbenchmark: issue rate of the Forward/Backward instruction mix on sm_100a -- FFMA whose coefficient is a uniform
// register filled by LDCU.128 from a __grid_constant__ kernel parameter (one LDCU.128 per R FFMAs), against plain FFMA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ldcu_mix ldcu_mix.cu && ./ldcu_mix
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define NC 64            // float4 coefficients per pass: 256 floats cannot stay in the 63 uniform registers
struct Coef { float4 c[NC]; };
template <int PER>       // FFMAs per coefficient quadruple: 4 (one per component) or 8 (two per component)
__global__ void k_mix(const __grid_constant__ Coef pc, float *out)
{
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const float4 c = pc.c[j];
            x[(4 * j + 0) & 15] = fmaf(x[(4 * j + 0) & 15], c.x, x[(4 * j + 5) & 15]);
            x[(4 * j + 1) & 15] = fmaf(x[(4 * j + 1) & 15], c.y, x[(4 * j + 6) & 15]);
            x[(4 * j + 2) & 15] = fmaf(x[(4 * j + 2) & 15], c.z, x[(4 * j + 7) & 15]);
            x[(4 * j + 3) & 15] = fmaf(x[(4 * j + 3) & 15], c.w, x[(4 * j + 8) & 15]);
            if (PER == 8) {
                x[(4 * j + 8) & 15] = fmaf(x[(4 * j + 8) & 15], c.x, x[(4 * j + 13) & 15]);
                x[(4 * j + 9) & 15] = fmaf(x[(4 * j + 9) & 15], c.y, x[(4 * j + 14) & 15]);
                x[(4 * j + 10) & 15] = fmaf(x[(4 * j + 10) & 15], c.z, x[(4 * j + 15) & 15]);
                x[(4 * j + 11) & 15] = fmaf(x[(4 * j + 11) & 15], c.w, x[(4 * j + 16) & 15]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_plain(float *out, float a)
{
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 4 * NC; j++) x[j & 15] = fmaf(x[j & 15], a, x[(j + 5) & 15]);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += x[j];
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
    Coef h;
    for (int j = 0; j < NC; j++) h.c[j] = make_float4(1.0f - 1e-6f * j, 1.0f - 2e-6f * j, 1.0f - 3e-6f * j, 1.0f - 4e-6f * j);
    const double cyc = clk_khz * 1e3;
    printf("SMs %d clock %.0f MHz; warp-instructions / clk / SM (FFMA only | FFMA + LDCU)\n", sms, clk_khz / 1e3);
    const int cfg[3][2] = {{6, 64}, {8, 128}, {8, 256}};       // 12, 32, 64 warps / SM
    for (int q = 0; q < 3; q++) {
        const int blocks = sms * cfg[q][0], threads = cfg[q][1];
        float *out; cudaMalloc(&out, (size_t)blocks * threads * 4);
        const double nw = (double)blocks * (threads / 32) * ITERS;
        float t0 = timeit([&] { k_plain<<<blocks, threads>>>(out, 0.999f); });
        float t4 = timeit([&] { k_mix<4><<<blocks, threads>>>(h, out); });
        float t8 = timeit([&] { k_mix<8><<<blocks, threads>>>(h, out); });
        const double r0 = nw * 4 * NC / (t0 * 1e-3) / cyc / sms;
        const double f4 = nw * 4 * NC / (t4 * 1e-3) / cyc / sms, a4 = nw * 5 * NC / (t4 * 1e-3) / cyc / sms;
        const double f8 = nw * 8 * NC / (t8 * 1e-3) / cyc / sms, a8 = nw * 9 * NC / (t8 * 1e-3) / cyc / sms;
        printf("%2d warps/SM: plain FFMA %.3f | 4 FFMA per LDCU.128: %.3f %.3f | 8 FFMA per LDCU.128: %.3f %.3f\n",
               cfg[q][0] * cfg[q][1] / 32, r0, f4, a4, f8, a8);
        cudaFree(out);
    }
    return 0;
}
