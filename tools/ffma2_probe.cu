// ffma2_probe.cu — issue cost of the packed FP32 FMA (FFMA2, fma.rn.f32x2) on sm_100a against the scalar FFMA, alone and
// interleaved with integer ALU work (the k_extend node step mixes both).  Prints FMA/clk/SM and warp-instructions/clk/SM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void ffma2(float &r0, float &r1, float b, float c) {
    unsigned long long a, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(r0), "f"(r1));
    unsigned long long bb, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(d));
}

template <int MODE> __global__ void __launch_bounds__(256) probe(float *out, float a, float b, int iters, uint32_t m) {
    float x[16];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = static_cast<float>(threadIdx.x + i) * 1e-3f;
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 7u + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) ffma2(x[i], x[i + 1], a, b);
        }
        if (MODE >= 2) {  // 8 integer ALU instructions (LOP3) next to 16 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = (u[i] & m) ^ (u[(i + 1) & 7] | 0x55u);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i]);
    if (s == 123.456f) out[0] = s;
}

template <int MODE> static void run(const char *name, float *out, int sms, double mhz, int fma_instr, int other) {
    const int iters = 20000, ctas = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<ctas, 256>>>(out, 1.0001f, 0.5f, 100, 0xFFFFu);
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, 1.0001f, 0.5f, iters, 0xFFFFu);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    const double warps = 8.0 * 8.0;  // per SM
    printf("  %-22s %8.3f ms  %6.1f FMA/clk/SM  %5.2f warp-instr/clk/SM\n", name, ms, warps * 32 * 16.0 * iters / cycles,
           warps * (fma_instr + other) * double(iters) / cycles);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, 4);
    printf("%s %d SMs %.0f MHz\n", prop.name, prop.multiProcessorCount, khz / 1000.0);
    run<0>("FFMA x16", out, prop.multiProcessorCount, khz / 1000.0, 16, 0);
    run<1>("FFMA2 x8", out, prop.multiProcessorCount, khz / 1000.0, 8, 0);
    run<2>("FFMA x16 + LOP3 x8", out, prop.multiProcessorCount, khz / 1000.0, 16, 8);
    run<3>("FFMA2 x8 + LOP3 x8", out, prop.multiProcessorCount, khz / 1000.0, 8, 8);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
