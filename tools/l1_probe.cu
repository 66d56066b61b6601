// l1_probe.cu — what a divergent node fetch costs the L1 data pipe on sm_100a (k_extend).
// Every warp runs ITER iterations; per iteration each lane fetches one 64-byte "node" from a table with a hashed
// index, in one of several ways.  Reported: SM cycles per warp-iteration per SM (all resident warps together), i.e.
// the reciprocal throughput of the load pattern when nothing else limits it.
//   own256   2 x LDG.256 of the lane's own node (what k_extend does)
//   pair256  2 x LDG.256, the two lanes of a pair fetch the two halves of ONE node per instruction (no exchange)
//   pairx    pair256 + the 8 SHFL + 16 SEL that hand every lane its own node
//   split256 2 x LDG.256, the two halves of a node in two arrays of 32-byte stride (all four sector banks per instruction)
//   split256na  the same with L1::no_allocate
//   one256   1 x LDG.256 (a 32-byte node)
//   own128   4 x LDG.128 of the lane's own node
//   quad128  4 x LDG.128, four lanes fetch the four quarters of ONE node per instruction
//   lds64    1 x LDS.64 at [entry][thread] with a random entry (the traversal stack's pop)
//   shfl8    8 x SHFL only
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l1_probe tools/l1_probe.cu ; run: build/l1_probe
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

struct f8 {
    float a, b, c, d, e, f, g, h;
};
__device__ __forceinline__ f8 ld8(const void *p) {
    f8 v;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a), "=f"(v.b), "=f"(v.c), "=f"(v.d), "=f"(v.e), "=f"(v.f), "=f"(v.g), "=f"(v.h)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ f8 ld8na(const void *p) {  // L1::no_allocate
    f8 v;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a), "=f"(v.b), "=f"(v.c), "=f"(v.d), "=f"(v.e), "=f"(v.f), "=f"(v.g), "=f"(v.h)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld4(const void *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float sum8(const f8 &v) { return ((v.a + v.b) + (v.c + v.d)) + ((v.e + v.f) + (v.g + v.h)); }
__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

template <int MODE> __global__ void __launch_bounds__(128) probe(const char *table, uint32_t mask, int iters, int active, float *out) {
    __shared__ uint2 s_stack[16 * 128];
    for (int i = threadIdx.x; i < 16 * 128; i += 128) s_stack[i] = make_uint2(i, i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.0f;
    if (static_cast<int>(lane) >= active) return;  // fewer active lanes (a multiple of 4)
    const uint32_t amask = active >= 32 ? 0xFFFFFFFFu : ((1u << active) - 1u);
    for (int it = 0; it < iters; ++it) {
        const uint32_t h_own = hash(gw * 0x9E3779B9u + it * 32u + lane) & mask;
        const uint32_t h_even = hash(gw * 0x9E3779B9u + it * 32u + (lane & ~1u)) & mask;
        const uint32_t h_odd = hash(gw * 0x9E3779B9u + it * 32u + (lane | 1u)) & mask;
        if (MODE == 0) {
            const char *p = table + static_cast<size_t>(h_own) * 64;
            acc += sum8(ld8(p)) + sum8(ld8(p + 32));
        } else if (MODE == 1 || MODE == 2) {
            const uint32_t par = lane & 1u;
            const f8 x = ld8(table + static_cast<size_t>(h_even) * 64 + par * 32);
            const f8 y = ld8(table + static_cast<size_t>(h_odd) * 64 + par * 32);
            if (MODE == 1) {
                acc += sum8(x) + sum8(y);
            } else {
                f8 s, k, r;
#define SELX(m) s.m = par ? x.m : y.m; k.m = par ? y.m : x.m; r.m = __shfl_xor_sync(amask, s.m, 1);
                SELX(a) SELX(b) SELX(c) SELX(d) SELX(e) SELX(f) SELX(g) SELX(h)
                acc += sum8(k) - sum8(r);
            }
        } else if (MODE == 8) {  // the two halves of a node in two separate arrays, 32-byte stride each
            const char *pa = table + static_cast<size_t>(h_own) * 32;
            const char *pb = table + (static_cast<size_t>(mask) + 1) * 32 + static_cast<size_t>(h_own) * 32;
            acc += sum8(ld8(pa)) + sum8(ld8(pb));
        } else if (MODE == 9) {  // split256 with L1::no_allocate
            const char *pa = table + static_cast<size_t>(h_own) * 32;
            const char *pb = table + (static_cast<size_t>(mask) + 1) * 32 + static_cast<size_t>(h_own) * 32;
            acc += sum8(ld8na(pa)) + sum8(ld8na(pb));
        } else if (MODE == 3) {
            acc += sum8(ld8(table + static_cast<size_t>(h_own) * 32));
        } else if (MODE == 4) {
            const char *p = table + static_cast<size_t>(h_own) * 64;
            const float4 a = ld4(p), b = ld4(p + 16), c = ld4(p + 32), d = ld4(p + 48);
            acc += (a.x + b.y) + (c.z + d.w) + (a.y + a.z + a.w + b.x + b.z + b.w + c.x + c.y + c.w + d.x + d.y + d.z);
        } else if (MODE == 5) {
            const uint32_t q = lane & 3u;
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t hj = hash(gw * 0x9E3779B9u + it * 32u + (lane & ~3u) + j) & mask;
                const float4 a = ld4(table + static_cast<size_t>(hj) * 64 + q * 16);
                t += (a.x + a.y) + (a.z + a.w);
            }
            acc += t;
        } else if (MODE == 6) {
            const uint2 e = s_stack[(h_own & 15u) * 128 + threadIdx.x];
            acc += __uint_as_float(e.x) + __uint_as_float(e.y);
        } else if (MODE == 7) {
            float t = __uint_as_float(h_own);
#pragma unroll
            for (int j = 0; j < 8; ++j) t += __shfl_xor_sync(amask, t, 1);
            acc += t;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int MODE> static void run(const char *name, const char *table, uint32_t mask, int active, float *out, int sms, float mhz) {
    const int iters = 2000, ctas = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<ctas, 128>>>(table, mask, 200, active, out);
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 128>>>(table, mask, iters, active, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 8 CTAs x 4 warps x iters warp-iterations
    const double cyc = ms * 1e-3 * mhz * 1e6 / (8.0 * 4.0 * iters);
    printf("  %-8s active=%2d  %7.3f ms  %7.2f cycles per warp-iteration per SM\n", name, active, ms, cyc);
}

int main(int argc, char **argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float mhz = khz / 1000.0f;
    printf("%s, %d SMs, %.0f MHz (attribute)\n", prop.name, prop.multiProcessorCount, mhz);
    float *out;
    cudaMalloc(&out, 4);
    const size_t sizes[3] = {size_t(32) << 10, size_t(32) << 20, size_t(1) << 30};
    const char *names[3] = {"32 KB table (L1-resident)", "32 MB table (L2-resident)", "1 GB table (HBM)"};
    const int only_size = argc > 1 ? atoi(argv[1]) : -1, only_mode = argc > 2 ? atoi(argv[2]) : -1;  // for ncu: one pattern
    for (int s = 0; s < 3; ++s) {
        if (only_size >= 0 && s != only_size) continue;
        char *table;
        cudaMalloc(&table, sizes[s]);
        cudaMemset(table, 0, sizes[s]);
        const uint32_t mask = static_cast<uint32_t>(sizes[s] / 64 - 1);
        printf("%s\n", names[s]);
        if (only_mode >= 0) {
            if (only_mode == 0) run<0>("own256", table, mask, 32, out, prop.multiProcessorCount, mhz);
            if (only_mode == 8) run<8>("split256", table, mask, 32, out, prop.multiProcessorCount, mhz);
            if (only_mode == 1) run<1>("pair256", table, mask, 32, out, prop.multiProcessorCount, mhz);
            if (only_mode == 3) run<3>("one256", table, mask, 32, out, prop.multiProcessorCount, mhz);
            if (only_mode == 4) run<4>("own128", table, mask, 32, out, prop.multiProcessorCount, mhz);
            cudaFree(table);
            continue;
        }
        for (int active : {32, 20}) {
            run<0>("own256", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<8>("split256", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<9>("split256na", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<1>("pair256", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<2>("pairx", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<3>("one256", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<4>("own128", table, mask, active, out, prop.multiProcessorCount, mhz);
            run<5>("quad128", table, mask, active, out, prop.multiProcessorCount, mhz);
            if (s == 0) {
                run<6>("lds64", table, mask, active, out, prop.multiProcessorCount, mhz);
                run<7>("shfl8", table, mask, active, out, prop.multiProcessorCount, mhz);
            }
        }
        cudaFree(table);
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
