// Microbenchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a (informs the Heisenberg inner loop).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0, p1, p2, p3, pa, pb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {  // 8 scalar FFMA
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            } else if (MODE == 1) {  // 4 packed FFMA2 (same flops)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            } else {  // 4 FFMA2 + 4 LOP3 (does the packed op free issue slots for the ALU pipe?)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
                unsigned u0 = __float_as_uint(x0), u1 = __float_as_uint(x1), u2 = __float_as_uint(x2), u3 = __float_as_uint(x3);
                u0 = (u0 ^ u1) & (u2 | 0x55u + u); u1 = (u1 ^ u2) & (u3 | 0x33u + u); u2 = (u2 ^ u3) & (u0 | 0x0fu + u); u3 = (u3 ^ u0) & (u1 | 0x77u + u);
                x0 = __uint_as_float(u0); x1 = __uint_as_float(u1); x2 = __uint_as_float(u2); x3 = __uint_as_float(u3);
            }
        }
    }
    float r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0)); r += lo + hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p1)); r += lo + hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2)); r += lo + hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p3)); r += lo + hi;
    if (r == 12345.678f) out[0] = r;
}
template <int MODE> void run(const char* name) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(out, 16, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * 16 * 8;
    printf("%-28s %.3f ms  %.2f Tfma/s (scalar-equivalent)\n", name, ms, fma / ms * 1e-9);
}
int main() { run<0>("8 FFMA"); run<1>("4 FFMA2"); run<2>("4 FFMA2 + ~8 LOP3"); run<0>("8 FFMA"); return 0; }
