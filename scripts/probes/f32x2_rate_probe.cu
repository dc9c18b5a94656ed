// f32x2_rate_probe.cu -- issue rate of the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2 from PTX
// add.f32x2 / mul.f32x2 / fma.rn.f32x2) against their scalar forms, and of a complex multiply written both ways.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate_probe f32x2_rate_probe.cu && ./f32x2_rate_probe
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 v) { return *reinterpret_cast<u64 *>(&v); }
__device__ __forceinline__ float2 up(u64 v) { return *reinterpret_cast<float2 *>(&v); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, int iters) {
    float2 v[8], w = make_float2(1.0000001f, 0.9999999f);
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, 1.0f - i * 1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { v[i].x = __fadd_rn(v[i].x, w.x); v[i].y = __fadd_rn(v[i].y, w.y); }           // 2 FADD
            if (MODE == 1) v[i] = up(add2(pk(v[i]), pk(w)));                                                   // 1 FADD2
            if (MODE == 2) { v[i].x = __fmaf_rn(v[i].x, w.x, w.y); v[i].y = __fmaf_rn(v[i].y, w.y, w.x); } // 2 FFMA
            if (MODE == 3) v[i] = up(fma2(pk(v[i]), pk(w), pk(w)));                                            // 1 FFMA2
            if (MODE == 4) {  // complex multiply, scalar: 2 FMUL + 2 FFMA
                const float2 a = v[i];
                v[i] = make_float2(__fmaf_rn(a.x, w.x, -__fmul_rn(a.y, w.y)), __fmaf_rn(a.x, w.y, __fmul_rn(a.y, w.x)));
            }
            if (MODE == 5) {  // complex multiply, packed: FMUL2 + FFMA2 on broadcast pairs
                const float2 a = v[i];
                const u64 t = mul2(pk(make_float2(a.y, a.y)), pk(make_float2(-w.y, w.x)));
                v[i] = up(fma2(pk(make_float2(a.x, a.x)), pk(w), t));
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
    for (int i = 0; i < 8; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float2 *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    const char *names[6] = {"2 x FADD ", "FADD2    ", "2 x FFMA ", "FFMA2    ", "cmul 2 FMUL + 2 FFMA", "cmul FMUL2 + FFMA2  "};
    void (*kern[6])(float2 *, int) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>};
    for (int m = 0; m < 6; ++m) {
        kern[m]<<<148 * 4, 256>>>(out, 10);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern[m]<<<148 * 4, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double elems = 8.0 * iters * 4 * 8;  // complex elements updated per SM, in warps (4 CTAs x 8 warps)
        printf("%s: %.3f ms  %.2f complex-element warp-updates/cycle/SM at %d MHz (%s)\n", names[m], ms,
               elems / (ms * 1e-3 * clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
