// fp64_rate_probe.cu -- what the FP64 pipe of one B200 SM sustains for the resampler's instruction mix.
// The sinc resampler accumulates acc = acc + c * x in f64 with SEPARATE multiply and add (the CPU restatement does
// not contract, so neither may the GPU).  This probe times R*CH independent chains per thread of (a) DMUL + DADD and
// (b) DFMA, with 4..32 warps per SM, and prints warp-instructions per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate_probe fp64_rate_probe.cu && ./fp64_rate_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(double *out, int iters, long long *cyc) {
    extern __shared__ double sm[];
    double acc[10], c[5], x[2];
    for (int i = 0; i < 10; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 5; ++i) c[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    x[0] = 0.999999; x[1] = 1.000001;
    if (threadIdx.x == 0) sm[0] = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            double p[10];
            // flip the exponent's lowest bit (c <-> 2c) so that ptxas cannot hoist the loop-invariant products
#pragma unroll
            for (int r = 0; r < 5; ++r) c[r] = __longlong_as_double(__double_as_longlong(c[r]) ^ (1LL << 52));
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(p[r * 2 + ch]) : "d"(c[r]), "d"(x[ch]));
#pragma unroll
            for (int i = 0; i < 10; ++i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(p[i]));
        } else {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[r * 2 + ch]) : "d"(c[r]), "d"(x[ch]));
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[r * 2 + ch]) : "d"(c[r]), "d"(x[ch]));
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 10; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 8 * 128 * sizeof(double));
    cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
    const int iters = 20000;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 2; ++mode)
        for (int bps : {1, 2, 3, 4, 6, 8}) {
            const int smem = (220 * 1024 / bps) & ~1023;
            auto kern = mode == 0 ? k<0> : k<1>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            kern<<<148 * bps, 128, smem>>>(out, iters, cyc);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            kern<<<148 * bps, 128, smem>>>(out, iters, cyc);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[148 * 8]; cudaMemcpy(h, cyc, 148 * bps * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148 * bps; ++i) avg += h[i]; avg /= 148 * bps;
            const double winstr = 20.0 * iters * 4 * bps;  // warp-instructions per SM
            // clock64 deltas turned out unreliable under this load (they imply > 4 instr/cycle); the event time and the
            // nominal max clock give the conservative figure
            printf("%s warps/SM %2d: %.3f ms, %.3f fp64 warp-instr/cycle/SM at %d MHz (clock64: %.0f cycles) (%s)\n",
                   mode ? "DFMA     " : "DMUL+DADD", 4 * bps, ms, winstr / (ms * 1e-3 * clk * 1e3), clk / 1000, avg,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
