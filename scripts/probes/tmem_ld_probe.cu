// tmem_ld_probe.cu -- hardware probe (not product code): TMEM -> register read bandwidth of one SM for the load shapes
// the FIR epilogue can use, with 4 / 8 / 16 warps loading back to back (each warp its own lane quadrant).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int SHAPE>  // 0: 16x256b.x2 (8 regs)  1: 16x256b.x4 (16 regs)  2: 32x32b.x16 (16 regs)  3: 32x32b.x32
__device__ __forceinline__ uint32_t ld(uint32_t ta) {
    uint32_t v[32];
    if (SHAPE == 0)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta));
    else if (SHAPE == 1)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta));
    else if (SHAPE == 2)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta));
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(ta));
    return v[0];
}
template <int SHAPE> __host__ __device__ constexpr int bytes_per_ld() { return (SHAPE == 0 ? 8 : SHAPE == 3 ? 32 : 16) * 32 * 4; }

template <int SHAPE, int PER_WAIT>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long *cycles, uint32_t *sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t ta = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < PER_WAIT; ++j) acc += ld<SHAPE>(ta + ((j * 32) & 255));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

template <int SHAPE, int PER_WAIT>
void run(const char *name, int warps, long long *dc, uint32_t *ds) {
    const int iters = 2000;
    probe<SHAPE, PER_WAIT><<<1, warps * 32, 0>>>(iters, dc, ds);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    const double bytes = (double)iters * PER_WAIT * bytes_per_ld<SHAPE>() * warps;
    printf("%-14s %2d warps, %d loads per wait: %6.1f B/cycle/SM  (%.1f cycles per load per warp)\n", name, warps, PER_WAIT,
           bytes / cyc, (double)cyc / (iters * PER_WAIT));
}

int main() {
    long long *dc; uint32_t *ds;
    CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&ds, 4096));
    for (int w : {4, 8, 16}) {
        run<0, 6>("16x256b.x2", w, dc, ds);
        run<1, 3>("16x256b.x4", w, dc, ds);
        run<2, 3>("32x32b.x16", w, dc, ds);
        run<3, 2>("32x32b.x32", w, dc, ds);
    }
    run<0, 1>("16x256b.x2", 16, dc, ds);
    run<3, 1>("32x32b.x32", 16, dc, ds);
    printf("tmem ld probe done\n");
    return 0;
}
