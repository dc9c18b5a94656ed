// bulk_rate_probe.cu -- hardware probe (not product code): what one SM's TMA moves with 1-D bulk copies
// (cp.async.bulk global -> shared, completion on an mbarrier), all 148 SMs streaming disjoint chunks of a large
// buffer the way the polyphase FIR's loader warp does: chunk w = blockIdx.x + it * gridDim.x, a ring of S slots,
// each chunk as P concurrent copies.  A second warp "consumes" a slot the moment it lands.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(64, 1) probe(const uint8_t *src, long long nchunks, int B, int S, int P, unsigned *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar0 + 8u * s, 1); mbar_init(bar0 + 8u * (64 + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t s0 = smem_u32(smem);
    if (warp == 0) {
        long long it = 0;
        for (long long w = blockIdx.x; w < nchunks; w += gridDim.x, ++it) {
            const int slot = (int)(it % S);
            if (it >= S) mbar_wait(bar0 + 8u * (64 + slot), (uint32_t)((it / S - 1) & 1));
            if (lane == 0) mbar_arrive_expect_tx(bar0 + 8u * slot, (uint32_t)B);
            __syncwarp();
            const int pb = B / P;
            if (lane < P) tma_bulk_g2s(s0 + (uint32_t)(slot * B + lane * pb), src + w * (long long)B + lane * pb, (uint32_t)pb, bar0 + 8u * slot);
        }
    } else {
        long long it = 0;
        unsigned acc = 0;
        for (long long w = blockIdx.x; w < nchunks; w += gridDim.x, ++it) {
            const int slot = (int)(it % S);
            mbar_wait(bar0 + 8u * slot, (uint32_t)((it / S) & 1));
            acc += smem[slot * B + lane * 4];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8u * (64 + slot));
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

int main() {
    const size_t total = (size_t)1 << 30;
    uint8_t *d;
    unsigned *sink;
    CK(cudaMalloc(&d, total + (1 << 20)));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 1, total + (1 << 20)));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Cfg { int B, S, P; };
    const Cfg cfgs[] = {{21120, 1, 1}, {21120, 2, 1}, {21120, 3, 1}, {21120, 4, 1}, {21120, 8, 1}, {42240, 2, 1}, {42240, 4, 1},
                        {10560, 4, 1}, {10560, 8, 1}, {10560, 16, 1}, {5280, 8, 1}, {5280, 32, 1}, {2048, 16, 1}, {2048, 64, 1},
                        {21120, 4, 2}, {21120, 4, 4}, {21120, 4, 11}, {21120, 8, 4}, {42240, 4, 8}, {640, 64, 1}};
    printf("%8s %4s %3s %10s %12s %10s\n", "chunk_B", "ring", "P", "ms", "GB/s", "B/clk/SM@1.9");
    for (const Cfg &c : cfgs) {
        const long long nchunks = (long long)(total / c.B);
        const size_t smem = (size_t)c.B * c.S;
        if (smem > 216 * 1024) continue;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            probe<<<sms, 64, smem>>>(d, nchunks, c.B, c.S, c.P, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double gbs = (double)nchunks * c.B / ms * 1e-6;
            if (rep == 1) printf("%8d %4d %3d %10.4f %12.1f %10.2f\n", c.B, c.S, c.P, ms, gbs, gbs / sms / 1.9);
        }
    }
    return 0;
}
