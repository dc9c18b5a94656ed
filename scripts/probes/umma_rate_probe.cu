// umma_rate_probe.cu -- hardware probe (not product code): issue-rate / operand-fetch bound of tcgen05.mma kind::i8
// M=128, K=32 B for several N, with NACC independent accumulators issued round-robin from ONE thread, unrolled.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo, int layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
}

template <int N, int NACC, int KS>
__global__ void __launch_bounds__(128, 1) rate(int groups, int pitch, int layout, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem) & 1023u)) & 1023u);
    uint8_t *sb = sa + 65536;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 + 131072; i += 128) sa[i] = (uint8_t)(i * 7 + 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (tid == 0) {
        const uint64_t ad0 = make_desc((uint32_t)__cvta_generic_to_shared(sa), 16, 8 * pitch, layout);
        const uint64_t bd0 = make_desc((uint32_t)__cvta_generic_to_shared(sb), 128, 256, 0);
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
#pragma unroll
                for (int a = 0; a < NACC; ++a)
                    mma_i8(tmem + a * N, ad0 + (uint64_t)((a * 128 * pitch + kk * 32) >> 4), bd0 + (uint64_t)((kk * N * 32) >> 4), idesc, kk > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb));
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mb), "r"(0u));
        cycles[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N, int NACC, int KS>
void run(const char *what, int pitch, int layout, long long *dc) {
    const int groups = 200;
    auto k = rate<N, NACC, KS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 131072 + 1024));
    k<<<1, 128, 65536 + 131072 + 1024>>>(groups, pitch, layout, dc);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cyc / (groups * KS * NACC);
    printf("%-8s N=%3d NACC=%d KS=%2d : %6.1f cycles/MMA  operand bytes/MMA %5d -> %5.1f B/cycle   outputs/cycle %.2f\n", what, N, NACC, KS,
           per, 4096 + N * 32, (4096 + N * 32) / per, 128.0 * (pitch / 2) / (per * KS));
}

int main() {
    long long *dc; CK(cudaMalloc(&dc, 8));
    run<48, 1, 5>("none", 16, 0, dc);  run<48, 4, 5>("none", 16, 0, dc); run<48, 4, 17>("none", 16, 0, dc);
    run<96, 1, 6>("sw32", 32, 6, dc);  run<96, 2, 6>("sw32", 32, 6, dc);  run<96, 1, 18>("sw32", 32, 6, dc); run<96, 2, 18>("sw32", 32, 6, dc);
    run<192, 1, 7>("sw64", 64, 4, dc); run<192, 2, 7>("sw64", 64, 4, dc); run<192, 1, 16>("sw64", 64, 4, dc); run<192, 2, 16>("sw64", 64, 4, dc);
    run<256, 1, 8>("sw64", 64, 4, dc); run<256, 2, 8>("sw64", 64, 4, dc);
    run<128, 1, 8>("sw64", 64, 4, dc); run<128, 2, 8>("sw64", 64, 4, dc); run<128, 4, 8>("sw64", 64, 4, dc);
    run<64, 4, 8>("none", 16, 0, dc);
    run<96, 2, 19>("sw64", 64, 4, dc); run<96, 1, 19>("sw64", 64, 4, dc); run<48, 4, 19>("sw64", 64, 4, dc); run<192, 1, 19>("sw64", 64, 4, dc);
    run<96, 2, 19>("sw32", 32, 6, dc);
    printf("rate probe done\n");
    return 0;
}
