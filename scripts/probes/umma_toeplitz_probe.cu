// umma_toeplitz_probe.cu -- hardware probe (not product code): does a tcgen05.mma shared-memory descriptor whose
// rows OVERLAP (row pitch 16/32/64/128 B inside one flat byte array) read the Toeplitz/Hankel operand we expect,
// for the no-swizzle and the 32/64/128-byte swizzle layouts, with kind::i8 (u8 x s8 -> s32)?  Also times
// back-to-back MMAs of several N to find the shared-memory operand-read bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_toeplitz_probe.cu && ./umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Cfg {
    int layout;      // descriptor layout_type: 0 none, 6 sw32, 4 sw64, 2 sw128
    int a_lbo, a_sbo, a_start;   // bytes
    int b_lbo, b_sbo;            // bytes (B always no-swizzle here)
    int N;
    int reps;        // MMAs issued (timing); result checked after reps accumulations when accumulate==0 on first
    int kind_f16;    // 0: i8, 1: f16
    int nacc;        // independent accumulators used round-robin (timing)
    int a_step;      // bytes added to the A start per MMA within a group of nacc (timing)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo, int layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version
    d |= (uint64_t)layout << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(const uint8_t *adata, int abytes, const uint8_t *bdata, int bbytes,
                                                Cfg c, int32_t *dout, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem) & 1023u)) & 1023u);  // 1024-aligned
    uint8_t *sb = sa + 32768;           // B region
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < abytes; i += 128) sa[i] = adata[i];
    for (int i = tid; i < bbytes; i += 128) sb[i] = bdata[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t ad0 = make_desc((uint32_t)__cvta_generic_to_shared(sa) + c.a_start, c.a_lbo, c.a_sbo, c.layout);
        const uint64_t bd = make_desc((uint32_t)__cvta_generic_to_shared(sb), c.b_lbo, c.b_sbo, 0);
        uint32_t idesc;
        if (c.kind_f16) idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
        else idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
        t0 = clock64();
        const int nacc = c.nacc > 0 ? c.nacc : 1;
        for (int r = 0; r < c.reps; ++r) {
            const uint32_t acc = r >= nacc;
            const uint32_t tmem = tmem_base_s + (uint32_t)((r % nacc) * c.N);
            const uint64_t ad = ad0 + (uint64_t)(((r % nacc) * c.a_step) >> 4);
            if (c.kind_f16)
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
            else
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb));
    }
    // everyone waits for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mb), "r"(0u));
        }
    }
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // read back: thread t of warp w holds lane 32w + t
    for (int col0 = 0; col0 < c.N; col0 += 8) {
        uint32_t v[8];
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + col0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int i = 0; i < 8; ++i) dout[tid * 256 + col0 + i] = (int32_t)v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static uint32_t swz(uint32_t L, int layout) {
    if (layout == 6) return L ^ (((L >> 7) & 1) << 4);
    if (layout == 4) return L ^ (((L >> 7) & 3) << 4);
    if (layout == 2) return L ^ (((L >> 7) & 7) << 4);
    return L;
}

int main() {
    const int ABYTES = 32768, BBYTES = 16384;
    std::vector<uint8_t> ha(ABYTES), hb(BBYTES);
    srand(1234);
    for (auto &v : ha) v = rand() & 255;
    uint8_t *da, *db; int32_t *dd; long long *dc;
    CK(cudaMalloc(&da, ABYTES)); CK(cudaMalloc(&db, BBYTES)); CK(cudaMalloc(&dd, 128 * 256 * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(da, ha.data(), ABYTES, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + BBYTES + 1024));
    struct T { const char *name; int layout, pitch, start, N; };
    const T tests[] = {
        {"none  P=16B start 0   ", 0, 16, 0, 48},    {"none  P=16B start 32  ", 0, 16, 32, 48},
        {"none  P=16B start 272 ", 0, 16, 272, 16},
        {"sw32  P=32B start 0   ", 6, 32, 0, 48},    {"sw32  P=32B start 32  ", 6, 32, 32, 48},
        {"sw32  P=32B start 128 ", 6, 32, 128, 48},  {"sw32  P=32B start 416 ", 6, 32, 416, 96},
        {"sw64  P=64B start 0   ", 4, 64, 0, 48},    {"sw64  P=64B start 32  ", 4, 64, 32, 48},
        {"sw64  P=64B start 64  ", 4, 64, 64, 48},   {"sw64  P=64B start 160 ", 4, 64, 160, 192},
        {"sw128 P=128B start 0  ", 2, 128, 0, 48},   {"sw128 P=128B start 32 ", 2, 128, 32, 48},
        {"sw128 P=128B start 128", 2, 128, 128, 48}, {"sw128 P=128B start 224", 2, 128, 224, 48},
    };
    for (const T &t : tests) {
        // B: N rows x 32 bytes (K-major, no swizzle): element (n,k) at (n/8)*256 + (k/16)*128 + (n%8)*16 + k%16
        std::vector<int8_t> B(t.N * 32);
        for (auto &v : B) v = (int8_t)((rand() & 255) - 128);
        std::fill(hb.begin(), hb.end(), 0);
        for (int n = 0; n < t.N; ++n)
            for (int k = 0; k < 32; ++k) hb[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + k % 16] = (uint8_t)B[n * 32 + k];
        CK(cudaMemcpy(db, hb.data(), BBYTES, cudaMemcpyHostToDevice));
        Cfg c{};
        c.layout = t.layout; c.a_start = t.start; c.N = t.N; c.reps = 1; c.kind_f16 = 0;
        c.a_lbo = (t.layout == 0) ? 16 : 16;   // K-direction chunk stride (ignored for swizzled K-major)
        c.a_sbo = 8 * t.pitch;
        c.b_lbo = 128; c.b_sbo = 256;
        CK(cudaMemset(dd, 0xff, 128 * 256 * 4));
        probe<<<1, 128, 32768 + BBYTES + 1024>>>(da, ABYTES, db, BBYTES, c, dd, dc);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> D(128 * 256);
        CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
        // model A: swizzle is a function of the absolute byte offset inside the 1024-aligned region
        // model B: swizzle is a function of the offset from the descriptor start
        int badA = 0, badB = 0, badFlat = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < t.N; ++n) {
                long long sA = 0, sB = 0, sF = 0;
                for (int k = 0; k < 32; ++k) {
                    const uint32_t L = (uint32_t)(t.start + t.pitch * r + k);
                    const uint32_t Lrel = (uint32_t)(t.pitch * r + k);
                    sA += (long long)ha[swz(L, t.layout)] * B[n * 32 + k];
                    sB += (long long)ha[t.start + swz(Lrel, t.layout)] * B[n * 32 + k];
                    sF += (long long)ha[L] * B[n * 32 + k];
                }
                const int32_t got = D[r * 256 + n];
                badA += got != (int32_t)sA; badB += got != (int32_t)sB; badFlat += got != (int32_t)sF;
            }
        printf("%s N=%3d : mismatches  modelA(abs-addr swizzle)=%d  modelB(start-relative)=%d  flat(no swizzle)=%d   D[0][0]=%d D[1][0]=%d\n",
               t.name, t.N, badA, badB, badFlat, D[0], D[256]);
    }
    // timing: cycles per MMA for several N, i8 and f16, no-swizzle overlapped rows
    const int Ns[] = {16, 48, 96, 192, 256};
    for (int kf = 0; kf < 2; ++kf)
        for (int N : Ns) {
            Cfg c{};
            c.layout = 0; c.a_start = 0; c.N = N; c.reps = 2000; c.kind_f16 = kf;
            c.a_lbo = 16; c.a_sbo = 128; c.b_lbo = 128; c.b_sbo = 256;
            probe<<<1, 128, 32768 + BBYTES + 1024>>>(da, ABYTES, db, BBYTES, c, dd, dc);
            CK(cudaDeviceSynchronize());
            long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
            printf("timing %s M=128 N=%3d K=32B : %.1f cycles per MMA (%d MMAs, 1 CTA)\n", kf ? "f16" : "i8 ", N, (double)cyc / c.reps, c.reps);
        }
    // same with a dense (non-overlapped) A: row pitch via SBO=512 (8 rows x 64 B would be sw64); here none-layout dense: lbo=128,sbo=256
    for (int N : Ns) {
        Cfg c{};
        c.layout = 0; c.a_start = 0; c.N = N; c.reps = 2000; c.kind_f16 = 0;
        c.a_lbo = 128; c.a_sbo = 256; c.b_lbo = 128; c.b_sbo = 256;
        probe<<<1, 128, 32768 + BBYTES + 1024>>>(da, ABYTES, db, BBYTES, c, dd, dc);
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
        printf("timing i8 dense-A M=128 N=%3d : %.1f cycles per MMA\n", N, (double)cyc / c.reps);
    }
    for (int nacc : {2, 4, 8})
        for (int N : {16, 48, 96, 192}) {
            if (nacc * N > 256) continue;
            Cfg c{};
            c.layout = 0; c.a_start = 0; c.N = N; c.reps = 4000; c.kind_f16 = 0; c.nacc = nacc; c.a_step = 2048;
            c.a_lbo = 16; c.a_sbo = 128; c.b_lbo = 128; c.b_sbo = 256;
            probe<<<1, 128, 32768 + BBYTES + 1024>>>(da, ABYTES, db, BBYTES, c, dd, dc);
            CK(cudaDeviceSynchronize());
            long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
            printf("timing i8 overlapped-A M=128 N=%3d, %d independent accumulators (A blocks 2 KB apart): %.1f cycles per MMA\n", N, nacc, (double)cyc / c.reps);
        }
    for (int layout : {6, 4})
        for (int N : {96, 192}) {
            const int nacc = 256 / N >= 2 ? 2 : 1;
            const int pitch = layout == 6 ? 32 : 64;
            Cfg c{};
            c.layout = layout; c.a_start = 0; c.N = N; c.reps = 4000; c.kind_f16 = 0; c.nacc = nacc; c.a_step = 128 * pitch;
            c.a_lbo = 16; c.a_sbo = 8 * pitch; c.b_lbo = 128; c.b_sbo = 256;
            probe<<<1, 128, 32768 + BBYTES + 1024>>>(da, ABYTES, db, BBYTES, c, dd, dc);
            CK(cudaDeviceSynchronize());
            long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
            printf("timing i8 swizzle-%dB overlapped-A M=128 N=%3d, %d accumulators: %.1f cycles per MMA\n", pitch, N, nacc, (double)cyc / c.reps);
        }
    printf("probe done\n");
    return 0;
}
