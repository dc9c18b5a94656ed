// fir_tc.cu -- K5: tensor-core FIR (+ Decimate) for u8 IQ input, recast as a Toeplitz GEMM.
//
// Same reference functions as fir.cu (RtlTcpSignal::next -> Fir::apply -> Decimate::next:
// src/rtltcp.rs:158-164, src/filter/fir.rs:23-32, src/signal/adapters/mod.rs:30-37), for the cases where
// the CUDA-core kernel is FP32-pipe-bound (K = 64: 128 FMA per 10-byte sample is 44 % of the HBM roofline
// at best).
//
//   kept output m = P*i + j  (row i, phase j < P = 8):   y_m = sum_k c[k] x[first + m*D - k]
//   A[i][s] = x[w0 + P*D*i + s]      rows are overlapping windows of ONE flat sample array (no im2col copy)
//   T[s][j] = c[j*D + K-1 + delta - s]   banded Toeplitz block, constant, built once on the host
//   Y = A * T                          one m16n8k16 MMA covers 16 windows x 8 phases x 16 window positions
//
// Exactness: the unpacked samples (b-128) are integers in [-128,127]: exact in fp16.  Every f32 tap is scaled by a
// power of two 2^S (so the largest sits just below 2^14; 2^-S-7 is applied, exactly, in the epilogue) and split into
// two fp16 terms hi+lo, 22 significant bits (2^-22 relative per tap: below the 2^-24..2^-20 the f32 accumulation of
// K terms carries anyway).  Each partial product (8 bits x 11 bits) is exact in the f32 accumulator, so the only
// roundings are the accumulations.  I and Q are kept as two fp16 planes in shared memory; the A fragments of a
// row block come from one `ldmatrix.x4` whose 32 row addresses are 16-byte-aligned offsets into the flat plane
// (P*D*2 bytes apart), which is what makes the Toeplitz operand free.  Accumulator fragments map to 64 consecutive
// complex outputs per 8 rows, so the epilogue is a fully coalesced 16-byte store per lane, no staging.
//
// (This is the legacy mma.sync tensor path.  ncu shows it already moves the FIR from FP32-bound to HBM/LSU-bound
// for K = 64; the tcgen05/TMEM version of the same Toeplitz mapping is the planned successor for K = 255.)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include "kernels.h"

namespace sdr {

namespace {

// outputs per window row P = 8 * NTL: with NTL = 2 two 8-wide MMA n-tiles share every A fragment (fewer ldmatrix
// wavefronts per MMA, slightly longer windows) -- used for real taps; complex taps already reuse each fragment 4x
constexpr int TCF_SPLITS = 2;

__device__ __forceinline__ void ldmatrix_x4(unsigned &r0, unsigned &r1, unsigned &r2, unsigned &r3, unsigned saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(saddr));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                        unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// two bytes -> half2 of (b - 128): 0x6400 | b is the fp16 number 1024 + b; subtracting 1152 is exact.
// selector picks bytes (i, j) of `word` into the low bytes of the two halves, 0x64 into the high bytes.
__device__ __forceinline__ unsigned centred_half2(unsigned word, unsigned selector) {
    const unsigned m = __byte_perm(word, 0x64646464u, selector);
    __half2 h = *reinterpret_cast<const __half2 *>(&m);
    h = __hsub2(h, __float2half2_rn(1152.0f));
    return *reinterpret_cast<unsigned *>(&h);
}
constexpr unsigned SEL_I = 0x4240u, SEL_Q = 0x4341u;  // word = I0 Q0 I1 Q1

struct TcArgs {
    FirArgs f;
    const uint2 *btab;  // [KS][NT][2 n-tiles][SPLITS][32] fragment-ordered Toeplitz taps for this call's delta
    float out_scale;    // 2^-(S+7), exact
    int KS;             // k-steps of 16 window positions
    int a0_mod;         // delta: window start minus the 8-aligned staging start
    int ntiles;         // tiles per channel
};

// 16-byte chunk swizzle of the fp16 planes: rows of an ldmatrix 8x8 block are 2*D chunks apart (P = 16), which for
// D = 1 would put rows r and r+4 on the same bank group; XOR-ing bit 3 into bit 0 separates them.
__device__ __forceinline__ int swz(int chunk) { return chunk ^ ((chunk >> 3) & 1); }

// Persistent, warp-private pipeline: every warp owns tiles of RB*16 window rows end to end -- it loads the raw
// bytes of its rows (plus the tap-length halo) with batched 16-byte loads, converts them into its private pair of
// fp16 planes, multiplies and stores -- so there is no CTA-wide barrier at all (only __syncwarp) and the warps of
// an SM drift apart: while some convert (ALU) others multiply (tensor) or wait on HBM.  The halo that neighbouring
// warps both stage (KS*16 of every RB*16*P*D samples) is an L2 hit, not HBM traffic.
template <int NW, int RB, bool TC>
__global__ void __launch_bounds__(NW * 32, (NW * 32 <= 128) ? 4 : 2) fir_mma_kernel(TcArgs a) {
    extern __shared__ uint4 smem16[];
    const FirArgs &f = a.f;
    const int D = f.D, K = f.K, KS = a.KS;
    constexpr int NT = TC ? 2 : 1;        // tap tables: real | (re, im)
    constexpr int NTL = TC ? 1 : 2;       // 8-wide n-tiles per window row
    constexpr int P = 8 * NTL;            // outputs per window row
    constexpr int ROWS = RB * 16;         // window rows per warp tile
    const int tile_out = ROWS * P;        // kept outputs per warp tile
    const int rowchunks = NTL * D;        // 16-byte chunks between window rows
    const int nchunks = (ROWS - 1) * rowchunks + 2 * KS + 1;  // chunks a warp tile touches (+1 for the swizzle)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint2 *bt = reinterpret_cast<uint2 *>(smem16);
    const int bt_words = KS * NT * NTL * TCF_SPLITS * 32;
    uint4 *pI = smem16 + (bt_words + 1) / 2 + warp * 2 * (nchunks + 1);
    uint4 *pQ = pI + nchunks + 1;
    for (int i = tid; i < bt_words; i += NW * 32) bt[i] = __ldg(a.btab + i);
    __syncthreads();

    // ldmatrix row address of this lane: matrix = lane/8; row = lane%8 + 8*(matrix&1); col0 = 8*(matrix>>1)
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lhalf = lane >> 4;
    const unsigned baseI = (unsigned)__cvta_generic_to_shared(pI), baseQ = (unsigned)__cvta_generic_to_shared(pQ);
    const int g = lane >> 2, t = lane & 3;
    const float sc = a.out_scale;
    const long long nwork = (long long)a.ntiles * f.n_ch;  // a.ntiles = warp tiles per channel

    // raw 8-sample chunk c of the tile whose window starts at input index w0 (history / zeros outside the block)
    auto load_chunk = [&](long long w0, int c, const unsigned char *in, const unsigned char *hist) -> uint4 {
        uint4 q = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);  // zero samples
        const long long s0 = w0 + 8LL * c;
        if (c < nchunks) {
            if (s0 >= 0 && s0 + 8 <= f.n_in) {
                q = __ldg(reinterpret_cast<const uint4 *>(in + 2 * s0));
            } else if (s0 + 8 <= 0 && s0 >= -(long long)f.HL) {
                q = __ldg(reinterpret_cast<const uint4 *>(hist + 2 * ((long long)f.HL + s0)));
            } else if (s0 < f.n_in && s0 + 8 > -(long long)f.HL) {
                unsigned short h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long s = s0 + i;
                    unsigned short v = 0x8080;
                    if (s >= 0) { if (s < f.n_in) v = *reinterpret_cast<const unsigned short *>(in + 2 * s); }
                    else if (s >= -(long long)f.HL) v = *reinterpret_cast<const unsigned short *>(hist + 2 * ((long long)f.HL + s));
                    h[i] = v;
                }
                q.x = h[0] | ((unsigned)h[1] << 16); q.y = h[2] | ((unsigned)h[3] << 16);
                q.z = h[4] | ((unsigned)h[5] << 16); q.w = h[6] | ((unsigned)h[7] << 16);
            }
        }
        return q;
    };
    auto tile_geom = [&](long long w, long long &m0, long long &w0, const unsigned char *&in, const unsigned char *&hist) {
        const int ch = (int)(w / a.ntiles);
        m0 = (w % a.ntiles) * tile_out;
        w0 = f.first + m0 * D - (K - 1) - a.a0_mod;  // multiple of 8 by construction
        in = (const unsigned char *)f.in + (long long)ch * f.in_stride * 2;
        hist = (const unsigned char *)f.hist + (long long)ch * f.hist_stride * 2;
    };
    auto convert_store = [&](const uint4 &q, int c) {
        if (c < nchunks) {
            uint4 vi, vq;
            vi.x = centred_half2(q.x, SEL_I); vq.x = centred_half2(q.x, SEL_Q);
            vi.y = centred_half2(q.y, SEL_I); vq.y = centred_half2(q.y, SEL_Q);
            vi.z = centred_half2(q.z, SEL_I); vq.z = centred_half2(q.z, SEL_Q);
            vi.w = centred_half2(q.w, SEL_I); vq.w = centred_half2(q.w, SEL_Q);
            const int pc = swz(c);
            pI[pc] = vi;
            pQ[pc] = vq;
        }
    };

    const long long wstride = (long long)gridDim.x * NW;
    long long w = (long long)blockIdx.x * NW + warp;
    // software pipeline: the first 128 chunks (4 per lane) of the NEXT tile are loaded into registers before the
    // MMAs of the current tile, so their HBM latency is hidden behind tensor work
    uint4 pre[4];
    if (w < nwork) {
        long long m0, w0; const unsigned char *in, *hist;
        tile_geom(w, m0, w0, in, hist);
#pragma unroll
        for (int u = 0; u < 4; ++u) pre[u] = load_chunk(w0, u * 32 + lane, in, hist);
    }
    for (; w < nwork; w += wstride) {
        long long m0, w0; const unsigned char *in, *hist;
        tile_geom(w, m0, w0, in, hist);
        const int ch = (int)(w / a.ntiles);

        // ---- stage: u8 IQ -> two fp16 planes of (b - 128) ----
#pragma unroll
        for (int u = 0; u < 4; ++u) convert_store(pre[u], u * 32 + lane);
        for (int c0 = 128; c0 < nchunks; c0 += 128) {  // windows longer than 1024 samples: the rest, 4 loads in flight
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = load_chunk(w0, c0 + u * 32 + lane, in, hist);
#pragma unroll
            for (int u = 0; u < 4; ++u) convert_store(q[u], c0 + u * 32 + lane);
        }
        __syncwarp();
        if (w + wstride < nwork) {
            long long m1, w1; const unsigned char *in1, *hist1;
            tile_geom(w + wstride, m1, w1, in1, hist1);
#pragma unroll
            for (int u = 0; u < 4; ++u) pre[u] = load_chunk(w1, u * 32 + lane, in1, hist1);
        }

        // ---- MMA ----
        float accI[RB][NTL][4], accQ[RB][NTL][4];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int n = 0; n < NTL; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) { accI[r][n][i] = 0.0f; accQ[r][n][i] = 0.0f; }
        for (int kk = 0; kk < KS; ++kk) {
            uint2 b[NT][NTL][TCF_SPLITS];
#pragma unroll
            for (int tb = 0; tb < NT; ++tb)
#pragma unroll
                for (int n = 0; n < NTL; ++n)
#pragma unroll
                    for (int s = 0; s < TCF_SPLITS; ++s)
                        b[tb][n][s] = bt[(((kk * NT + tb) * NTL + n) * TCF_SPLITS + s) * 32 + lane];
            unsigned fi[RB][4], fq[RB][4];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int chunk = (r * 16 + lrow) * rowchunks + 2 * kk + lhalf;
                const unsigned off = 16u * (unsigned)swz(chunk);
                ldmatrix_x4(fi[r][0], fi[r][1], fi[r][2], fi[r][3], baseI + off);
                ldmatrix_x4(fq[r][0], fq[r][1], fq[r][2], fq[r][3], baseQ + off);
            }
            // dependent MMAs (same accumulator) are issued RB*NTL*2 apart so the tensor pipe never waits on itself
#pragma unroll
            for (int s = 0; s < TCF_SPLITS; ++s) {
#pragma unroll
                for (int r = 0; r < RB; ++r)
#pragma unroll
                    for (int n = 0; n < NTL; ++n) {
                        mma_f16(accI[r][n], fi[r][0], fi[r][1], fi[r][2], fi[r][3], b[0][n][s].x, b[0][n][s].y);  // yI += xI*cr
                        mma_f16(accQ[r][n], fq[r][0], fq[r][1], fq[r][2], fq[r][3], b[0][n][s].x, b[0][n][s].y);  // yQ += xQ*cr
                    }
                if (TC) {
                    // (v*c).re = vr*cr - vi*ci ; (v*c).im = vr*ci + vi*cr ; the sign goes on the sample side
#pragma unroll
                    for (int r = 0; r < RB; ++r)
#pragma unroll
                        for (int n = 0; n < NTL; ++n) {
                            mma_f16(accI[r][n], fq[r][0] ^ 0x80008000u, fq[r][1] ^ 0x80008000u, fq[r][2] ^ 0x80008000u,
                                    fq[r][3] ^ 0x80008000u, b[NT - 1][n][s].x, b[NT - 1][n][s].y);              // yI -= xQ*ci
                            mma_f16(accQ[r][n], fi[r][0], fi[r][1], fi[r][2], fi[r][3], b[NT - 1][n][s].x,
                                    b[NT - 1][n][s].y);                                                          // yQ += xI*ci
                        }
                }
            }
        }

        // ---- epilogue: lane (g,t) holds I/Q of outputs m = P*(row + g + 8h) + 8n + 2t, +1 ----
        float2 *out = (float2 *)f.out + (long long)ch * f.out_stride;
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int n = 0; n < NTL; ++n) {
                    const long long m = m0 + (long long)(r * 16 + g + 8 * h) * P + 8 * n + 2 * t;
                    const float4 v = make_float4(accI[r][n][2 * h] * sc, accQ[r][n][2 * h] * sc,
                                                 accI[r][n][2 * h + 1] * sc, accQ[r][n][2 * h + 1] * sc);
                    if (m + 1 < f.n_out) *reinterpret_cast<float4 *>(out + m) = v;
                    else if (m < f.n_out) out[m] = make_float2(v.x, v.y);
                }
        __syncwarp();  // the planes are rewritten by the next tile
    }
}

inline unsigned short half_bits_rn(float v) {
    return __half_as_ushort(__float2half_rn(v));
}
inline float half_to_float(unsigned short b) {
    return __half2float(__ushort_as_half(b));
}

}  // namespace

static int tc_p(bool tc) { return tc ? 8 : 16; }
int fir_tc_ksteps(int K, int D, bool tc) { return ((tc_p(tc) - 1) * D + K + 7 + 15) / 16; }



// host: fragment-ordered Toeplitz tables for the 8 possible alignments delta.  Layout [delta][kk][table][split][lane].
// Fragment of mma.m16n8k16 B (16x8, "col"): lane (g = lane/4, t = lane%4) holds b0 = {T[2t][g], T[2t+1][g]},
// b1 = {T[2t+8][g], T[2t+9][g]}.
float fir_tc_build_tables(const float *taps, int K, bool tc, int D, std::vector<uint2> &out) {
    const int KS = fir_tc_ksteps(K, D, tc), NT = tc ? 2 : 1, NTL = tc_p(tc) / 8;
    out.assign((size_t)8 * KS * NT * NTL * TCF_SPLITS * 32, make_uint2(0, 0));
    // power-of-two scale: the largest |tap| lands in [2^13, 2^14), far from fp16 overflow (65504) and with the
    // lo terms of even very small taps still above the fp16 subnormal step
    float cmax = 0.0f;
    for (int i = 0; i < K * (tc ? 2 : 1); ++i) cmax = std::fmax(cmax, std::fabs(taps[i]));
    int e = 0;
    if (cmax > 0.0f && std::isfinite(cmax)) {
        std::frexp(cmax, &e);  // cmax = m * 2^e, m in [0.5, 1)
    }
    const int S = 14 - e;   // cmax * 2^S in [2^13, 2^14)
    const float up = std::ldexp(1.0f, S);
    auto split = [up](float c, unsigned short s[TCF_SPLITS]) {
        float r = c * up;  // exact (power of two)
        for (int i = 0; i < TCF_SPLITS; ++i) {
            s[i] = half_bits_rn(r);
            r -= half_to_float(s[i]);
        }
    };
    for (int delta = 0; delta < 8; ++delta)
        for (int kk = 0; kk < KS; ++kk)
            for (int tb = 0; tb < NT; ++tb)
              for (int nt = 0; nt < NTL; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = (lane >> 2) + 8 * nt, t = lane & 3;
                    unsigned short v[4][TCF_SPLITS];
                    const int srow[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
                    for (int el = 0; el < 4; ++el) {
                        const int s = kk * 16 + srow[el];
                        const int k = g * D + (K - 1) + delta - s;
                        float c = 0.0f;
                        if (k >= 0 && k < K) c = tc ? taps[2 * k + tb] : taps[k];
                        split(c, v[el]);
                    }
                    for (int sp = 0; sp < TCF_SPLITS; ++sp) {
                        uint2 w;
                        w.x = (unsigned)v[0][sp] | ((unsigned)v[1][sp] << 16);
                        w.y = (unsigned)v[2][sp] | ((unsigned)v[3][sp] << 16);
                        out[(((((size_t)delta * KS + kk) * NT + tb) * NTL + nt) * TCF_SPLITS + sp) * 32 + lane] = w;
                    }
                }
    return std::ldexp(1.0f, -S - 7);  // epilogue scale: undo 2^S and apply the unpack's 1/128
}

// returns SDR_ERR_UNSUPPORTED when the tensor path does not apply (caller falls back to fir_launch)
int fir_tc_launch(const FirArgs &f, bool tc, const uint2 *d_tables, float out_scale, cudaStream_t st) {
    if (f.n_out <= 0) return SDR_OK;
    if (((uintptr_t)f.in & 15) || ((uintptr_t)f.hist & 15) || ((uintptr_t)f.out & 15) || (f.out_stride & 1) ||
        (f.in_stride & 7) || (f.hist_stride & 7))
        return SDR_ERR_UNSUPPORTED;
    const int KS = fir_tc_ksteps(f.K, f.D, tc), NT = tc ? 2 : 1, TCF_P = tc_p(tc), NTL = TCF_P / 8;
    TcArgs a;
    a.f = f;
    a.KS = KS;
    a.out_scale = out_scale;
    // delta = (first - (K-1)) mod 8, so that w0 is a multiple of 8 for every tile (tile_out*D is a multiple of 8)
    long long d = (f.first - (f.K - 1)) % 8;
    if (d < 0) d += 8;
    a.a0_mod = (int)d;
    const size_t tab_words = (size_t)KS * NT * NTL * TCF_SPLITS * 32;
    a.btab = d_tables + (size_t)d * tab_words;
    const int sms = current_sm_count();
    auto smem_of = [&](int NW, int RB) -> size_t {
        const long long nchunks = (long long)(RB * 16 - 1) * NTL * f.D + 2 * KS + 1;
        return ((tab_words + 1) / 2 + (size_t)NW * 2 * (nchunks + 1)) * 16;
    };
    auto launch = [&](auto kern, int NW, int RB) -> int {
        const int tile_out = RB * 16 * TCF_P;
        const size_t smem = smem_of(NW, RB);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        a.ntiles = (int)((f.n_out + tile_out - 1) / tile_out);
        const long long nwork = (long long)a.ntiles * f.n_ch;
        int per_sm = (int)std::min<size_t>(NW == 4 ? 4 : 2, (220 * 1024) / smem);
        if (per_sm < 1) per_sm = 1;
        const unsigned grid = (unsigned)std::min<long long>((nwork + NW - 1) / NW, (long long)sms * per_sm);
        kern<<<grid, NW * 32, smem, st>>>(a);
        count_launch();
        return launch_status();
    };
    // 4 warps per CTA, up to 4 CTAs per SM; RB = 2 amortises the tap-fragment loads over two row blocks.  The
    // private planes grow with the decimation D, so large D falls back to one row block per warp / fewer CTAs.
    const int cfgs[3][2] = {{4, 2}, {4, 1}, {2, 1}};
    int pick = -1;
    for (int i = 0; i < 3 && pick < 0; ++i)
        if (smem_of(cfgs[i][0], cfgs[i][1]) <= 55 * 1024) pick = i;
    for (int i = 0; i < 3 && pick < 0; ++i)
        if (smem_of(cfgs[i][0], cfgs[i][1]) <= 110 * 1024) pick = i;
    for (int i = 2; i >= 0 && pick < 0; --i)
        if (smem_of(cfgs[i][0], cfgs[i][1]) <= 200 * 1024) pick = i;
    if (pick < 0) return SDR_ERR_UNSUPPORTED;
    if (tc) {
        switch (pick) {
            case 0: return launch(fir_mma_kernel<4, 2, true>, 4, 2);
            case 1: return launch(fir_mma_kernel<4, 1, true>, 4, 1);
            default: return launch(fir_mma_kernel<2, 1, true>, 2, 1);
        }
    }
    switch (pick) {
        case 0: return launch(fir_mma_kernel<4, 2, false>, 4, 2);
        case 1: return launch(fir_mma_kernel<4, 1, false>, 4, 1);
        default: return launch(fir_mma_kernel<2, 1, false>, 2, 1);
    }
}

}  // namespace sdr
