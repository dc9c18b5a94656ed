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
// Exactness: the unpacked samples (b-128) are integers in [-128,127]: exact in bf16.  Every f32 tap (pre-scaled by
// the exact factor 1/128) is split into three bf16 terms hi+mid+lo = c exactly (3 x 8 significant bits), so each
// partial product is exact in the f32 accumulator and the only roundings are the accumulations themselves -- the
// same count as a CPU f32 sum.  I and Q are kept as two bf16 planes in shared memory; the A fragments of a
// row block come from one `ldmatrix.x4` whose 32 row addresses are 16-byte-aligned offsets into the flat plane
// (P*D*2 bytes apart), which is what makes the Toeplitz operand free.  Accumulator fragments map to 64 consecutive
// complex outputs per 8 rows, so the epilogue is a fully coalesced 16-byte store per lane, no staging.
//
// (This is the legacy mma.sync tensor path.  ncu shows it already moves the FIR from FP32-bound to HBM/LSU-bound
// for K = 64; the tcgen05/TMEM version of the same Toeplitz mapping is the planned successor for K = 255.)
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "kernels.h"

namespace sdr {

namespace {

constexpr int TCF_WARPS = 8;
constexpr int TCF_P = 8;  // outputs per window row

__device__ __forceinline__ void ldmatrix_x4(unsigned &r0, unsigned &r1, unsigned &r2, unsigned &r3, unsigned saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// (b - 128) as bf16 pairs: 0x4B000000|b is the float 2^23 + b; subtracting 2^23 + 128 is exact
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned *>(&v);
}
__device__ __forceinline__ float centred_byte(unsigned word, int idx) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (unsigned)idx)) - 8388736.0f;
}

struct TcArgs {
    FirArgs f;
    const uint2 *btab;  // [KS][NT][3][32] fragment-ordered Toeplitz taps for this call's delta
    int KS;             // k-steps of 16 window positions
    int a0_mod;         // delta: window start minus the 8-aligned staging start
    long long in0;      // not used
};

// RB row blocks (of 16 windows x 8 phases) per warp; TC: complex taps
template <int RB, bool TC>
__global__ void __launch_bounds__(TCF_WARPS * 32, 2) fir_mma_kernel(TcArgs a) {
    extern __shared__ uint4 smem16[];
    const FirArgs &f = a.f;
    const int D = f.D, K = f.K, KS = a.KS;
    constexpr int NT = TC ? 2 : 1;                       // tap tables: real | (re, im)
    const int tile_out = TCF_WARPS * RB * 16 * TCF_P;    // kept outputs per CTA
    const int rowstep = TCF_P * D;                        // elements between window rows
    const int span = (TCF_WARPS * RB * 16 - 1) * rowstep + KS * 16;  // plane elements the CTA touches
    const int plane = (span + 8 + 7) & ~7;                // padded plane length (elements)
    __nv_bfloat16 *pI = reinterpret_cast<__nv_bfloat16 *>(smem16);
    __nv_bfloat16 *pQ = pI + plane;
    uint2 *bt = reinterpret_cast<uint2 *>(pQ + plane);    // KS*NT*3*32 uint2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = blockIdx.y;
    const long long m0 = (long long)blockIdx.x * tile_out;                  // first kept output of the CTA
    // window start of row 0 (input coordinates) and the 8-aligned staging start below it
    const long long w0 = f.first + m0 * D - (K - 1) - a.a0_mod;              // multiple of 8 by construction
    const unsigned char *in = (const unsigned char *)f.in + (long long)ch * f.in_stride * 2;
    const unsigned char *hist = (const unsigned char *)f.hist + (long long)ch * f.hist_stride * 2;

    for (int i = tid; i < KS * NT * 3 * 32; i += TCF_WARPS * 32) bt[i] = __ldg(a.btab + i);

    // ---- stage: u8 IQ -> two bf16 planes of (b - 128) ----
    for (int c = tid; c < plane / 8; c += TCF_WARPS * 32) {
        const long long s0 = w0 + 8LL * c;
        uint4 q;
        if (s0 >= 0 && s0 + 8 <= f.n_in) {
            q = __ldg(reinterpret_cast<const uint4 *>(in + 2 * s0));
        } else if (s0 < 0 && s0 + 8 <= 0 && s0 >= -(long long)f.HL) {
            q = __ldg(reinterpret_cast<const uint4 *>(hist + 2 * ((long long)f.HL + s0)));
        } else {
            unsigned short h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long s = s0 + i;
                unsigned short v = 0x8080;  // the zero sample
                if (s >= 0) { if (s < f.n_in) v = *reinterpret_cast<const unsigned short *>(in + 2 * s); }
                else if (s >= -(long long)f.HL) v = *reinterpret_cast<const unsigned short *>(hist + 2 * ((long long)f.HL + s));
                h[i] = v;
            }
            q.x = h[0] | ((unsigned)h[1] << 16); q.y = h[2] | ((unsigned)h[3] << 16);
            q.z = h[4] | ((unsigned)h[5] << 16); q.w = h[6] | ((unsigned)h[7] << 16);
        }
        uint4 vi, vq;
        vi.x = pack_bf16x2(centred_byte(q.x, 0), centred_byte(q.x, 2)); vq.x = pack_bf16x2(centred_byte(q.x, 1), centred_byte(q.x, 3));
        vi.y = pack_bf16x2(centred_byte(q.y, 0), centred_byte(q.y, 2)); vq.y = pack_bf16x2(centred_byte(q.y, 1), centred_byte(q.y, 3));
        vi.z = pack_bf16x2(centred_byte(q.z, 0), centred_byte(q.z, 2)); vq.z = pack_bf16x2(centred_byte(q.z, 1), centred_byte(q.z, 3));
        vi.w = pack_bf16x2(centred_byte(q.w, 0), centred_byte(q.w, 2)); vq.w = pack_bf16x2(centred_byte(q.w, 1), centred_byte(q.w, 3));
        reinterpret_cast<uint4 *>(pI)[c] = vi;
        reinterpret_cast<uint4 *>(pQ)[c] = vq;
    }
    __syncthreads();

    // ---- MMA: this warp owns RB row blocks ----
    float accI[RB][4], accQ[RB][4];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) { accI[r][i] = 0.0f; accQ[r][i] = 0.0f; }
    // ldmatrix row address of this lane: matrix = lane/8; row = lane%8 + 8*(matrix&1); col0 = 8*(matrix>>1)
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lcol = 8 * (lane >> 4);
    const unsigned baseI = (unsigned)__cvta_generic_to_shared(pI), baseQ = (unsigned)__cvta_generic_to_shared(pQ);
    const int row0 = warp * RB * 16;
    for (int kk = 0; kk < KS; ++kk) {
        uint2 b[NT][3];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int s = 0; s < 3; ++s) b[t][s] = bt[((kk * NT + t) * 3 + s) * 32 + lane];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const unsigned off = 2u * (unsigned)((row0 + r * 16 + lrow) * rowstep + kk * 16 + lcol);
            unsigned i0, i1, i2, i3, q0, q1, q2, q3;
            ldmatrix_x4(i0, i1, i2, i3, baseI + off);
            ldmatrix_x4(q0, q1, q2, q3, baseQ + off);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                mma_bf16(accI[r], i0, i1, i2, i3, b[0][s].x, b[0][s].y);   // yI += xI * cr
                mma_bf16(accQ[r], q0, q1, q2, q3, b[0][s].x, b[0][s].y);   // yQ += xQ * cr
                if (TC) {
                    // (v*c).re = vr*cr - vi*ci ; (v*c).im = vr*ci + vi*cr : table 1 holds ci, sign applied to the sample side
                    mma_bf16(accI[r], q0 ^ 0x80008000u, q1 ^ 0x80008000u, q2 ^ 0x80008000u, q3 ^ 0x80008000u,
                             b[NT - 1][s].x, b[NT - 1][s].y);                // yI -= xQ * ci
                    mma_bf16(accQ[r], i0, i1, i2, i3, b[NT - 1][s].x, b[NT - 1][s].y);  // yQ += xI * ci
                }
            }
        }
    }

    // ---- epilogue: lane (g,t) holds I/Q of outputs m = P*(row+g) + 2t, +1 and the same 8 rows further ----
    float2 *out = (float2 *)f.out + (long long)ch * f.out_stride;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long m = m0 + (long long)(row0 + r * 16 + g + 8 * h) * TCF_P + 2 * t;
            const float4 v = make_float4(accI[r][2 * h], accQ[r][2 * h], accI[r][2 * h + 1], accQ[r][2 * h + 1]);
            if (m + 1 < f.n_out) *reinterpret_cast<float4 *>(out + m) = v;
            else if (m < f.n_out) out[m] = make_float2(v.x, v.y);
        }
    }
}

inline unsigned short bf16_bits_rn(float v) {
    unsigned u;
    std::memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (unsigned short)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
inline float bf16_to_float(unsigned short b) {
    unsigned u = (unsigned)b << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

int fir_tc_ksteps(int K, int D) { return ((TCF_P - 1) * D + K + 7 + 15) / 16; }

size_t fir_tc_table_words(int K, int D, bool tc) { return (size_t)fir_tc_ksteps(K, D) * (tc ? 2 : 1) * 3 * 32; }

// host: fragment-ordered Toeplitz tables for the 8 possible alignments delta.  Layout [delta][kk][table][split][lane].
// Fragment of mma.m16n8k16 B (16x8, "col"): lane (g = lane/4, t = lane%4) holds b0 = {T[2t][g], T[2t+1][g]},
// b1 = {T[2t+8][g], T[2t+9][g]}.
void fir_tc_build_tables(const float *taps, int K, bool tc, int D, std::vector<uint2> &out) {
    const int KS = fir_tc_ksteps(K, D), NT = tc ? 2 : 1;
    out.assign((size_t)8 * KS * NT * 3 * 32, make_uint2(0, 0));
    auto split3 = [](float c, unsigned short s[3]) {
        const float cs = c * 0.0078125f;  // exact: the 1/128 of the unpack lives in the taps
        float r = cs;
        for (int i = 0; i < 3; ++i) {
            s[i] = bf16_bits_rn(r);
            r -= bf16_to_float(s[i]);
        }
    };
    for (int delta = 0; delta < 8; ++delta)
        for (int kk = 0; kk < KS; ++kk)
            for (int tb = 0; tb < NT; ++tb)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3;
                    unsigned short v[4][3];
                    const int srow[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
                    for (int e = 0; e < 4; ++e) {
                        const int s = kk * 16 + srow[e];
                        const int k = g * D + (K - 1) + delta - s;
                        float c = 0.0f;
                        if (k >= 0 && k < K) c = tc ? taps[2 * k + tb] : taps[k];
                        split3(c, v[e]);
                    }
                    for (int sp = 0; sp < 3; ++sp) {
                        uint2 w;
                        w.x = (unsigned)v[0][sp] | ((unsigned)v[1][sp] << 16);
                        w.y = (unsigned)v[2][sp] | ((unsigned)v[3][sp] << 16);
                        out[((((size_t)delta * KS + kk) * NT + tb) * 3 + sp) * 32 + lane] = w;
                    }
                }
}

// returns SDR_ERR_UNSUPPORTED when the tensor path does not apply (caller falls back to fir_launch)
int fir_tc_launch(const FirArgs &f, bool tc, const uint2 *d_tables, cudaStream_t st) {
    if (f.n_out <= 0) return SDR_OK;
    if (((uintptr_t)f.in & 15) || ((uintptr_t)f.hist & 15) || ((uintptr_t)f.out & 15) || (f.out_stride & 1) ||
        (f.in_stride & 7) || (f.hist_stride & 7) || f.n_ch > 65535)
        return SDR_ERR_UNSUPPORTED;
    const int KS = fir_tc_ksteps(f.K, f.D), NT = tc ? 2 : 1;
    TcArgs a;
    a.f = f;
    a.KS = KS;
    a.in0 = 0;
    // delta = (first - (K-1)) mod 8, so that w0 is a multiple of 8 for every tile (tile_out*D is a multiple of 8)
    long long d = (f.first - (f.K - 1)) % 8;
    if (d < 0) d += 8;
    a.a0_mod = (int)d;
    a.btab = d_tables + (size_t)d * KS * NT * 3 * 32;
    auto launch = [&](auto kern, int RB) -> int {
        const int tile_out = TCF_WARPS * RB * 16 * TCF_P;
        const long long span = (long long)(TCF_WARPS * RB * 16 - 1) * TCF_P * f.D + KS * 16;
        const long long plane = (span + 8 + 7) & ~7LL;
        const size_t smem = (size_t)plane * 2 * 2 + (size_t)KS * NT * 3 * 32 * sizeof(uint2);
        if (smem > 200 * 1024) return SDR_ERR_UNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        dim3 grid((unsigned)((f.n_out + tile_out - 1) / tile_out), (unsigned)f.n_ch);
        kern<<<grid, TCF_WARPS * 32, smem, st>>>(a);
        count_launch();
        return launch_status();
    };
    // bigger row-block counts amortise the tap-fragment loads; the shared window grows with D
    const long long per_rb = (long long)TCF_WARPS * 16 * TCF_P * f.D * 4;  // plane bytes (I+Q) per RB step
    int RB = 4;
    while (RB > 1 && per_rb * RB + f.K * 4 > 96 * 1024) RB >>= 1;
    if (tc) {
        if (RB == 4) return launch(fir_mma_kernel<4, true>, 4);
        if (RB == 2) return launch(fir_mma_kernel<2, true>, 2);
        return launch(fir_mma_kernel<1, true>, 1);
    }
    if (RB == 4) return launch(fir_mma_kernel<4, false>, 4);
    if (RB == 2) return launch(fir_mma_kernel<2, false>, 2);
    return launch(fir_mma_kernel<1, false>, 1);
}

}  // namespace sdr
