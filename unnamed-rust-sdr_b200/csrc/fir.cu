// fir.cu -- K1/K2: fused u8-IQ unpack + FIR (+ Decimate) on CUDA cores.
//
// Replaces  RtlTcpSignal::next (src/rtltcp.rs:158-164) -> signal::Filter::next
// (src/signal/adapters/mod.rs:94-96) -> Fir::apply (src/filter/fir.rs:23-32) -> Decimate::next
// (src/signal/adapters/mod.rs:30-37) of the reference, computing only the kept outputs.
//
//   fir_rb_kernel      D == 1.  A CTA stages a window of TILE + Kp samples (tap-length halo on the
//                      left, taken from the carried history for the first tile) in shared memory,
//                      unpacking u8 IQ on the way in with 16-byte loads.  Each thread owns 8
//                      consecutive outputs and slides a 16-sample register window over the taps:
//                      one 64-byte shared read feeds 128 FMAs.  The shared layout pads every
//                      8-sample chunk to 80 bytes so the per-thread 128-bit reads are conflict free.
//                      Outputs go back through shared memory so global stores are 512 B per warp.
//   fir_generic_kernel any D / format / alignment: one output per thread, operands from L1/L2.
//
// STRICT variants reproduce the reference's arithmetic exactly: f32 multiply then add, taps in
// ascending k, no FMA contraction (src/filter/convolve.rs:13-15) -> bit-identical outputs.
#include "kernels.h"

namespace sdr {

namespace {

constexpr int RB_THREADS = 256;
constexpr int RB_O = 8;
constexpr int RB_TILE = RB_THREADS * RB_O;  // 2048 outputs per CTA
constexpr int RB_MAX_KP = 4096;

template <int FMT> struct Elem;
template <> struct Elem<SDR_FMT_U8IQ> { static constexpr int bytes = 2; };
template <> struct Elem<SDR_FMT_C64> { static constexpr int bytes = 8; };
template <> struct Elem<SDR_FMT_F32> { static constexpr int bytes = 4; };

template <int FMT>
__device__ __forceinline__ float2 load_one(const char *base, long long idx) {
    if (FMT == SDR_FMT_U8IQ) {
        return unpack_iq_u16(*reinterpret_cast<const uint16_t *>(base + idx * 2));
    } else if (FMT == SDR_FMT_C64) {
        return *reinterpret_cast<const float2 *>(base + idx * 8);
    } else {
        return make_float2(*reinterpret_cast<const float *>(base + idx * 4), 0.0f);
    }
}

// sample at stream-relative index idx (may be negative: carried history, then zeros)
template <int FMT>
__device__ __forceinline__ float2 load_sample(const char *in, const char *hist, long long idx, int HL) {
    if (idx >= 0) return load_one<FMT>(in, idx);
    if (idx >= -(long long)HL) return load_one<FMT>(hist, HL + idx);
    return make_float2(0.0f, 0.0f);
}

template <bool TC, bool STRICT>
__device__ __forceinline__ void mac(float2 &acc, const float2 x, const float cr, const float ci) {
    if (!TC) {
        if (STRICT) {
            acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, cr));
            acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, cr));
        } else {
            acc.x = fmaf(x.x, cr, acc.x);
            acc.y = fmaf(x.y, cr, acc.y);
        }
    } else {
        if (STRICT) {  // (v*c).re = v.re*c.re - v.im*c.im ; (v*c).im = v.re*c.im + v.im*c.re ; then +=
            const float pr = __fsub_rn(__fmul_rn(x.x, cr), __fmul_rn(x.y, ci));
            const float pi = __fadd_rn(__fmul_rn(x.x, ci), __fmul_rn(x.y, cr));
            acc.x = __fadd_rn(acc.x, pr);
            acc.y = __fadd_rn(acc.y, pi);
        } else {
            acc.x = fmaf(x.x, cr, acc.x);
            acc.x = fmaf(-x.y, ci, acc.x);
            acc.y = fmaf(x.x, ci, acc.y);
            acc.y = fmaf(x.y, cr, acc.y);
        }
    }
}

// ---------------------------------------------------------------------------------------
template <int FMT, bool TC, bool STRICT>
__global__ void fir_generic_kernel(FirArgs a) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (m >= a.n_out) return;
    constexpr int ES = Elem<FMT>::bytes;
    const char *in = (const char *)a.in + (long long)ch * a.in_stride * ES;
    const char *hist = (const char *)a.hist + (long long)ch * a.hist_stride * ES;
    const long long s = a.first + m * (long long)a.D;
    float2 acc = make_float2(0.0f, 0.0f);
    for (int k = 0; k < a.K; ++k) {
        const float2 x = load_sample<FMT>(in, hist, s - k, a.HL);
        const float cr = TC ? __ldg(a.taps + 2 * k) : __ldg(a.taps + k);
        const float ci = TC ? __ldg(a.taps + 2 * k + 1) : 0.0f;
        mac<TC, STRICT>(acc, x, cr, ci);
    }
    if (FMT == SDR_FMT_F32)
        ((float *)a.out)[(long long)ch * a.out_stride + m] = acc.x;
    else
        ((float2 *)a.out)[(long long)ch * a.out_stride + m] = acc;
}

// ---------------------------------------------------------------------------------------
// shared window layout: sample j of the window lives at float2 index 10*(j>>3) + (j&7)
__device__ __forceinline__ int win_idx(int j) { return 10 * (j >> 3) + (j & 7); }

template <int FMT>
__device__ __forceinline__ void load_chunk8(const char *p, float2 (&v)[8]) {  // 8 consecutive samples, 16B aligned
    if (FMT == SDR_FMT_U8IQ) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
        v[0] = unpack_iq16(q.x, 0); v[1] = unpack_iq16(q.x, 1);
        v[2] = unpack_iq16(q.y, 0); v[3] = unpack_iq16(q.y, 1);
        v[4] = unpack_iq16(q.z, 0); v[5] = unpack_iq16(q.z, 1);
        v[6] = unpack_iq16(q.w, 0); v[7] = unpack_iq16(q.w, 1);
    } else {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = __ldg(q + i);
            v[2 * i] = make_float2(f.x, f.y);
            v[2 * i + 1] = make_float2(f.z, f.w);
        }
    }
}

template <int FMT, bool TC, bool STRICT>
__global__ void __launch_bounds__(RB_THREADS, 2) fir_rb_kernel(FirArgs a) {
    extern __shared__ float4 smem4[];
    float2 *win = reinterpret_cast<float2 *>(smem4);
    constexpr int ES = Elem<FMT>::bytes;
    const int Kp = a.Kp;
    const int nchunks = (RB_TILE + Kp) >> 3;
    float *staps = reinterpret_cast<float *>(win + 10 * nchunks);
    const int tid = threadIdx.x;
    const int ch = blockIdx.y;
    const long long tile0 = (long long)blockIdx.x * RB_TILE;
    const char *in = (const char *)a.in + (long long)ch * a.in_stride * ES;
    const char *hist = (const char *)a.hist + (long long)ch * a.hist_stride * ES;

    for (int i = tid; i < Kp * (TC ? 2 : 1); i += RB_THREADS) staps[i] = __ldg(a.taps + i);

    // ---- stage the window: samples tile0 - Kp .. tile0 + TILE - 1 (HL == Kp) ----
    for (int c = tid; c < nchunks; c += RB_THREADS) {
        const long long s0 = tile0 - Kp + 8LL * c;
        float2 v[8];
        if (s0 >= 0) {
            if (s0 + 8 <= a.n_in) {
                load_chunk8<FMT>(in + s0 * ES, v);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[i] = (s0 + i < a.n_in) ? load_one<FMT>(in, s0 + i) : make_float2(0.0f, 0.0f);
            }
        } else {
            load_chunk8<FMT>(hist + ((long long)a.HL + s0) * ES, v);
        }
        float4 *dst = reinterpret_cast<float4 *>(win + 10 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
    }
    __syncthreads();

    // ---- 8 outputs per thread, sliding 16-sample register window ----
    // window sample index of x[n0 + o - k] is  Kp + 8*tid + o - k   (n0 = tile0 + 8*tid)
    float2 acc[RB_O];
#pragma unroll
    for (int o = 0; o < RB_O; ++o) acc[o] = make_float2(0.0f, 0.0f);
    float2 w[16];
    const int Q = Kp >> 3;
    {
        const float4 *src = reinterpret_cast<const float4 *>(win + 10 * (tid + Q));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = src[i];
            w[8 + 2 * i] = make_float2(f.x, f.y);
            w[8 + 2 * i + 1] = make_float2(f.z, f.w);
        }
    }
    for (int kc = 0; kc < Q; ++kc) {
        const float4 *src = reinterpret_cast<const float4 *>(win + 10 * (tid + Q - kc - 1));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = src[i];
            w[2 * i] = make_float2(f.x, f.y);
            w[2 * i + 1] = make_float2(f.z, f.w);
        }
        float cr[8], ci[8];
        if (!TC) {
            const float4 t0 = *reinterpret_cast<const float4 *>(staps + 8 * kc);
            const float4 t1 = *reinterpret_cast<const float4 *>(staps + 8 * kc + 4);
            cr[0] = t0.x; cr[1] = t0.y; cr[2] = t0.z; cr[3] = t0.w;
            cr[4] = t1.x; cr[5] = t1.y; cr[6] = t1.z; cr[7] = t1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i) ci[i] = 0.0f;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t = *reinterpret_cast<const float4 *>(staps + 16 * kc + 4 * i);
                cr[2 * i] = t.x; ci[2 * i] = t.y; cr[2 * i + 1] = t.z; ci[2 * i + 1] = t.w;
            }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            // reference order: Fir::apply never reads beyond tap K-1 (fir.rs:27-30); the zero padding up to Kp would
            // turn an Inf / NaN sample just outside the window into NaN (Inf * 0) on up to 7 outputs
            if (STRICT && 8 * kc + kk >= a.K) break;
#pragma unroll
            for (int o = 0; o < RB_O; ++o) mac<TC, STRICT>(acc[o], w[8 + o - kk], cr[kk], ci[kk]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) w[8 + i] = w[i];
    }
    __syncthreads();

    // ---- outputs through shared memory for 512-byte warp stores ----
    {
        float4 *dst = reinterpret_cast<float4 *>(win + 10 * tid);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[2 * i].x, acc[2 * i].y, acc[2 * i + 1].x, acc[2 * i + 1].y);
    }
    __syncthreads();
    float2 *out = (float2 *)a.out + (long long)ch * a.out_stride;
    for (int f = tid; f < RB_TILE / 2; f += RB_THREADS) {
        const int j = 2 * f;
        const long long g = tile0 + j;
        const float4 v = *reinterpret_cast<const float4 *>(win + win_idx(j));
        if (g + 1 < a.n_out) {
            *reinterpret_cast<float4 *>(out + g) = v;
        } else if (g < a.n_out) {
            out[g] = make_float2(v.x, v.y);
        }
    }
}

template <typename T>
__global__ void hist_update_kernel(const T *in, const T *hist_old, T *hist_new, int HL, long long n_in,
                                   long long in_stride, long long hist_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (i >= HL) return;
    const long long s = n_in - HL + i;
    const T v = (s >= 0) ? in[(long long)ch * in_stride + s] : hist_old[(long long)ch * hist_stride + HL + s];
    hist_new[(long long)ch * hist_stride + i] = v;
}

__global__ void fill_u16_kernel(uint16_t *p, long long n, uint16_t v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void unpack_kernel(const uint8_t *__restrict__ iq, long long n, float2 *__restrict__ out) {
    // 8 samples (16 bytes in, 64 bytes out) per thread when aligned; scalar otherwise
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long s0 = 8 * c;
    if (s0 >= n) return;
    if (s0 + 8 <= n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(iq + 2 * s0));
        float4 *dst = reinterpret_cast<float4 *>(out + s0);
        dst[0] = make_float4(unpack_byte(q.x, 0), unpack_byte(q.x, 1), unpack_byte(q.x, 2), unpack_byte(q.x, 3));
        dst[1] = make_float4(unpack_byte(q.y, 0), unpack_byte(q.y, 1), unpack_byte(q.y, 2), unpack_byte(q.y, 3));
        dst[2] = make_float4(unpack_byte(q.z, 0), unpack_byte(q.z, 1), unpack_byte(q.z, 2), unpack_byte(q.z, 3));
        dst[3] = make_float4(unpack_byte(q.w, 0), unpack_byte(q.w, 1), unpack_byte(q.w, 2), unpack_byte(q.w, 3));
    } else {
        for (long long s = s0; s < n; ++s)
            out[s] = unpack_iq_u16(*reinterpret_cast<const uint16_t *>(iq + 2 * s));
    }
}

template <int FMT, bool TC, bool STRICT>
int launch_variant(const FirArgs &a, bool use_rb, cudaStream_t st) {
    if (use_rb) {
        const size_t smem = (size_t)((RB_TILE + a.Kp) >> 3) * 80 + (size_t)a.Kp * 4 * (TC ? 2 : 1);
        auto kern = fir_rb_kernel<FMT, TC, STRICT>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_status(e);
        }
        dim3 grid((unsigned)((a.n_out + RB_TILE - 1) / RB_TILE), (unsigned)a.n_ch);
        kern<<<grid, RB_THREADS, smem, st>>>(a);
    } else {
        dim3 grid((unsigned)((a.n_out + 255) / 256), (unsigned)a.n_ch);
        fir_generic_kernel<FMT, TC, STRICT><<<grid, 256, 0, st>>>(a);
    }
    count_launch();
    return launch_status();
}

template <int FMT>
int launch_fmt(const FirArgs &a, bool tc, bool strict, bool use_rb, cudaStream_t st) {
    if (tc) return strict ? launch_variant<FMT, true, true>(a, use_rb, st) : launch_variant<FMT, true, false>(a, use_rb, st);
    return strict ? launch_variant<FMT, false, true>(a, use_rb, st) : launch_variant<FMT, false, false>(a, use_rb, st);
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

int fir_launch(const FirArgs &a, int fmt, bool tc, bool strict, cudaStream_t st, int *path) {
    if (path) *path = strict ? 2 : 1;
    if (a.n_out <= 0) return SDR_OK;
    if (a.n_ch > 65535) return SDR_ERR_INVALID_ARG;
    bool use_rb = (a.D == 1) && (fmt != SDR_FMT_F32) && (a.Kp <= RB_MAX_KP) && (a.HL == a.Kp) &&
                  aligned16(a.in) && aligned16(a.hist) && aligned16(a.out) && (a.out_stride % 2 == 0);
    if (fmt == SDR_FMT_U8IQ) use_rb = use_rb && (a.in_stride % 8 == 0) && (a.hist_stride % 8 == 0);
    if (fmt == SDR_FMT_C64) use_rb = use_rb && (a.in_stride % 2 == 0) && (a.hist_stride % 2 == 0);
    switch (fmt) {
        case SDR_FMT_U8IQ: return launch_fmt<SDR_FMT_U8IQ>(a, tc, strict, use_rb, st);
        case SDR_FMT_C64: return launch_fmt<SDR_FMT_C64>(a, tc, strict, use_rb, st);
        case SDR_FMT_F32:
            if (tc) return SDR_ERR_INVALID_ARG;
            return strict ? launch_variant<SDR_FMT_F32, false, true>(a, false, st)
                          : launch_variant<SDR_FMT_F32, false, false>(a, false, st);
    }
    return SDR_ERR_INVALID_ARG;
}

int fir_hist_update(const void *in, const void *hist_old, void *hist_new, int fmt, int HL, long long n_in,
                    long long in_stride, long long hist_stride, int n_ch, cudaStream_t st) {
    if (HL <= 0) return SDR_OK;
    dim3 grid((unsigned)((HL + 127) / 128), (unsigned)n_ch);
    if (fmt == SDR_FMT_U8IQ)
        hist_update_kernel<uint16_t><<<grid, 128, 0, st>>>((const uint16_t *)in, (const uint16_t *)hist_old,
                                                           (uint16_t *)hist_new, HL, n_in, in_stride, hist_stride);
    else if (fmt == SDR_FMT_C64)
        hist_update_kernel<float2><<<grid, 128, 0, st>>>((const float2 *)in, (const float2 *)hist_old,
                                                         (float2 *)hist_new, HL, n_in, in_stride, hist_stride);
    else
        hist_update_kernel<float><<<grid, 128, 0, st>>>((const float *)in, (const float *)hist_old,
                                                        (float *)hist_new, HL, n_in, in_stride, hist_stride);
    count_launch();
    return launch_status();
}

int fir_fill_hist(void *hist, int fmt, long long n_elems, cudaStream_t st) {
    if (n_elems <= 0) return SDR_OK;
    if (fmt == SDR_FMT_U8IQ) {  // the zero sample is the byte pair (128, 128): (128-128)/128 = 0.0
        fill_u16_kernel<<<(unsigned)((n_elems + 255) / 256), 256, 0, st>>>((uint16_t *)hist, n_elems, 0x8080);
        count_launch();
        return launch_status();
    }
    return cuda_status(cudaMemsetAsync(hist, 0, (size_t)n_elems * (fmt == SDR_FMT_C64 ? 8 : 4), st));
}

int unpack_launch(const uint8_t *iq, size_t n, float *out, cudaStream_t st) {
    if (n == 0) return SDR_OK;
    if (!aligned16(iq) || !aligned16(out)) return SDR_ERR_MISALIGNED;
    const long long chunks = ((long long)n + 7) / 8;
    unpack_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(iq, (long long)n, (float2 *)out);
    count_launch();
    return launch_status();
}

}  // namespace sdr
