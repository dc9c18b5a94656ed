// pll.cu -- K4: batched PLL, one lane per independent stream (the loop is a recurrence:
// the NCO value computed at sample n is an input of sample n+1, src/filter/pll.rs:70-76).
//
// Replaces Pll::apply (src/filter/pll.rs:70-85) with its Biquad sub-filters
// (src/filter/biquad.rs:43-56) for n_streams streams at once.
//
// Every f32 operation is issued in the reference's order with explicit _rn intrinsics so the
// compiler cannot contract mul+add into FMA (rustc never does).  The three transcendental calls
// (atan2, cos, sin) are the only place where results can differ from the host libm:
//   default   : evaluated in f64 and rounded to f32 (differs from glibc only at rare near-ties)
//   FAST_MATH : CUDA's f32 atan2f / sincosf (1-2 ulp)
// Data movement: a warp moves [32 streams x 32 samples] tiles between HBM and shared memory with
// 256-byte row segments, so the lane-per-stream inner loop reads shared memory, not strided HBM.
#include "kernels.h"

namespace sdr {

namespace {

constexpr float PI_F = 3.14159265358979323846f;
constexpr int PLL_CHUNK = 32;  // samples per staged tile

struct Biquad1 {
    float b0, b1, b2, na1, na2;
    int kind;
};

__device__ __forceinline__ float bq_apply(const Biquad1 &c, float v, float &x1, float &x2, float &y1, float &y2) {
    if (c.kind == SDR_BQ_IDENTITY) return v;
    // out = 0; out += v*b0; out += x1*b1; out += x2*b2; out += y1*na1; out += y2*na2  (biquad.rs:44-49)
    float out = __fadd_rn(0.0f, __fmul_rn(v, c.b0));
    out = __fadd_rn(out, __fmul_rn(x1, c.b1));
    out = __fadd_rn(out, __fmul_rn(x2, c.b2));
    out = __fadd_rn(out, __fmul_rn(y1, c.na1));
    out = __fadd_rn(out, __fmul_rn(y2, c.na2));
    x2 = x1; x1 = v;
    y2 = y1; y1 = out;
    return out;
}

template <bool FAST>
__device__ __forceinline__ void pll_step(const PllParams &p, const Biquad1 &lf, const Biquad1 &of, const Biquad1 &kf,
                                         PllState &s, float xr, float xi, float &out, uint8_t &locked) {
    // c = value * self.value.conj()           (pll.rs:71)   other = (vre, -vim)
    const float o_re = s.vre, o_im = -s.vim;
    const float cr = __fsub_rn(__fmul_rn(xr, o_re), __fmul_rn(xi, o_im));
    const float ci = __fadd_rn(__fmul_rn(xr, o_im), __fmul_rn(xi, o_re));
    // loopfilter.apply(c): Biquad<f32, Complex<f32>> acts on re and im independently
    const float lr = bq_apply(lf, cr, s.lx1r, s.lx2r, s.ly1r, s.ly2r);
    const float li = bq_apply(lf, ci, s.lx1i, s.lx2i, s.ly1i, s.ly2i);
    // phasedif = arg * gain                   (pll.rs:72)
    float arg;
    if (FAST) arg = atan2f(li, lr);
    else arg = (float)atan2((double)li, (double)lr);
    const float phasedif = __fmul_rn(arg, p.gain);
    // nphase += reference + phasedif; nphase = nphase.fract()      (pll.rs:73-74)
    float nph = __fadd_rn(s.nphase, __fadd_rn(p.reference, phasedif));
    nph = __fsub_rn(nph, truncf(nph));
    s.nphase = nph;
    // phase = 2.0 * PI * nphase; value = from_polar(1.0, phase)     (pll.rs:75-76)
    const float phase = __fmul_rn(2.0f * PI_F, nph);
    if (FAST) {
        float sn, cs;
        sincosf(phase, &sn, &cs);
        s.vre = cs;
        s.vim = sn;
    } else {
        double sn, cs;
        sincos((double)phase, &sn, &cs);
        s.vre = (float)cs;
        s.vim = (float)sn;
    }
    // locked = lockfilter.apply(c.re); output = outputfilter.apply(phasedif * rate)   (pll.rs:78-79)
    const float lk = bq_apply(kf, cr, s.kx1, s.kx2, s.ky1, s.ky2);
    out = bq_apply(of, __fmul_rn(phasedif, p.rate), s.ox1, s.ox2, s.oy1, s.oy2);
    locked = lk > 0.01f ? 1 : 0;
}

__device__ __forceinline__ Biquad1 make_bq(const float *c, int kind) {
    Biquad1 b;
    b.b0 = c[0]; b.b1 = c[1]; b.b2 = c[2]; b.na1 = c[3]; b.na2 = c[4];
    b.kind = kind;
    return b;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// One warp (= one CTA) owns 32 streams.  [32 streams x 32 samples] tiles are staged through shared memory: the
// next tile streams in with cp.async (256-byte row segments) while the lanes step the current one, so the
// sequential per-stream loop never waits on HBM; outputs leave as 128-byte row segments.
template <bool FAST>
__global__ void __launch_bounds__(32) pll_kernel(const float2 *__restrict__ in, long long n, long long in_stride,
                                                 float *__restrict__ out, uint8_t *__restrict__ locked,
                                                 long long out_stride, const PllParams *__restrict__ params,
                                                 int params_shared, PllState *__restrict__ state, int n_streams) {
    __shared__ float2 s_in[2][32][PLL_CHUNK + 1];
    __shared__ float s_out[32][PLL_CHUNK + 1];
    __shared__ uint8_t s_lk[32][PLL_CHUNK + 4];
    const int lane = threadIdx.x;
    const int stream0 = blockIdx.x * 32;
    if (stream0 >= n_streams) return;
    const int my = stream0 + lane;
    const bool live = my < n_streams;
    PllParams p = params[params_shared ? 0 : (live ? my : stream0)];
    PllState st = state[live ? my : stream0];
    const Biquad1 lf = make_bq(p.lc, p.lk), of = make_bq(p.oc, p.ok), kf = make_bq(p.kc, p.kk);
    const int nrows = min(32, n_streams - stream0);
    const long long ntiles = (n + PLL_CHUNK - 1) / PLL_CHUNK;

    auto issue = [&](long long tile, int buf) {
        const long long base = tile * PLL_CHUNK;
        const int cnt = (int)min((long long)PLL_CHUNK, n - base);
        if (lane < cnt) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
                if (r < nrows) cp_async8(&s_in[buf][r][lane], in + (long long)(stream0 + r) * in_stride + base + lane);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (long long tile = 0; tile < ntiles; ++tile) {
        const int buf = (int)(tile & 1);
        const long long base = tile * PLL_CHUNK;
        const int cnt = (int)min((long long)PLL_CHUNK, n - base);
        if (tile + 1 < ntiles) {
            issue(tile + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        if (live) {
            for (int i = 0; i < cnt; ++i) {
                const float2 x = s_in[buf][lane][i];
                float o;
                uint8_t l;
                pll_step<FAST>(p, lf, of, kf, st, x.x, x.y, o, l);
                s_out[lane][i] = o;
                s_lk[lane][i] = l;
            }
        }
        __syncwarp();
        if (lane < cnt) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
                if (r < nrows) {
                    out[(long long)(stream0 + r) * out_stride + base + lane] = s_out[r][lane];
                    locked[(long long)(stream0 + r) * out_stride + base + lane] = s_lk[r][lane];
                }
        }
        __syncwarp();
    }
    if (live) state[my] = st;
}

}  // namespace

int pll_launch(const float2 *in, long long n, long long in_stride, float *out, uint8_t *locked, long long out_stride,
               const PllParams *params, int params_shared, PllState *state, int n_streams, bool fast_math,
               cudaStream_t st) {
    if (n <= 0 || n_streams <= 0) return SDR_OK;
    const unsigned grid = (unsigned)((n_streams + 31) / 32);
    if (fast_math)
        pll_kernel<true><<<grid, 32, 0, st>>>(in, n, in_stride, out, locked, out_stride, params, params_shared, state, n_streams);
    else
        pll_kernel<false><<<grid, 32, 0, st>>>(in, n, in_stride, out, locked, out_stride, params, params_shared, state, n_streams);
    count_launch();
    return launch_status();
}

}  // namespace sdr
