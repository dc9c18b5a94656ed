// resample.cu -- rate conversion kernels behind the libsamplerate-shaped sdr_src_* API.
//
// Replaces libsamplerate's src_process as called by SampleRate::process (src/resample.rs:46-67).
// libsamplerate is not in the reference tree; the arithmetic is the "sdr-src" specification of
// DESIGN.md (ZOH / linear / windowed-sinc with libsamplerate's structure, closed-form output
// positions pos + m*step so that every output is independent and can be computed in parallel).
// All position and coefficient arithmetic is f64 with explicit _rn intrinsics (no FMA
// contraction), so a GPU output equals the CPU restatement's bit for bit.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "kernels.h"

namespace sdr {

namespace {

__device__ __forceinline__ double out_position(double pos, long long m, double step) {
    return __dadd_rn(pos, __dmul_rn((double)m, step));
}

__global__ void src_zoh_linear_kernel(SrcLaunch s) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= s.n_out * s.channels) return;
    const long long m = gid / s.channels;
    const int c = (int)(gid % s.channels);
    const double P = out_position(s.pos, m, s.step);
    const double fl = floor(P);
    const long long i = (long long)fl + 1;  // right neighbour, in input coordinates
    const float left = s.v[(i - 1 + s.origin) * s.channels + c];
    if (s.type == SDR_SRC_LINEAR) {
        const float right = s.v[(i + s.origin) * s.channels + c];
        const double f = __dsub_rn(P, fl);
        s.out[gid] = (float)__dadd_rn((double)left, __dmul_rn(f, (double)__fsub_rn(right, left)));
    } else {
        s.out[gid] = left;
    }
}

__global__ void src_sinc_kernel(SrcLaunch s) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= s.n_out * s.channels) return;
    const long long m = gid / s.channels;
    const int c = (int)(gid % s.channels);
    const double T = out_position(s.pos, m, s.step);
    const long long i0 = (long long)floor(T);
    const float *__restrict__ tab = s.table;
    double left = 0.0, right = 0.0;
    for (long long j = i0 - s.wc - 1; j <= i0; ++j) {  // left wing, far -> near
        if (j + s.origin < 0 || j + s.origin >= s.have) continue;
        const double fi = __dmul_rn(__dsub_rn(T, (double)j), s.rq);
        const long long k = (long long)fi;
        if (k >= s.half_len) continue;
        const double fr = __dsub_rn(fi, (double)k);
        const double t0 = (double)__ldg(tab + k), t1 = (double)__ldg(tab + k + 1);
        const double co = __dadd_rn(t0, __dmul_rn(fr, __dsub_rn(t1, t0)));
        left = __dadd_rn(left, __dmul_rn(co, (double)s.v[(j + s.origin) * s.channels + c]));
    }
    for (long long j = i0 + s.wc + 1; j > i0; --j) {  // right wing, far -> near
        if (j + s.origin < 0 || j + s.origin >= s.have) continue;
        const double fi = __dmul_rn(__dsub_rn((double)j, T), s.rq);
        const long long k = (long long)fi;
        if (k >= s.half_len) continue;
        const double fr = __dsub_rn(fi, (double)k);
        const double t0 = (double)__ldg(tab + k), t1 = (double)__ldg(tab + k + 1);
        const double co = __dadd_rn(t0, __dmul_rn(fr, __dsub_rn(t1, t0)));
        right = __dadd_rn(right, __dmul_rn(co, (double)s.v[(j + s.origin) * s.channels + c]));
    }
    s.out[gid] = (float)__dmul_rn(s.rho, __dadd_rn(left, right));
}

// ---- polyphase fast path ---------------------------------------------------------------------------------------
// When step = 1/ratio is k/q with q in {1,2,4,8,16} and the start position lies on a 2^-16 grid, every position
// pos + m*step and every distance T - j is EXACT in f64, so the coefficient of tap distance d depends only on
// (m mod q, d): q tables of wc + 2 entries per wing, computed once per call with the very operations of
// src_sinc_kernel (bit-identical coefficients), and the per-output work is reduced to the f64 multiply-add chain in
// the specified order (left wing far -> near, right wing far -> near).  The reference's ratios (0.2, 0.08, 1/3 is not
// dyadic and stays on the per-tap kernel) hit it.  A CTA stages its window of input frames in shared memory as f64.
struct PolyArgs {
    SrcLaunch s;
    int q;
    double frac[16];   // T - floor(T) for outputs m = 0..q-1
    int dmaxL[16], dmaxR[16];  // largest tap distance inside the table (k < half_len), per phase
    long long Wp;      // row pitch of the coefficient tables (even, >= wc + 4)
    int S;             // q * step (an integer by construction)
    int zero;          // 0
    unsigned ib_lo, ib_hi;  // R-outputs kernel: CTAs ib_lo..ib_hi are interior (see the kernel); lo > hi = none
};

// coefficient tables: [copy][wing][phase][Wp].  copy 0 holds c[d] at d, copy 1 holds c[d] at d + 1, so that the pair
// (c[d-1], c[d]) is one aligned 16-byte load from whichever copy puts it on an even index.
__device__ __forceinline__ const double *coef_row(const PolyArgs &a, int copy, int wing, int p) {
    return a.s.coef + (((long long)copy * 2 + wing) * a.q + p) * a.Wp;
}
static const double *coef_row_host(const PolyArgs &a, int copy, int wing, int p) {
    return a.s.coef + (((long long)copy * 2 + wing) * a.q + p) * a.Wp;
}

__global__ void src_sinc_coef_kernel(PolyArgs a) {
    const long long W = a.s.wc + 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 2LL * a.q * W) return;
    const int wing = (int)(gid / (a.q * W));
    const int p = (int)((gid / W) % a.q);
    const long long d = gid % W;
    // left: T - j = d + frac ; right: j - T = d - frac   (both exact)
    const double dist = wing == 0 ? __dadd_rn((double)d, a.frac[p]) : __dsub_rn((double)d, a.frac[p]);
    double co = 0.0;
    if (dist >= 0.0) {
        const double fi = __dmul_rn(dist, a.s.rq);
        const long long k = (long long)fi;
        if (k < a.s.half_len) {
            const double fr = __dsub_rn(fi, (double)k);
            const double t0 = (double)__ldg(a.s.table + k), t1 = (double)__ldg(a.s.table + k + 1);
            co = __dadd_rn(t0, __dmul_rn(fr, __dsub_rn(t1, t0)));
        }
    }
    const_cast<double *>(coef_row(a, 0, wing, p))[d] = co;
    const_cast<double *>(coef_row(a, 1, wing, p))[d + 1] = co;
}

template <int CH>
__global__ void __launch_bounds__(128) src_sinc_poly_kernel(PolyArgs a) {
    extern __shared__ double xs[];
    const SrcLaunch &s = a.s;
    const int tid = threadIdx.x;
    const long long mb0 = (long long)blockIdx.x * 128;
    const long long mb1 = min(mb0 + 127, s.n_out - 1);
    const long long i_first = (long long)floor(out_position(s.pos, mb0, s.step));
    const long long i_last = (long long)floor(out_position(s.pos, mb1, s.step));
    long long jlo = i_first - s.wc - 1, jhi = i_last + s.wc + 1;
    if (jlo < 0) jlo = 0;
    if (jhi > s.have - 1) jhi = s.have - 1;
    const long long nfr = jhi - jlo + 1;
    for (long long i = tid; i < nfr * CH; i += 128) xs[i] = (double)s.v[jlo * CH + i];
    __syncthreads();
    const long long m = mb0 + tid;
    if (m >= s.n_out) return;
    const double T = out_position(s.pos, m, s.step);
    const long long i0 = (long long)floor(T);
    const int p = (int)(m % a.q);
    double left[CH], right[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { left[c] = 0.0; right[c] = 0.0; }
    {   // left wing, far -> near: j = i0 - d, d from dmax down to 0, inside [0, have)
        long long dhi = min((long long)a.dmaxL[p], min(s.wc + 1, i0));
        long long dlo = max(0LL, i0 - (s.have - 1));
        const double *co = coef_row(a, 0, 0, p);
        for (long long d = dhi; d >= dlo; --d) {
            const double c0 = __ldg(co + d);
            const double *x = xs + (i0 - d - jlo) * CH;
#pragma unroll
            for (int c = 0; c < CH; ++c) left[c] = __dadd_rn(left[c], __dmul_rn(c0, x[c]));
        }
    }
    {   // right wing, far -> near: j = i0 + d, d from dmax down to 1
        long long dhi = min((long long)a.dmaxR[p], min(s.wc + 1, s.have - 1 - i0));
        long long dlo = max(1LL, -i0);
        const double *co = coef_row(a, 0, 1, p);
        for (long long d = dhi; d >= dlo; --d) {
            const double c0 = __ldg(co + d);
            const double *x = xs + (i0 + d - jlo) * CH;
#pragma unroll
            for (int c = 0; c < CH; ++c) right[c] = __dadd_rn(right[c], __dmul_rn(c0, x[c]));
        }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) s.out[m * CH + c] = (float)__dmul_rn(s.rho, __dadd_rn(left[c], right[c]));
}

// R outputs per thread.  The per-tap kernel above is bound by the load pipe: every tap of every output costs one
// 16-byte shared load of its input frame (4 wavefronts per warp) plus one coefficient load, against 2 cycles of f64
// issue.  Outputs m, m + q, m + 2q, ... share a phase and their windows are shifted by the INTEGER S = q * step, so a
// thread that owns R of them loads each input frame once and feeds R accumulator sets (coefficient distance d + r*S
// for output r).  Each output still sees its own taps in the specified order (left wing far -> near, right wing far
// -> near), so the bits equal the R = 1 kernel's.  R is odd: with S odd the lane stride R*S stays odd and the shared
// loads stay conflict-free.  The edges (start/end of the stream, ragged phases) run in a predicated loop either side
// of the all-outputs-active steady loop.
// Coefficients through the uniform datapath.  A coefficient is the same for the 32 lanes of a warp (q = 1: one phase),
// yet a global or shared load still writes 32 copies back into the register file -- 256 B per tap and output, which at
// the load pipe's 128 B/cycle/SM costs as much as the tap's four f64 instructions and made the kernel writeback-bound
// (ncu: lsu_writeback_active 77 % of active cycles, FP64 pipe 54 %).  From __constant__ memory with a warp-uniform
// index the compiler emits LDCU.64 into a uniform register that DMUL takes as an operand: no load-pipe traffic at all.
constexpr int kConstCoef = 7680;  // doubles (60 KB of the 64 KB bank)
__constant__ double c_coef[kConstCoef];

template <typename XT, int CH>
__device__ __forceinline__ void load_frame(const XT *x, double (&v)[CH]) {
    if constexpr (CH == 2 && sizeof(XT) == 4) {
        const float2 f = *reinterpret_cast<const float2 *>(x);
        v[0] = (double)f.x;
        v[1] = (double)f.y;
    } else if constexpr (CH == 2) {
        const double2 f = *reinterpret_cast<const double2 *>(x);
        v[0] = f.x;
        v[1] = f.y;
    } else {
        v[0] = (double)x[0];
    }
}

template <int K, int KEND, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (K < KEND) {
        f(std::integral_constant<int, K>{});
        static_for<K + 1, KEND>(f);
    }
}

// taps t = t_from down to t_to of one wing for outputs RLO..RHI of the thread: frame xb + t * xstep, coefficient
// c_coef[cbase + t + r*S].  Every bound and coefficient index is warp-uniform (kernel parameters only); the frame
// address is the only per-thread quantity.
template <int CH, int R, typename XT, int RLO, int RHI, int UNR>
__device__ __forceinline__ void uni_seg(double (&acc)[R][CH], const XT *xb, int xstep, int cbase, int S, int t_from,
                                        int t_to, int tz) {
#pragma unroll UNR
    for (int t = t_from; t >= t_to; --t) {
        double xv[CH];
        load_frame<XT, CH>(xb + t * xstep, xv);
#pragma unroll
        for (int r = RLO; r <= RHI; ++r) {
            const double c0 = c_coef[cbase + t + r * S + tz];
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[r][c] = __dadd_rn(acc[r][c], __dmul_rn(c0, xv[c]));
        }
    }
}

// Interior CTAs of the R-outputs kernel (the host finds them: every output exists and no wing is clipped by the ends
// of the stream; q == 1).  Loop bounds and coefficient indices derive from kernel parameters only, so the compiler
// keeps them in uniform registers and reads the coefficients with LDCU straight into DMUL operands.
template <int CH, int R, typename XT>
__device__ __forceinline__ void poly_interior_cta(const PolyArgs &a, long long blk, XT *xs) {
    const SrcLaunch &s = a.s;
    const int tid = threadIdx.x;
    const int opb = 128 * R;
    const long long mb0 = blk * opb;
    const long long i_first = (long long)floor(out_position(s.pos, mb0, s.step));
    const long long i_last = (long long)floor(out_position(s.pos, mb0 + opb - 1, s.step));
    const long long jlo = i_first - s.wc - 1, jhi = i_last + s.wc + 1;  // inside [0, have) for an interior CTA
    const long long nfr = jhi - jlo + 1;
    for (long long i = tid; i < nfr * CH; i += 128) xs[i] = (XT)s.v[jlo * CH + i];
    __syncthreads();
    const long long m0 = mb0 + (long long)tid * R;
    const int S = a.S;
    const int DL = (int)min((long long)a.dmaxL[0], s.wc + 1), DR = (int)min((long long)a.dmaxR[0], s.wc + 1);
    const long long i0 = (long long)floor(out_position(s.pos, m0, s.step));
    const XT *xb = xs + (i0 - jlo) * CH;
    // One segment routine serves both wings.  Left wing: frame i0 - t, distance t + r*S, output r active while
    // -r*S <= t <= DL - r*S.  Right wing: frame i0 + u, distance u - r*S, active while 1 + r*S <= u <= DR + r*S; with
    // r' = R-1-r and t = u - (R-1)*S this reads distance t + r'*S, active while 1 - r'*S <= t <= DR - r'*S -- the left
    // wing's shape with lower bound 1, the frames walked in the other direction and the outputs in reverse order.
    // (The wing loop is deliberately NOT unrolled and the ramps are not unrolled either: past a certain code size
    // ptxas gives up on uniform registers for the whole kernel and falls back to per-thread LDC.)
    // A zero the compiler cannot prove uniform (lane * a parameter that is 0).  With a provably uniform index ptxas
    // sometimes emits LDCU into uniform registers and sometimes per-thread LDC, depending on code size; LDCU
    // measured 10 % slower here (its latency is not hidden by the warp scheduler), so the index is made formally
    // per-thread and the loads are always LDC.  Either way they bypass the load/store pipe.
    const int tz = (int)(threadIdx.x & 31) * a.zero;
    double left[R][CH], acc[R][CH];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[r][c] = 0.0;
        const int D = (int)min((long long)(w ? a.dmaxR[0] : a.dmaxL[0]), s.wc + 1);
        const int L0 = w;
        const int cbase = w ? (int)a.Wp : 0;
        const int xstep = w ? CH : -CH;
        const XT *xw = w ? xb + (R - 1) * S * CH : xb;
        static_for<0, R - 1>([&](auto k) {  // ramp-in: outputs 0..k
            uni_seg<CH, R, XT, 0, k.value, 4>(acc, xw, xstep, cbase, S, D - k.value * S, D - (k.value + 1) * S + 1, tz);
        });
        uni_seg<CH, R, XT, 0, R - 1, 4>(acc, xw, xstep, cbase, S, D - (R - 1) * S, L0, tz);  // every output active
        static_for<0, R - 1>([&](auto k) {  // ramp-out: outputs k+1..R-1
            uni_seg<CH, R, XT, k.value + 1, R - 1, 4>(acc, xw, xstep, cbase, S, L0 - k.value * S - 1,
                                                      L0 - (k.value + 1) * S, tz);
        });
        if (w == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < CH; ++c) left[r][c] = acc[r][c];
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c)
            s.out[(m0 + r) * CH + c] = (float)__dmul_rn(s.rho, __dadd_rn(left[r][c], acc[R - 1 - r][c]));
}

template <int CH, int R, typename XT>
__device__ __forceinline__ void poly_general_cta(const PolyArgs &a, long long blk, XT *xs) {
    const SrcLaunch &s = a.s;
    const int tid = threadIdx.x;
    const int opb = 128 * R;  // outputs per CTA; 128 is a multiple of q
    const long long mb0 = blk * opb;
    const long long mb1 = min(mb0 + opb - 1, s.n_out - 1);
    const long long i_first = (long long)floor(out_position(s.pos, mb0, s.step));
    const long long i_last = (long long)floor(out_position(s.pos, mb1, s.step));
    long long jlo = i_first - s.wc - 1, jhi = i_last + s.wc + 1;
    if (jlo < 0) jlo = 0;
    if (jhi > s.have - 1) jhi = s.have - 1;
    const long long nfr = jhi - jlo + 1;
    for (long long i = tid; i < nfr * CH; i += 128) xs[i] = (XT)s.v[jlo * CH + i];
    __syncthreads();
    const int q = a.q;
    const int p = tid % q;
    const long long m0 = mb0 + (long long)(tid / q) * (q * R) + p;
    if (m0 >= s.n_out) return;
    const int S = (int)__dmul_rn((double)q, s.step);  // exact integer by poly_plan
    const long long i0 = (long long)floor(out_position(s.pos, m0, s.step));
    const XT *xb = xs + (i0 - jlo) * CH;  // frame i0 of output 0
    int nr = 0;                               // outputs of this thread inside the call
#pragma unroll
    for (int r = 0; r < R; ++r) nr += (m0 + (long long)r * q < s.n_out) ? 1 : 0;

    double acc[2][R][CH];
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[w][r][c] = 0.0;

#pragma unroll
    for (int w = 0; w < 2; ++w) {
        // wing 0 (left):  frame i0 - t, distance of output r = t + r*S      (sg = +1)
        // wing 1 (right): frame i0 + t, distance of output r = t - r*S      (sg = -1)
        const int sg = w == 0 ? 1 : -1;
        const double *co = coef_row(a, 0, w, p);
        int lo[R], hi[R];          // active range of the loop variable t for output r
        int t_first = INT_MIN, t_last = INT_MAX, st_hi = INT_MAX, st_lo = INT_MIN;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long ir = i0 + (long long)r * S;
            long long dhi, dlo;
            if (w == 0) {
                dhi = min((long long)a.dmaxL[p], min(s.wc + 1, ir));
                dlo = max(0LL, ir - (s.have - 1));
            } else {
                dhi = min((long long)a.dmaxR[p], min(s.wc + 1, s.have - 1 - ir));
                dlo = max(1LL, -ir);
            }
            if (r >= nr) { dhi = -1; dlo = 0; }  // not an output of this call: empty range
            if (dhi >= dlo) {
                hi[r] = (int)dhi - sg * r * S;
                lo[r] = (int)dlo - sg * r * S;
                t_first = max(t_first, hi[r]);
                t_last = min(t_last, lo[r]);
                st_hi = min(st_hi, hi[r]);
                st_lo = max(st_lo, lo[r]);
            } else {
                hi[r] = INT_MIN;
                lo[r] = INT_MAX;
                st_hi = INT_MIN;  // no steady part when an output is empty
            }
        }
        if (t_first == INT_MIN) continue;  // nothing on this wing
        if (st_hi < st_lo) { st_hi = t_last - 1; st_lo = t_last; }  // steady part empty: one predicated sweep
        auto sweep = [&](int from, int to) {  // predicated: t = from down to to
            for (int t = from; t >= to; --t) {
                bool any = false;
#pragma unroll
                for (int r = 0; r < R; ++r) any |= (t <= hi[r] && t >= lo[r]);
                if (!any) continue;
                double xv[CH];
                load_frame<XT, CH>(xb - (long long)sg * t * CH, xv);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (t <= hi[r] && t >= lo[r]) {
                        const double c0 = __ldg(co + t + sg * r * S);
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            acc[w][r][c] = __dadd_rn(acc[w][r][c], __dmul_rn(c0, xv[c]));
                    }
                }
            }
        };
        if (st_hi == t_last - 1) {
            sweep(t_first, t_last);
        } else {
            sweep(t_first, st_hi + 1);
            // every output active: taps t and t - 1 per iteration, their coefficients (c[d - 1], c[d]) as one aligned
            // 16-byte load from the copy whose index of c[d - 1] is even (fixed per output for the whole loop)
            const double *pr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d1 = st_hi + sg * r * S - 1;  // distance of the pair's low element at the first iteration
                pr[r] = coef_row(a, d1 & 1, w, p) + (d1 & 1) + sg * r * S - 1;  // + t -> the pair of (t, t - 1)
            }
            int t = st_hi;
            if (t - 1 >= st_lo) {
                // software pipeline: the loads of iteration i + 1 are issued before the arithmetic of iteration i
                double2 cp[R];
                double xv0[CH], xv1[CH];
                auto load = [&](int tt, double2 (&c2)[R], double (&a0)[CH], double (&a1)[CH]) {
#pragma unroll
                    for (int r = 0; r < R; ++r) c2[r] = __ldg(reinterpret_cast<const double2 *>(pr[r] + tt));
                    load_frame<XT, CH>(xb - (long long)sg * tt * CH, a0);
                    load_frame<XT, CH>(xb - (long long)sg * (tt - 1) * CH, a1);
                };
                load(t, cp, xv0, xv1);
                for (; t - 1 >= st_lo; t -= 2) {
                    double2 cn[R];
                    double xn0[CH], xn1[CH];
                    const int tn = t - 3 >= st_lo ? t - 2 : t;  // last iteration: a harmless reload
                    load(tn, cn, xn0, xn1);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            acc[w][r][c] = __dadd_rn(acc[w][r][c], __dmul_rn(cp[r].y, xv0[c]));
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            acc[w][r][c] = __dadd_rn(acc[w][r][c], __dmul_rn(cp[r].x, xv1[c]));
#pragma unroll
                    for (int r = 0; r < R; ++r) cp[r] = cn[r];
#pragma unroll
                    for (int c = 0; c < CH; ++c) { xv0[c] = xn0[c]; xv1[c] = xn1[c]; }
                }
            }
            st_lo = t + 1;  // an odd leftover tap joins the closing sweep
            sweep(st_lo - 1, t_last);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nr) {
            const long long m = m0 + (long long)r * q;
#pragma unroll
            for (int c = 0; c < CH; ++c)
                s.out[m * CH + c] = (float)__dmul_rn(s.rho, __dadd_rn(acc[0][r][c], acc[1][r][c]));
        }
    }
}

// One grid for both kinds of CTA.  The few CTAs that touch the ends of the stream (predicated loops, several times
// slower) come FIRST in the grid so that they run alongside the interior ones instead of forming a tail.
template <int CH, int R, typename XT>
__global__ void __launch_bounds__(128) src_sinc_poly_r_kernel(PolyArgs a) {
    extern __shared__ __align__(16) unsigned char xs_raw[];
    XT *xs = reinterpret_cast<XT *>(xs_raw);
    const bool have_int = a.ib_lo <= a.ib_hi;
    const unsigned n_int = have_int ? a.ib_hi - a.ib_lo + 1 : 0;
    const unsigned n_edge = gridDim.x - n_int;
    if (blockIdx.x < n_edge) {
        const unsigned e = blockIdx.x;
        const long long blk = have_int && e >= a.ib_lo ? (long long)e + n_int : (long long)e;
        poly_general_cta<CH, R, XT>(a, blk, xs);
    } else {
        poly_interior_cta<CH, R, XT>(a, (long long)a.ib_lo + (blockIdx.x - n_edge), xs);
    }
}

// ---- coefficient design (host, f64, done once per converter type) ---------------------------
struct SincSpec { int increment; size_t half_len; double fc, beta; };
// geometry of libsamplerate's three tables (entries per zero crossing, half length); the
// Kaiser-windowed-sinc shape is this project's own (~97 / 97 / 145 dB).
const SincSpec kSpec[3] = {
    {2381, 340239, 0.9666, 15.0},  // SDR_SRC_SINC_BEST_QUALITY
    {491, 22438, 0.932, 9.73},     // SDR_SRC_SINC_MEDIUM_QUALITY
    {128, 2464, 0.84, 9.73},       // SDR_SRC_SINC_FASTEST
};

double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

std::vector<float> g_tab[3];
std::once_flag g_once[3];

}  // namespace

size_t src_sinc_table_host(int type, const float **table, int *increment) {
    if (type < 0 || type > 2) return 0;
    std::call_once(g_once[type], [type]() {
        const SincSpec &s = kSpec[type];
        std::vector<float> &t = g_tab[type];
        t.resize(s.half_len + 2);
        const double i0b = bessel_i0(s.beta);
        for (size_t k = 0; k <= s.half_len + 1; ++k) {
            double v = 0.0;
            if (k <= s.half_len) {
                const double u = (double)k / (double)s.increment;
                const double a = M_PI * s.fc * u;
                const double sinc = (k == 0) ? 1.0 : std::sin(a) / a;
                const double r = (double)k / (double)s.half_len;
                const double w = bessel_i0(s.beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
                v = s.fc * sinc * w;
            }
            t[k] = (float)v;
        }
    });
    if (table) *table = g_tab[type].data();
    if (increment) *increment = kSpec[type].increment;
    return kSpec[type].half_len;
}

double src_sinc_wing(int type, double ratio, double *rq, double *rho, long long *wc) {
    const SincSpec &sp = kSpec[type];
    const double r = ratio < 1.0 ? ratio : 1.0;
    const double q = r * (double)sp.increment;
    const double wing = (double)sp.half_len / q;
    if (rq) *rq = q;
    if (rho) *rho = r;
    if (wc) *wc = (long long)std::ceil(wing);
    return wing;
}

// ---- tensor-core path for integer steps (see kernels.h) -----------------------------------------------------------
// taps of the equivalent FIR: g[k] = rho * coef(|W - k|), k = 0 .. 2W, coef = the spec's linear interpolation in the
// half table evaluated in f64 (the same expression as src_sinc_coef_kernel / the per-tap kernel), rounded to f32 once.
int src_fast_branch_len(long long wc, int S) { return (int)((2 * (wc + 1) + 1 + S - 1) / S); }

void src_fast_taps(int type, double ratio, int S, std::vector<float> &branches) {
    double rq, rho;
    long long wc;
    src_sinc_wing(type, ratio, &rq, &rho, &wc);
    const float *tab = nullptr;
    int inc = 0;
    const long long half_len = (long long)src_sinc_table_host(type, &tab, &inc);
    const long long W = wc + 1, Kt = 2 * W + 1;
    const int Kb = src_fast_branch_len(wc, S);
    branches.assign((size_t)S * Kb, 0.0f);
    for (long long k = 0; k < Kt; ++k) {
        const long long d = k <= W ? W - k : k - W;
        const double fi = (double)d * rq;
        const long long idx = (long long)fi;
        double co = 0.0;
        if (idx < half_len) {
            const double fr = fi - (double)idx;
            co = (double)tab[idx] + fr * ((double)tab[idx + 1] - (double)tab[idx]);
        }
        branches[(size_t)(k % S) * Kb + (size_t)(k / S)] = (float)(rho * co);
    }
}

__global__ void src_fast_sum_kernel(const float2 *__restrict__ br, long long pitch, int S, float2 *__restrict__ out, long long n_out) {
    // two frames (16 bytes) per thread; the last thread takes the odd frame alone
    const long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_out) return;
    if (i + 1 < n_out) {
        float4 acc = __ldg(reinterpret_cast<const float4 *>(br + i));
        for (int s = 1; s < S; ++s) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(br + (long long)s * pitch + i));
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        *reinterpret_cast<float4 *>(out + i) = acc;
    } else {
        float2 acc = __ldg(br + i);
        for (int s = 1; s < S; ++s) {
            const float2 v = __ldg(br + (long long)s * pitch + i);
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
        }
        out[i] = acc;
    }
}

// pitch (frames) even and the rows / out 16-byte aligned
int src_fast_sum(const float *branches, long long pitch, int S, float *out, long long n_out, cudaStream_t st) {
    const long long nt = (n_out + 1) / 2;
    src_fast_sum_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>((const float2 *)branches, pitch, S, (float2 *)out, n_out);
    count_launch();
    return launch_status();
}

// the polyphase path applies when every position of the call is exact in f64 (see above); fills PolyArgs
static bool poly_plan(const SrcLaunch &s, PolyArgs &a) {
    static const bool disabled = std::getenv("SDR_SRC_NO_POLY") != nullptr;  // A/B switch for the parity test
    if (disabled || !s.coef || s.channels > 2 || s.origin != 0 || s.wc + 2 > (1 << 20)) return false;
    int q = 0;
    for (int c : {1, 2, 4, 8, 16}) {
        const double v = s.step * c;
        if (v == std::floor(v) && v < 1e9) { q = c; break; }
    }
    if (!q) return false;
    const double g = s.pos * 65536.0;
    if (g != std::floor(g) || std::fabs(s.pos) > 1e9 || (double)s.n_out * s.step > 1e9) return false;
    a.s = s;
    a.q = q;
    for (int p = 0; p < q; ++p) {
        const double T = s.pos + (double)p * s.step;  // exact
        a.frac[p] = T - std::floor(T);
        // largest distance whose table index stays inside: same f64 products as the kernel
        int dl = -1, dr = -1;
        for (long long d = 0; d <= s.wc + 1; ++d) {
            const double fl = ((double)d + a.frac[p]) * s.rq;
            if ((long long)fl < s.half_len) dl = (int)d; else break;
        }
        for (long long d = 1; d <= s.wc + 1; ++d) {
            const double fr = ((double)d - a.frac[p]) * s.rq;
            if ((long long)fr < s.half_len) dr = (int)d; else break;
        }
        a.dmaxL[p] = dl;
        a.dmaxR[p] = dr;
    }
    for (int p = q; p < 16; ++p) { a.frac[p] = 0.0; a.dmaxL[p] = -1; a.dmaxR[p] = -1; }
    return true;
}

int src_launch(const SrcLaunch &s, cudaStream_t st) {
    const long long total = s.n_out * s.channels;
    if (total <= 0) return SDR_OK;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (s.type == SDR_SRC_ZERO_ORDER_HOLD || s.type == SDR_SRC_LINEAR) {
        src_zoh_linear_kernel<<<grid, 128, 0, st>>>(s);
        count_launch();
        return launch_status();
    }
    PolyArgs a;
    if (poly_plan(s, a)) {
        static const int r_force = std::getenv("SDR_SRC_R") ? std::atoi(std::getenv("SDR_SRC_R")) : 0;  // A/B switch
        a.Wp = (s.wc + 5) & ~1LL;
        auto smem_for = [&](int r) {
            const size_t frames = (size_t)((128.0 * r - 1.0) * s.step) + 2 * (size_t)s.wc + 8;
            return frames * s.channels * sizeof(double);
        };
        // R outputs per thread when the call is long enough to fill the GPU with such CTAs and the window fits
        int R = 1;
        for (int r : {3, 5}) {  // 3 measured faster than 5 (more resident warps); 5 stays selectable for experiments
            if (r_force && r != r_force) continue;
            if (smem_for(r) <= 100 * 1024 && (r_force || s.n_out >= 128LL * r * 64)) { R = r; break; }
        }
        if (r_force == 1) R = 1;
        const size_t smem = smem_for(R);
        if (smem <= 200 * 1024) {
            const long long ncoef = 2LL * a.q * (s.wc + 2);
            src_sinc_coef_kernel<<<(unsigned)((ncoef + 255) / 256), 256, 0, st>>>(a);
            count_launch();
            if (R > 1) {
                static const bool no_uni = std::getenv("SDR_SRC_NO_UNIFORM") != nullptr;  // A/B switch
                const bool uni = !no_uni && a.q == 1 && 2 * a.Wp <= kConstCoef &&
                                 std::min(a.dmaxL[0], a.dmaxR[0]) >= (long long)R * (long long)s.step;
                void (*kr)(PolyArgs) =
                    s.channels == 1 ? (R == 5 ? src_sinc_poly_r_kernel<1, 5, double> : src_sinc_poly_r_kernel<1, 3, double>)
                                    : (R == 5 ? src_sinc_poly_r_kernel<2, 5, double> : src_sinc_poly_r_kernel<2, 3, double>);
                const long long opb = 128LL * R, nblk = (s.n_out + opb - 1) / opb;
                a.S = (int)(a.q * s.step);
                a.zero = 0;
                a.ib_lo = 1;
                a.ib_hi = 0;
                if (uni) {
                    // interior CTAs: all outputs exist and no wing is clipped by the ends of the stream.  i_first and
                    // i_last grow with the CTA index, so the interior CTAs form one contiguous range.
                    auto i_first = [&](long long b) { return (long long)std::floor(s.pos + (double)(b * opb) * s.step); };
                    auto is_interior = [&](long long b) {
                        if ((b + 1) * opb > s.n_out) return false;
                        const long long i_last = (long long)std::floor(s.pos + (double)((b + 1) * opb - 1) * s.step);
                        return i_first(b) >= s.wc + 1 && i_last + s.wc + 1 <= s.have - 1;
                    };
                    long long lo = 0, hi = nblk - 1;
                    while (lo < nblk && i_first(lo) < s.wc + 1) ++lo;
                    while (hi >= lo && !is_interior(hi)) --hi;
                    if (lo <= hi && is_interior(lo)) { a.ib_lo = (unsigned)lo; a.ib_hi = (unsigned)hi; }
                }
                const bool have_int = a.ib_lo <= a.ib_hi;
                cudaError_t e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return cuda_status(e);
                if (have_int) {
                    // one __constant__ table per device: a launch on another stream must not overwrite it while an
                    // earlier kernel still reads it, so the copy waits for the previous user's event
                    static std::mutex mu;
                    static cudaEvent_t last_use[64] = {};
                    int dev = 0;
                    cudaGetDevice(&dev);
                    std::lock_guard<std::mutex> lk(mu);
                    cudaEvent_t &ev = last_use[dev & 63];
                    if (!ev) {
                        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return cuda_status(cudaGetLastError());
                    } else {
                        cudaStreamWaitEvent(st, ev, 0);
                    }
                    e = cudaMemcpyToSymbolAsync(c_coef, coef_row_host(a, 0, 0, 0), 2 * a.Wp * sizeof(double), 0,
                                                cudaMemcpyDeviceToDevice, st);
                    if (e != cudaSuccess) return cuda_status(e);
                    kr<<<(unsigned)nblk, 128, smem, st>>>(a);
                    count_launch();
                    cudaEventRecord(ev, st);
                } else {
                    kr<<<(unsigned)nblk, 128, smem, st>>>(a);
                    count_launch();
                }
                return launch_status();
            }
            auto kern = s.channels == 1 ? src_sinc_poly_kernel<1> : src_sinc_poly_kernel<2>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_status(e);
            kern<<<(unsigned)((s.n_out + 127) / 128), 128, smem, st>>>(a);
            count_launch();
            return launch_status();
        }
    }
    src_sinc_kernel<<<grid, 128, 0, st>>>(s);
    count_launch();
    return launch_status();
}

}  // namespace sdr
