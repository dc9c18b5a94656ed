// kernels.h -- internal launch interface between the C-ABI layer (api.cu) and the kernel files.
#pragma once
#include "common.cuh"

namespace sdr {

// ---------------------------------------------------------------------------------------
// FIR (fir.cu)
// ---------------------------------------------------------------------------------------
struct FirArgs {
    const void *in;    // n_ch rows of n_in elements (raw input format)
    const void *hist;  // n_ch rows of HL elements, raw input format; hist[HL-1] is the newest
    void *out;         // n_ch rows of n_out elements (c64 or f32)
    const float *taps; // device; Kp entries (real) or Kp (re,im) pairs, zero padded beyond K
    long long n_in, in_stride, out_stride, hist_stride, n_out;
    long long first;   // index (in new-input coordinates) of the first kept output
    int K, Kp, HL, D, n_ch;
    int accumulate = 0;  // c64 tcgen05 kernel only: out += result
    // c64 tcgen05 kernel only, gather mode (the polyphase branches of the rational resampler): when in_step != 0, sample u
    // of channel c is in[in_step * u + in_off - c] for indices inside [0, in_limit) and zero outside (no hist, no strides)
    long long in_step = 0, in_off = 0, in_limit = 0;
};
// fmt: sdr_format_t.  Picks the register-blocked kernel when D == 1 and alignment allows,
// the generic one otherwise.  *path: 1 = direct (FMA), 2 = strict order.
int fir_launch(const FirArgs &a, int fmt, bool taps_complex, bool strict, cudaStream_t st, int *path);
// hist_new = last HL elements of concat(hist_old, in[0..n_in))
int fir_hist_update(const void *in, const void *hist_old, void *hist_new, int fmt, int HL, long long n_in,
                    long long in_stride, long long hist_stride, int n_ch, cudaStream_t st);
int fir_fill_hist(void *hist, int fmt, long long n_elems, cudaStream_t st);  // "zero" samples (128 for u8)

int unpack_launch(const uint8_t *iq, size_t n, float *out, cudaStream_t st);

// tensor-core Toeplitz FIR for u8 IQ input (fir_tc.cu)
}  // namespace sdr
#include <vector>
namespace sdr {
int fir_tc_ksteps(int K, int D, bool taps_complex);
// returns the exact power-of-two epilogue scale
float fir_tc_build_tables(const float *taps, int K, bool taps_complex, int D, std::vector<uint2> &out);
// SDR_ERR_UNSUPPORTED = the tensor path does not apply to this call (alignment / shared-memory budget)
int fir_tc_launch(const FirArgs &a, bool taps_complex, const uint2 *d_tables, float out_scale, cudaStream_t st);

// tcgen05 / TMEM Toeplitz FIR (+ Decimate) for u8 IQ input (fir_umma.cu).  R = samples between window rows,
// PC = output candidates per row (== R when D == 1).
int fir_umma_ksteps(int K, int R, int PC);
bool fir_umma_geometry(int K, int D, bool taps_complex, bool want_planar, int *R, int *PC, int *planar);
// mode: 0 interleaved bytes, 1 planar (real taps, opt-in), 2 polyphase planes (Decimate with D in 5..12)
bool fir_umma_build_tables(const float *taps, int K, bool taps_complex, int R, int PC, int mode, int D,
                           std::vector<uint8_t> &out, int magic[2][3], float sc[3]);
int fir_umma_launch(const FirArgs &a, int R, int PC, int mode, const uint8_t *d_tables, const int magic[2][3],
                    const float sc[3], cudaStream_t st);

// tcgen05 / TMEM Toeplitz FIR for c64 input, real taps, D == 1 (fir_umma_c64.cu): bf16 split of samples and taps,
// ns = 3 (six products, f32-grade accuracy) or 2 (three products, < 1e-5 of max|y|)
bool fir_umma_c64_applies(int K, int D, bool taps_complex, int ns);
bool fir_umma_c64_build_tables(const float *taps, int K, int ns, std::vector<uint8_t> &out);
// branch_tab_stride != 0: channel c uses the table set at d_tables + c * branch_tab_stride (per-channel taps; every CTA is
// bound to one channel, n_ch <= SM count)
int fir_umma_c64_launch(const FirArgs &a, int ns, const uint8_t *d_tables, cudaStream_t st, long long branch_tab_stride = 0);

// ---------------------------------------------------------------------------------------
// FFT (fft.cu)
// ---------------------------------------------------------------------------------------
struct FftArgs {
    const void *in;
    float2 *out;
    const float2 *tw;   // W_n^k, k in [0, n)  (device)
    long long batches;
    int log_n;          // n = 1 << log_n, 4 <= log_n <= 16
    int fmt;            // sdr_format_t
    unsigned flags;     // SDR_FFT_*
    float norm;         // 1.0f / sqrtf((float)n)
    int *work = nullptr; // n >= 2^14: batches + 1 ints of device scratch (ticket + per-transform counters), or null
};
int fft_pow2_launch(const FftArgs &a, cudaStream_t st);
// direct O(n^2) DFT for tiny / odd sizes handled without Bluestein (n <= 64)
int fft_naive_launch(const void *in, float2 *out, const float2 *tw, long long batches, int n, int fmt,
                     unsigned flags, float norm, cudaStream_t st);
// n = 2^17 .. 2^27, plain c64 -> c64, N = N1 * N2 through two n-element scratch buffers (five launches); tw1 / tw2 are
// the W tables of lengths 2^(log_n / 2) and 2^(log_n - log_n / 2)
int fft_huge_launch(const float2 *in, float2 *out, float2 *s1, float2 *s2, const float2 *tw1, const float2 *tw2,
                    int log_n, long long batches, int *work, cudaStream_t st);
// input format -> c64 ; c64 spectrum -> shift / norm / rfft selection of fft.rs:14-26,34-36
int fft_convert_launch(const void *in, float2 *a, long long total, int fmt, cudaStream_t st);
int fft_finish_launch(const float2 *a, float2 *out, long long batches, long long n, unsigned flags, float norm,
                      cudaStream_t st);
// Bluestein helpers: a[j] = x[j] * chirp[j] (zero padded to m), and the final pointwise stage
int bluestein_pre_launch(const void *in, float2 *a, const float2 *chirp, long long batches, int n, int m,
                         int fmt, cudaStream_t st);
int bluestein_mul_launch(float2 *a, const float2 *bfft, long long batches, int m, cudaStream_t st);
int bluestein_post_launch(const float2 *a, float2 *out, const float2 *chirp, long long batches, int n, int m,
                          unsigned flags, float norm, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// PLL (pll.cu)
// ---------------------------------------------------------------------------------------
struct PllParams {      // per stream, device resident
    float reference;    // already divided by rate (pll.rs:51)
    float gain, rate;
    float lc[5], oc[5], kc[5];  // loop / output / lock biquad coefficients b0,b1,b2,na1,na2
    int lk, ok, kk;             // 0 = Identity
};
struct PllState {       // per stream, device resident
    float nphase, vre, vim;
    float lx1r, lx1i, lx2r, lx2i, ly1r, ly1i, ly2r, ly2i;  // loop filter (complex)
    float ox1, ox2, oy1, oy2;                              // output filter
    float kx1, kx2, ky1, ky2;                              // lock filter
};
int pll_launch(const float2 *in, long long n, long long in_stride, float *out, uint8_t *locked,
               long long out_stride, const PllParams *params, int params_shared, PllState *state,
               int n_streams, bool fast_math, bool any_identity, cudaStream_t st);

// stand-alone biquad stream filter: n_seq real sequences (a complex stream is two), element i of sequence s at
// FM stereo pieces (src/main.rs:49,62-71,76-80): pilot Pll + (mono, diff) decode, demodulator output map, L/R matrix
int pll_stereo_launch(const float *in, long long n, long long in_stride, float *out_md, long long out_stride,
                      const PllParams *params, int params_shared, PllState *state, int n_streams, bool fast_math,
                      cudaStream_t st);
int fm_demod_map_launch(const float *v, const uint8_t *locked, float *out, long long n, cudaStream_t st);
int fm_matrix_launch(const float *md, long long md_stride, float *lr, long long lr_stride, int rows, long long n_frames,
                     cudaStream_t st);
// in[(s / W) * in_stride * W + i * W + s % W], W = 1 (f32) or 2 (c64); coef / state: 5 / 4 floats per sequence
int biquad_launch(const float *in, long long n, long long in_stride, float *out, long long out_stride, int W,
                  const float *coef, const int *kind, int coef_shared, float *state, int n_seq, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// resampler (resample.cu)
// ---------------------------------------------------------------------------------------
struct SrcLaunch {
    const float *v;     // device: carried frames followed by this call's input frames (interleaved)
    long long have;     // frames in v
    long long origin;   // frame index f (in the position's coordinate system) lives at v[f + origin]
    float *out;
    long long n_out;    // frames to produce
    int channels, type;
    double pos, step;   // output m sits at pos + m*step
    // sinc only
    const float *table; // device half table (half_len + 2 entries)
    long long half_len;
    double rq, rho;
    long long wc;
    // polyphase fast path (filled by src_launch when the positions are exact): device scratch for the per-phase wing
    // coefficients, 2 copies * 2 wings * 16 phases * (wc + 6) doubles, or null to force the per-tap kernel
    double *coef = nullptr;
};
int src_launch(const SrcLaunch &s, cudaStream_t st);
// Rational-ratio sinc conversion as a polyphase decimating FIR on the tensor cores (integer step S = 1/ratio, integer
// positions, 2 channels = one c64 stream): y[m] = sum_s FIR_{g_s}(x_s)[m], x_s[u] = v[S u + P + W - s], g_s[t] = g[S t + s],
// g[k] = rho * coef(|W - k|), W = wc + 1.  src_fast_taps fills the S branch filters (each Kb taps, zero padded); the FIR
// kernel gathers x_s itself (FirArgs::in_step).
int src_fast_branch_len(long long wc, int S);
void src_fast_taps(int type, double ratio, int S, std::vector<float> &branches);
// out[m] = ((b_0[m] + b_1[m]) + b_2[m]) + ... over the S branch outputs (rows of `pitch` frames)
int src_fast_sum(const float *branches, long long pitch, int S, float *out, long long n_out, cudaStream_t st);
// the windowed-sinc half table of converter `type` (0..2), computed once on the host in f64
size_t src_sinc_table_host(int type, const float **table, int *increment);
double src_sinc_wing(int type, double ratio, double *rq, double *rho, long long *wc);

}  // namespace sdr
