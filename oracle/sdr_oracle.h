/*
 * sdr_oracle.h -- CPU ORACLE for the sample-stream hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a from-scratch C++ restatement of the arithmetic of agrif/unnamed-rust-sdr
 * (reference tree, read-only).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libsdr_b200.so) never
 * links, loads or calls anything in this directory.
 *
 * Parity status of each block (see DESIGN.md "Oracle"):
 *   - unpack, FIR, decimate, take/skip/block counts, biquad, PLL, freq_sweep, fft
 *     post-processing: restated line by line from reference source that IS in the tree
 *     (citations on each function).  The reference holds no tests / golden vectors, so the
 *     pins are the known-answer tests derived from its semantics (tests/test_oracle.py).
 *   - FFT arithmetic (rustfft 3.0, Cargo.toml:18) and resampler arithmetic (C libsamplerate
 *     via libsamplerate-sys, Cargo.toml:24-26) live in dependencies that are NOT in the
 *     tree: PARITY UNPINNED at those two boundaries.  FFT is anchored on the mathematical
 *     DFT (f64) and numpy; the resampler follows libsamplerate's published algorithm
 *     structure with an own coefficient design (documented divergence).
 *
 * Build: oracle/Makefile  (g++ -O2 -ffp-contract=off: no FMA contraction, like rustc).
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sample kinds -------------------------------------------------------------------- */
#define ORC_KIND_F32 1 /* f32 samples            (Fir<f32,f32>)                          */
#define ORC_KIND_C64 2 /* Complex<f32> samples   (Fir<_,Complex<f32>>), interleaved re,im */

/* ---- a1: rtl_tcp unpack  (src/rtltcp.rs:158-164) --------------------------------------- */
void orc_unpack_u8iq(const uint8_t *iq, size_t n_samples, float *out_c64);

/* ---- a2: Fir<C,A>  (src/filter/fir.rs:7-33, src/filter/convolve.rs:13-15) --------------- */
typedef struct orc_fir orc_fir_t;
orc_fir_t *orc_fir_new(const float *taps, size_t n_taps, int taps_complex, int sample_kind);
void orc_fir_free(orc_fir_t *);
void orc_fir_reset(orc_fir_t *);
orc_fir_t *orc_fir_clone(const orc_fir_t *);
/* one Fir::apply per input element, in order; out has n elements of sample_kind */
void orc_fir_apply(orc_fir_t *, const float *in, size_t n, float *out);
/* f64 truth: zero history, same definition y[n] = sum_k c[k] x[n-k]; out is double */
void orc_fir_f64(const float *taps, size_t n_taps, int taps_complex, int sample_kind,
                 const float *in, size_t n, double *out);

/* ---- a4: Decimate  (src/signal/adapters/mod.rs:14-41) ---------------------------------- */
/* wait = (rate_in / rate_out).round() as usize, all in f32 (mod.rs:22) */
size_t orc_decimate_wait(float rate_in, float rate_out);
/* keeps in[(k+1)*wait-1]; *phase = number of inputs already discarded in the current group
 * (0 on a fresh stream), updated on return.  elem_floats = 1 (f32) or 2 (c64).  returns n_out */
size_t orc_decimate(const float *in, size_t n, size_t wait, int elem_floats, size_t *phase,
                    float *out);

/* ---- Take/Skip/Window/Block sample counts (adapters/mod.rs:174,249,279; block.rs:117) --- */
size_t orc_round_count(float rate, float duration); /* (rate*duration).round() as usize */
size_t orc_block_size(float size, float rate);      /* (size*rate).ceil() as usize      */
/* Times (src/signal/times.rs:17-21): t[i] = (i as f32)/rate */
void orc_times(float rate, size_t start, size_t n, float *out);

/* ---- a7: fft / rfft  (src/fft.rs:3-37) -------------------------------------------------- */
/* plain forward DFT, unnormalised, f32 arithmetic, twiddles from f64 (re-planned per call
 * like fft.rs:10-11).  Any n >= 1.  in/out interleaved c64. */
void orc_fft_f32(const float *in, size_t n, float *out);
/* the whole of fft.rs:3-28: labels[i] = (i - n/2) as f32 * (rate/n), vals = X[(i-n/2) mod n]*norm */
void orc_fft_shifted(const float *in, size_t n, float rate, float *labels, float *vals_c64);
/* rfft (fft.rs:30-37): real input, returns n - n/2 entries */
size_t orc_rfft_shifted(const float *in_real, size_t n, float rate, float *labels, float *vals_c64);
/* f64 truth DFT of f32 data (exact definition, any n) */
void orc_dft_f64(const float *in, size_t n, double *out);
/* batched helper used by the CPU baseline: `batches` consecutive n-point transforms of u8 IQ
 * (unpack + orc_fft_shifted values only), spread over `threads` std::threads. */
void orc_fft_batch_u8(const uint8_t *iq, size_t n, size_t batches, int threads, float *vals_c64);
void orc_fft_batch_c64(const float *in, size_t n, size_t batches, int shifted, int threads,
                       float *vals_c64);

/* ---- a9: Biquad / BiquadD  (src/filter/biquad.rs) ------------------------------------- */
#define ORC_BQ_IDENTITY 0 /* filter::Identity (src/filter/simple.rs:4-19) */
#define ORC_BQ_LOWPASS 1
#define ORC_BQ_HIGHPASS 2
#define ORC_BQ_BANDPASS 3
#define ORC_BQ_NOTCH 4
#define ORC_BQ_LR 5
/* coef[5] = b0,b1,b2,na1,na2 after Biquad::new's division by a0 (biquad.rs:25-38) */
void orc_biquad_design(int kind, float p0, float p1, float rate, float coef[5]);
typedef struct orc_biquad orc_biquad_t;
orc_biquad_t *orc_biquad_new(int kind, float p0, float p1, float rate, int sample_kind);
void orc_biquad_free(orc_biquad_t *);
void orc_biquad_apply(orc_biquad_t *, const float *in, size_t n, float *out);

/* ---- a8: Pll  (src/filter/pll.rs) -------------------------------------------------------- */
typedef struct {
    float reference, gain;          /* PllDesign::new(reference, gain, ..) pll.rs:26-36 */
    int loop_kind;  float loop_p0, loop_p1;     /* loopfilter   design (on Complex<f32>) */
    int out_kind;   float out_p0, out_p1;       /* outputfilter design (on f32)          */
    int lock_kind;  float lock_p0, lock_p1;     /* lockfilter   design (on f32)          */
} orc_pll_design_t;
typedef struct orc_pll orc_pll_t;
orc_pll_t *orc_pll_new(const orc_pll_design_t *, float rate);
void orc_pll_free(orc_pll_t *);
/* out[i] = output value, locked[i] = 1 iff Some(..) (locked > 0.01) */
void orc_pll_apply(orc_pll_t *, const float *in_c64, size_t n, float *out, uint8_t *locked);
void orc_pll_state(const orc_pll_t *, float *nphase, float *value_re, float *value_im);
/* the stereo-decode closure of src/main.rs:62-71 around a pilot Pll: out = (mono, diff) frames */
void orc_fm_stereo_decode(orc_pll_t *, const float *v, size_t n, float *out_mono_diff);

/* ---- FreqSweep  (src/signal/sources.rs:116-194) ---------------------------------------- */
/* freq_sweep(rate, df, warmup, start..end); writes up to cap samples, returns total length */
size_t orc_freq_sweep(float rate, float df, int warmup, float range_start, float range_end,
                      float *out_freq, float *out_c64, size_t cap);

/* ---- a5/a6: resample ------------------------------------------------------------------- */
/* libsamplerate-shaped API (src/resample.rs binds src_new/src_process/... :5,36,61,73,81,89,95,105).
 * Arithmetic = own "sdr-src" specification; PARITY UNPINNED. */
#define ORC_SRC_SINC_BEST_QUALITY 0
#define ORC_SRC_SINC_MEDIUM_QUALITY 1
#define ORC_SRC_SINC_FASTEST 2
#define ORC_SRC_ZERO_ORDER_HOLD 3
#define ORC_SRC_LINEAR 4
typedef struct {
    const float *data_in;
    float *data_out;
    long input_frames, output_frames;
    long input_frames_used, output_frames_gen;
    int end_of_input;
    double src_ratio;
} orc_src_data_t;
typedef struct orc_src orc_src_t;
orc_src_t *orc_src_new(int converter_type, int channels, int *error);
orc_src_t *orc_src_delete(orc_src_t *);
int orc_src_process(orc_src_t *, orc_src_data_t *);
int orc_src_reset(orc_src_t *);
orc_src_t *orc_src_clone(const orc_src_t *, int *error);
int orc_src_set_ratio(orc_src_t *, double ratio);
int orc_src_get_channels(const orc_src_t *);
long orc_src_history_frames(const orc_src_t *);
const char *orc_src_strerror(int error);
/* the coefficient design shared with the product (pure function of the type):
 * returns half-length (number of table entries), *increment = entries per zero crossing */
size_t orc_src_sinc_table(int converter_type, const float **table, int *increment);
/* signal::Resample adaptor loop (src/signal/adapters/resample.rs:38-82) over a finite input:
 * 4096-frame input buffer, 4096-frame output capacity.  returns frames written (<= cap) */
size_t orc_resample_signal(const float *in, size_t n_frames, int channels, int converter_type,
                           double ratio, float *out, size_t cap);

/* ---- CPU-baseline helpers (multi-threaded over independent units; used by bench.py) ---- */
/* FIR over a u8 IQ capture: unpack + Fir::apply per sample + Decimate(wait); the stream is cut
 * into `threads` contiguous ranges with a (n_taps-1) halo so every output is identical to the
 * single-thread result.  returns n_out */
size_t orc_fir_u8_mt(const uint8_t *iq, size_t n, const float *taps, size_t n_taps,
                     int taps_complex, size_t wait, int threads, float *out_c64);
/* channelizer: n_ch independent streams (channel-major c64, n per channel):
 * Fir(taps) -> Pll(design) per channel */
void orc_channelizer_mt(const float *in_c64, size_t n_ch, size_t n, const float *taps,
                        size_t n_taps, const orc_pll_design_t *, float rate, int threads,
                        float *out, uint8_t *locked);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
