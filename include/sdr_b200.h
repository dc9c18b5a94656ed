/*
 * sdr_b200.h -- C ABI of libsdr_b200.so: the B200 (sm_100a) implementation of the
 * sample-stream hot path of agrif/unnamed-rust-sdr.
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface
 * (file:line in the reference tree) it replaces; INTEGRATION.md shows the Rust `extern "C"`
 * block + safe wrappers a maintainer would add.  Conventions (modelled on the reference's one
 * existing FFI, src/resample.rs:33-110 over libsamplerate):
 *   - opaque handles; constructors return NULL and write an error code through `int *err`
 *   - every other call returns an int status, 0 = OK (see sdr_strerror)
 *   - buffers are caller-owned, plain pointers + element counts; nothing is returned allocated
 *   - `*_process` / `*_exec` take HOST pointers (any alignment) and are synchronous
 *   - `*_dev` variants take DEVICE pointers on the handle's device (16-byte aligned) and are
 *     asynchronous on the handle's stream
 *   - a handle is bound to one device + one stream; it may migrate between host threads but
 *     must not be used from two threads at once (Send, not Sync -- resample.rs:25)
 *   - there is NO CPU fallback: without a CUDA device constructors fail with SDR_ERR_NO_DEVICE
 *
 * Sample formats:
 *   SDR_FMT_U8IQ  interleaved u8 I,Q as sent by rtl_tcp; unpacked on the fly exactly as
 *                 RtlTcpSignal::next does: (b as f32 - 128.0) / 128.0   (src/rtltcp.rs:158-164)
 *   SDR_FMT_C64   num::Complex<f32>, interleaved re,im (8 bytes)
 *   SDR_FMT_F32   f32 real samples
 */
#ifndef SDR_B200_H
#define SDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDR_B200_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------- */
/* 1..22 keep libsamplerate's numbering because src/resample.rs:209-269 hard-codes that map. */
#define SDR_OK 0
#define SDR_ERR_MALLOC_FAILED 1
#define SDR_ERR_BAD_STATE 2
#define SDR_ERR_BAD_DATA 3
#define SDR_ERR_BAD_DATA_PTR 4
#define SDR_ERR_BAD_SRC_RATIO 6
#define SDR_ERR_BAD_CONVERTER 10
#define SDR_ERR_BAD_CHANNEL_COUNT 11
#define SDR_ERR_DATA_OVERLAP 16
#define SDR_ERR_INVALID_ARG 100
#define SDR_ERR_UNSUPPORTED 101
#define SDR_ERR_NO_DEVICE 102
#define SDR_ERR_NULL_HANDLE 103
#define SDR_ERR_OUTPUT_TOO_SMALL 104
#define SDR_ERR_MISALIGNED 105
#define SDR_ERR_CUDA_BASE 1000 /* 1000 + cudaError_t */

const char *sdr_strerror(int code);
int sdr_abi_version(void);
/* number of CUDA devices visible (0 if none / no driver) */
int sdr_device_count(void);
/* name, SM count and memory of device `dev` as a short text line */
int sdr_device_info(int dev, char *buf, size_t cap);

typedef enum { SDR_FMT_U8IQ = 0, SDR_FMT_C64 = 1, SDR_FMT_F32 = 2 } sdr_format_t;

/* ======================================================================================
 * unpack -- RtlTcpSignal::next  (src/rtltcp.rs:158-164).  bit-exact.
 * ====================================================================================== */
int sdr_unpack_u8iq(const uint8_t *iq, size_t n_samples, float *out_c64, int device);
int sdr_unpack_u8iq_dev(const uint8_t *iq, size_t n_samples, float *out_c64, int device, void *stream);

/* ======================================================================================
 * FIR -- filter::Fir<C,A> (src/filter/fir.rs:7-33) driven through signal::Filter
 * (src/signal/adapters/mod.rs:67-100) and, when decimation > 1, signal::Decimate
 * (src/signal/adapters/mod.rs:14-41) fused behind it so that only kept outputs are computed.
 *
 *   y[n] = sum_{k<K} taps[k] * x[n-k],  x[<0] = 0            (fir.rs:23-32)
 *   kept outputs: n = (j+1)*D - 1 over the whole stream      (adapters/mod.rs:30-37)
 * The handle carries the last K-1 inputs and the decimation phase across calls, so feeding a
 * stream in blocks of any size gives the same samples as one call.
 * ====================================================================================== */
/* Default (flags = 0) for u8 IQ input is the integer tcgen05 path: taps are quantised to 24-bit fixed point relative to
 * the largest |tap| (a tap below 2^-24 of the largest contributes nothing; error < 1e-6 of max|y|).
 * STRICT_ORDER selects the reference-order CUDA-core kernel instead. */
#define SDR_FIR_STRICT_ORDER 1u /* f32 mul then add, k ascending, no FMA: bit-identical to Fir::apply */
#define SDR_FIR_NO_TENSOR 2u    /* never take a tensor-core path */
#define SDR_FIR_NO_TCGEN05 4u   /* never take the tcgen05/TMEM path (the mma.sync Toeplitz path may still run) */
#define SDR_FIR_SPLIT2 16u      /* c64 input on tcgen05: two bf16 terms per operand (3 products, error < 1e-5 of max|y|) instead of
                                  * three (6 products, the reference's f32 accuracy); ~1.5x less tensor work */
#define SDR_FIR_PLANAR 8u       /* real taps: tcgen05 kernel on de-interleaved I / Q byte planes (same bits; see fir_umma.cu) */

typedef struct {
    const float *taps;   /* n_taps f32, or n_taps (re,im) pairs when taps_complex */
    size_t n_taps;       /* K >= 1 */
    int taps_complex;    /* 0: Fir<f32,A>   1: Fir<Complex<f32>,Complex<f32>> */
    int input_format;    /* sdr_format_t; output is C64 for U8IQ/C64 input, F32 for F32 input */
    size_t decimation;   /* Decimate `wait` D >= 1 (0 is rejected: the reference underflows) */
    size_t n_channels;   /* independent streams sharing the taps (>= 1) */
    unsigned flags;
    int device;
    void *stream;        /* cudaStream_t to run on; NULL = handle creates its own (pass cudaStreamLegacy,
                            (void*)0x1, to run on the legacy default stream) */
} sdr_fir_config_t;

typedef struct sdr_fir sdr_fir_t;
sdr_fir_t *sdr_fir_create(const sdr_fir_config_t *cfg, int *err);
void sdr_fir_destroy(sdr_fir_t *);
int sdr_fir_reset(sdr_fir_t *);                        /* zero history and phase (Fir::new state) */
sdr_fir_t *sdr_fir_clone(const sdr_fir_t *, int *err); /* #[derive(Clone)] on Fir, fir.rs:6 */
/* number of outputs a call with n_in new inputs per channel will produce, given the current phase */
size_t sdr_fir_output_count(const sdr_fir_t *, size_t n_in);
/* in: n_channels rows of n_in elements, row stride in_stride elements (ignored if 1 channel)
 * out: n_channels rows, row stride out_stride elements; out_cap = elements available per row.
 * *n_used = inputs consumed per channel (== n_in on success), *n_out = outputs per channel. */
int sdr_fir_process(sdr_fir_t *, const void *in, size_t n_in, size_t in_stride, void *out,
                    size_t out_cap, size_t out_stride, size_t *n_used, size_t *n_out);
int sdr_fir_process_dev(sdr_fir_t *, const void *in, size_t n_in, size_t in_stride, void *out,
                        size_t out_cap, size_t out_stride, size_t *n_used, size_t *n_out);
/* which kernel family the last process call used: 0 none, 1 CUDA-core direct, 2 CUDA-core
 * strict-order, 3 tensor-core Toeplitz (mma.sync), 4 tensor-core Toeplitz (tcgen05 / TMEM, integer, u8 IQ input),
 * 5 tensor-core Toeplitz (tcgen05 / TMEM, bf16-split operands, c64 input with real taps) */
int sdr_fir_last_path(const sdr_fir_t *);

/* Decimate::new's `wait` = (rate_in / rate_out).round() as usize in f32 (adapters/mod.rs:22) */
size_t sdr_decimate_wait(float rate_in, float rate_out);
/* Take/Skip/Window: (rate * duration).round() as usize (adapters/mod.rs:174,249,279) */
size_t sdr_duration_samples(float rate, float duration);
/* Block: (size * rate).ceil() as usize (adapters/block.rs:117) */
size_t sdr_block_samples(float size, float rate);

/* ======================================================================================
 * FFT -- fft::fft / fft::rfft (src/fft.rs:3-37); the arithmetic replaces rustfft's
 * FFTplanner::plan_fft + process (fft.rs:10-12).
 *   X[k] = sum_n x[n] e^{-2 pi i nk/N}                               (unnormalised forward)
 *   SDR_FFT_SHIFT: out[i] = X[(i - N/2) mod N]                       (fft.rs:15,18-25)
 *   SDR_FFT_NORM : out *= 1.0/(N as f32).sqrt(), multiplied in f32   (fft.rs:16,25)
 *   SDR_FFT_RFFT : input is F32 real, output keeps entries N/2..N-1 of the shifted spectrum
 *                  (rfft's drain(0..len/2), fft.rs:34-36); implies SHIFT
 * A plan is immutable after creation; `batches` consecutive N-point blocks per call.
 * ====================================================================================== */
#define SDR_FFT_SHIFT 1u
#define SDR_FFT_NORM 2u
#define SDR_FFT_RFFT 4u

typedef struct {
    size_t n;          /* transform length: any n >= 1 up to 2^27 (powers of two) or 2^26 (any other length,
                        * Bluestein); SDR_ERR_UNSUPPORTED beyond.  fft::fft takes a whole finite signal */
    int input_format;  /* U8IQ, C64, or F32 (with SDR_FFT_RFFT) */
    unsigned flags;
    int device;
    void *stream;
} sdr_fft_config_t;

typedef struct sdr_fft sdr_fft_t;
sdr_fft_t *sdr_fft_create(const sdr_fft_config_t *cfg, int *err);
void sdr_fft_destroy(sdr_fft_t *);
/* output elements (c64) per transform: n, or n - n/2 with SDR_FFT_RFFT */
size_t sdr_fft_output_len(const sdr_fft_t *);
int sdr_fft_exec(sdr_fft_t *, const void *in, size_t batches, float *out_c64);
int sdr_fft_exec_dev(sdr_fft_t *, const void *in, size_t batches, float *out_c64);
/* the frequency labels of fft.rs:14-24: labels[i] = (i - n/2) as f32 * (rate / n as f32);
 * with rfft != 0 only the kept n - n/2 entries are written.  Host-side, pure. */
int sdr_fft_labels(size_t n, float rate, int rfft, float *labels);

/* ======================================================================================
 * PLL -- filter::Pll / PllDesign (src/filter/pll.rs:4-85) with Biquad / BiquadD / Identity
 * sub-filters (src/filter/biquad.rs:5-154, src/filter/simple.rs:4-19).  One GPU lane per
 * independent stream (the loop itself is sequential).
 * ====================================================================================== */
typedef enum {
    SDR_BQ_IDENTITY = 0, /* filter::Identity */
    SDR_BQ_LOWPASS = 1,  /* BiquadD::LowPass(freq, q)  */
    SDR_BQ_HIGHPASS = 2, /* BiquadD::HighPass(freq, q) */
    SDR_BQ_BANDPASS = 3, /* BiquadD::BandPass(freq, q) */
    SDR_BQ_NOTCH = 4,    /* BiquadD::Notch(freq, q)    */
    SDR_BQ_LR = 5        /* BiquadD::Lr(decayrate)     */
} sdr_biquad_kind_t;

typedef struct { int kind; float p0, p1; } sdr_biquad_design_t;

/* PllDesign::new(reference, gain, loopfilter, outputfilter, lockfilter)  pll.rs:26-36 */
typedef struct {
    float reference, gain;
    sdr_biquad_design_t loopfilter, outputfilter, lockfilter;
} sdr_pll_design_t;

/* BiquadD::design + Biquad::new (biquad.rs:25-38,83-154): coef = b0,b1,b2,na1,na2.  Host, pure. */
int sdr_biquad_design(const sdr_biquad_design_t *d, float rate, float coef[5]);

/* atan2 / sin / cos of the PLL recurrence (pll.rs:72,76).  Default (flags = 0): f32 routines within ~1 ulp of libm's
 * atan2f / sinf / cosf, arranged for the depth of the dependent chain (219 cycles per sample).  SDR_PLL_F64_MATH: the
 * same functions evaluated in f64 and rounded to f32 (equal to libm's result for all but ~1e-4 of the arguments), 455
 * cycles per sample.  Both meet the same parity bars against the CPU oracle.  SDR_PLL_FAST_MATH is
 * accepted for source compatibility and has no effect (it selected the f32 routines when f64 was the default). */
#define SDR_PLL_FAST_MATH 1u
#define SDR_PLL_F64_MATH 2u
/* always run the general kernel (per-sample filter-kind tests, fract for any step size) instead of the one specialised
 * for designs without Identity sub-filters and with |reference / rate| + pi |gain| < 1.  Same results bit for bit where
 * both apply (tests/test_gpu_pll_resample.py); for testing. */
#define SDR_PLL_GENERAL_KERNEL 4u

typedef struct {
    const sdr_pll_design_t *designs; /* n_designs entries */
    size_t n_designs;                /* 1 (shared by all streams) or n_streams */
    size_t n_streams;
    float rate;                      /* FilterDesign::design(rate)  pll.rs:48 */
    unsigned flags;
    int device;
    void *stream;
} sdr_pll_config_t;

typedef struct sdr_pll sdr_pll_t;
sdr_pll_t *sdr_pll_create(const sdr_pll_config_t *cfg, int *err);
void sdr_pll_destroy(sdr_pll_t *);
int sdr_pll_reset(sdr_pll_t *);
sdr_pll_t *sdr_pll_clone(const sdr_pll_t *, int *err); /* #[derive(Clone)] pll.rs:12 */
/* in: n_streams rows of n c64 samples (row stride in_stride elements);
 * out: f32 value, locked: 1 iff the reference returns Some(..) (locked > 0.01, pll.rs:80-84) */
int sdr_pll_process(sdr_pll_t *, const float *in_c64, size_t n, size_t in_stride, float *out,
                    uint8_t *locked, size_t out_stride);
int sdr_pll_process_dev(sdr_pll_t *, const float *in_c64, size_t n, size_t in_stride, float *out,
                        uint8_t *locked, size_t out_stride);
/* FM stereo decode around a pilot-tone Pll -- the closure of src/main.rs:62-71:
 *     mono = v * 0.5;  diff = Some(_) = pll.apply(Complex::new(v, 0.0)) ? (v / pll.value.powi(2)).re * 0.5 : 0.0
 * v: n_streams rows of n f32 samples (row stride in_stride samples); out_mono_diff: rows of n
 * (mono, diff) frames (row stride out_stride frames).  Advances the same carried state as process. */
int sdr_pll_stereo_decode(sdr_pll_t *, const float *v, size_t n, size_t in_stride, float *out_mono_diff,
                          size_t out_stride);
int sdr_pll_stereo_decode_dev(sdr_pll_t *, const float *v, size_t n, size_t in_stride, float *out_mono_diff,
                              size_t out_stride);
/* the public fields Pll::nphase / Pll::value (pll.rs:21-22) of one stream; synchronises */
int sdr_pll_get_state(sdr_pll_t *, size_t stream_index, float *nphase, float *value_re, float *value_im);

/* ======================================================================================
 * Biquad as a stream filter -- Filter::apply of filter::Biquad<f32, A> (src/filter/biquad.rs:40-56)
 * driven by signal::Filter (src/signal/adapters/mod.rs:94-96), e.g. the Lr de-emphasis filters of
 * src/main.rs:52,75-80.  A = f32 or Complex<f32> (both parts filtered independently with the
 * same real coefficients).  n_streams independent streams, one GPU lane per real sequence; the
 * operation order of biquad.rs:44-49 is kept (no FMA), so results are bit-identical.
 * ====================================================================================== */
typedef struct {
    const sdr_biquad_design_t *designs; /* n_designs entries */
    size_t n_designs;                   /* 1 (shared by all streams) or n_streams */
    size_t n_streams;
    float rate;                         /* FilterDesign::design(rate), biquad.rs:83 */
    int sample_complex;                 /* 0: Biquad<f32,f32>, 1: Biquad<f32,Complex<f32>> */
    int device;
    void *stream;
} sdr_biquad_config_t;

typedef struct sdr_biquad sdr_biquad_t;
sdr_biquad_t *sdr_biquad_create(const sdr_biquad_config_t *cfg, int *err);
void sdr_biquad_destroy(sdr_biquad_t *);
int sdr_biquad_reset(sdr_biquad_t *);                        /* zero x1, x2, y1, y2 (Biquad::new, biquad.rs:25-38) */
sdr_biquad_t *sdr_biquad_clone(const sdr_biquad_t *, int *err); /* #[derive(Clone)] biquad.rs:4 */
/* in / out: n_streams rows of n samples (f32 or c64), row strides in samples (ignored for one stream) */
int sdr_biquad_process(sdr_biquad_t *, const float *in, size_t n, size_t in_stride, float *out, size_t out_stride);
int sdr_biquad_process_dev(sdr_biquad_t *, const float *in, size_t n, size_t in_stride, float *out, size_t out_stride);

/* ======================================================================================
 * channelizer -- n_channels x ( Fir<f32,Complex<f32>> -> Pll ): config C4.  The FIR output
 * feeds the PLL on chip; only the PLL output leaves.  Same carried state as the two parts.
 * ====================================================================================== */
typedef struct sdr_channelizer sdr_channelizer_t;
sdr_channelizer_t *sdr_channelizer_create(const sdr_fir_config_t *fir, const sdr_pll_config_t *pll, int *err);
void sdr_channelizer_destroy(sdr_channelizer_t *);
int sdr_channelizer_reset(sdr_channelizer_t *);
int sdr_channelizer_process(sdr_channelizer_t *, const void *in, size_t n, size_t in_stride, float *out,
                            uint8_t *locked, size_t out_stride);
int sdr_channelizer_process_dev(sdr_channelizer_t *, const void *in, size_t n, size_t in_stride,
                                float *out, uint8_t *locked, size_t out_stride);

/* ======================================================================================
 * resample -- resample::SampleRate<A> (src/resample.rs:11-110) binds libsamplerate's
 * src_new / src_process / src_reset / src_clone / src_delete / src_set_ratio /
 * src_get_channels / src_strerror (resample.rs:36,61,73,81,105,95,89,196).  The functions
 * below have the same signatures and struct layout, so resample.rs only changes its `use`.
 * Arithmetic: the "sdr-src" specification in DESIGN.md (libsamplerate itself is not in the
 * reference tree: parity unpinned).
 * ====================================================================================== */
enum {
    SDR_SRC_SINC_BEST_QUALITY = 0,
    SDR_SRC_SINC_MEDIUM_QUALITY = 1,
    SDR_SRC_SINC_FASTEST = 2,
    SDR_SRC_ZERO_ORDER_HOLD = 3,
    SDR_SRC_LINEAR = 4
};

typedef struct { /* == libsamplerate SRC_DATA, built at resample.rs:49-58 */
    const float *data_in;
    float *data_out;
    long input_frames, output_frames;
    long input_frames_used, output_frames_gen;
    int end_of_input;
    double src_ratio;
} SDR_SRC_DATA;

typedef struct sdr_src SDR_SRC_STATE;
SDR_SRC_STATE *sdr_src_new(int converter_type, int channels, int *error);
/* same, choosing device/stream (sdr_src_new uses device 0 and its own stream) */
SDR_SRC_STATE *sdr_src_new_on(int converter_type, int channels, int device, void *stream, int *error);
SDR_SRC_STATE *sdr_src_delete(SDR_SRC_STATE *);
int sdr_src_process(SDR_SRC_STATE *, SDR_SRC_DATA *);
/* data_in / data_out are device pointers; counts are still returned synchronously (they are
 * computed on the host), the samples land asynchronously on the handle's stream */
int sdr_src_process_dev(SDR_SRC_STATE *, SDR_SRC_DATA *);
int sdr_src_reset(SDR_SRC_STATE *);
SDR_SRC_STATE *sdr_src_clone(SDR_SRC_STATE *, int *error);
int sdr_src_set_ratio(SDR_SRC_STATE *, double new_ratio);
int sdr_src_get_channels(SDR_SRC_STATE *);
/* frames of input history the handle carries between calls (sinc converters: bounded by the filter
 * wing + what the last call consumed; a call that stops at `output_frames` hands the input it did
 * not need back through input_frames_used, as libsamplerate's src_process does) */
long sdr_src_history_frames(SDR_SRC_STATE *);
/* Sinc converters, long calls (>= 8192 output frames) at an integer step 1/ratio on 2-channel (c64) data run as a
 * polyphase decimating FIR on the tensor cores: within 1e-5 of max|y| of the converter's f64 specification.
 * exact != 0 keeps the f64 kernels for every call (equal to the specification up to the final f32 rounding). */
int sdr_src_set_exact(SDR_SRC_STATE *, int exact);
const char *sdr_src_strerror(int error);
const char *sdr_src_get_name(int converter_type);        /* resample.rs:125 */
const char *sdr_src_get_description(int converter_type); /* resample.rs:133 */
const char *sdr_src_get_version(void);                   /* resample.rs:5 */
/* the windowed-sinc half table of a sinc converter (host pointer, half_len + 2 entries) */
size_t sdr_src_sinc_table(int converter_type, const float **table, int *increment);

/* ======================================================================================
 * stream timing helper for benchmarks: CUDA-event time of everything queued on the handle's
 * stream between begin and end (ms).  Pure measurement, no effect on results.
 * ====================================================================================== */
typedef struct sdr_timer sdr_timer_t;
sdr_timer_t *sdr_timer_create(int device, void *stream, int *err);
void sdr_timer_destroy(sdr_timer_t *);
int sdr_timer_begin(sdr_timer_t *);
int sdr_timer_end(sdr_timer_t *, float *elapsed_ms); /* synchronises the stream */
/* count of kernels this library has launched in this process (all handles) */
uint64_t sdr_kernel_launch_count(void);

/* ======================================================================================
 * FM broadcast stereo receiver -- the signal chain of src/main.rs:32-81 for a batch of
 * stations, every intermediate device-resident:
 *   u8 IQ (rtltcp.rs:158-164) -> Pll demodulator (main.rs:41-49) -> / 75000 -> resample_with(
 *   SincFastest, 144 kHz) (main.rs:50) -> pilot Pll + (mono, diff) decode (main.rs:54-71) ->
 *   resample(48 kHz) (main.rs:73) -> Lr de-emphasis + (mono+diff, mono-diff) (main.rs:52,75-80).
 * Carried state: both PLLs, both converters and the de-emphasis filters of every station, so a
 * stream may be fed in calls of any size.
 * ====================================================================================== */
typedef struct {
    size_t n_stations;
    float rate;     /* input sample rate; main.rs:32 uses 1 800 000 */
    float pilot;    /* pilot tone in Hz; 0 selects main.rs:54's 19 000 */
    unsigned flags; /* SDR_PLL_F64_MATH */
    int device;
    void *stream;
} sdr_fm_config_t;

typedef struct sdr_fm sdr_fm_t;
sdr_fm_t *sdr_fm_create(const sdr_fm_config_t *cfg, int *err);
void sdr_fm_destroy(sdr_fm_t *);
int sdr_fm_reset(sdr_fm_t *);
float sdr_fm_output_rate(const sdr_fm_t *);          /* 48000 */
/* the out_cap sdr_fm_process[_dev] requires for n input samples: a smaller one is rejected with
 * SDR_ERR_OUTPUT_TOO_SMALL before any state changes (the call can be retried with a larger buffer) */
size_t sdr_fm_max_output(const sdr_fm_t *, size_t n);
/* iq: n_stations rows of n u8 IQ samples (2n bytes each, row stride in_stride BYTES);
 * out: n_stations rows of (left, right) f32 frames at 48 kHz, capacity out_cap frames per row,
 * row stride out_stride frames; *n_out = frames written per row (the same for every station).
 * end_of_input != 0 flushes the converters the way signal::Resample does at the end of its
 * upstream (adapters/resample.rs:45-65).  _dev: iq / out are device pointers, rows 16-byte
 * aligned (in_stride % 16 == 0 when n_stations > 1), samples land asynchronously on the stream. */
int sdr_fm_process(sdr_fm_t *, const uint8_t *iq, size_t n, size_t in_stride, float *out, size_t out_cap,
                   size_t out_stride, size_t *n_out, int end_of_input);
int sdr_fm_process_dev(sdr_fm_t *, const uint8_t *iq, size_t n, size_t in_stride, float *out, size_t out_cap,
                       size_t out_stride, size_t *n_out, int end_of_input);

/* ======================================================================================
 * sliding-window spectra -- the spectrum path of examples/live.rs:30-39:
 *     sig.window(duration).decimate(fps).map(|w| fft::fft(from_iter(rate, w...)))
 * Window (src/signal/adapters/mod.rs:271-303) holds the last `window` samples (initially
 * zeros) and yields after every input sample; Decimate (mod.rs:14-41) keeps every hop-th, so
 * kept window j ends at input sample (j + 1) * hop - 1; each kept window goes through
 * fft::fft (src/fft.rs:3-28; flags as sdr_fft_config_t, any length).
 *   window = sdr_duration_samples(rate, duration),  hop = sdr_decimate_wait(rate, fps)
 * Carried state: the last window - 1 samples and the stream position.
 * ====================================================================================== */
typedef struct {
    size_t window;    /* samples per window = transform length */
    size_t hop;       /* Decimate's wait D >= 1 */
    int input_format; /* U8IQ or C64 */
    unsigned flags;   /* SDR_FFT_SHIFT | SDR_FFT_NORM */
    int device;
    void *stream;
} sdr_window_fft_config_t;

typedef struct sdr_window_fft sdr_window_fft_t;
sdr_window_fft_t *sdr_window_fft_create(const sdr_window_fft_config_t *cfg, int *err);
void sdr_window_fft_destroy(sdr_window_fft_t *);
int sdr_window_fft_reset(sdr_window_fft_t *);
size_t sdr_window_fft_size(const sdr_window_fft_t *);
/* windows the next n_in input samples complete */
size_t sdr_window_fft_output_count(const sdr_window_fft_t *, size_t n_in);
/* out_c64: *n_windows consecutive spectra of `window` c64 values each (capacity out_cap windows) */
int sdr_window_fft_process(sdr_window_fft_t *, const void *in, size_t n_in, float *out_c64, size_t out_cap,
                           size_t *n_windows);
int sdr_window_fft_process_dev(sdr_window_fft_t *, const void *in, size_t n_in, float *out_c64, size_t out_cap,
                               size_t *n_windows);

#ifdef __cplusplus
}
#endif
#endif /* SDR_B200_H */
