//! gpu.rs -- safe wrappers a maintainer would add to the crate (`mod gpu;` in src/lib.rs).
//! NOT COMPILED HERE (no Rust toolchain in this image); kept deliberately thin so that everything with
//! behaviour lives behind the C ABI, where it is tested from C++ and Python.
//!
//! What changes in the reference tree:
//!   * src/resample.rs:1      `use libsamplerate_sys::*;`  ->  `use crate::sdr_b200_sys::{sdr_src_new as src_new,
//!                             sdr_src_process as src_process, sdr_src_reset as src_reset, sdr_src_clone as src_clone,
//!                             sdr_src_delete as src_delete, sdr_src_set_ratio as src_set_ratio,
//!                             sdr_src_get_channels as src_get_channels, sdr_src_strerror as src_strerror,
//!                             sdr_src_get_name as src_get_name, sdr_src_get_description as src_get_description,
//!                             sdr_src_get_version as src_get_version, SDR_SRC_STATE as SRC_STATE, SDR_SRC_DATA as SRC_DATA};`
//!                             (+ the five SRC_* converter constants).  Nothing else in resample.rs changes.
//!   * src/fft.rs:10-12       the rustfft planner + process become `gpu::fft_shifted(&data, rate)`.
//!   * src/signal/adapters/mod.rs  `Filter<S, Fir<..>>` gets a block-buffered `next()` (below), shaped exactly like
//!                             signal::Resample::next (src/signal/adapters/resample.rs:38-82).
use crate::sdr_b200_sys::*;
use crate::Signal;
use num::Complex;
use std::os::raw::c_void;

#[derive(Debug)]
pub struct Error(pub i32);
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        let s = unsafe { std::ffi::CStr::from_ptr(sdr_strerror(self.0)) };
        write!(f, "{}", s.to_string_lossy())
    }
}
fn check(rc: i32) -> Result<(), Error> { if rc == SDR_OK { Ok(()) } else { Err(Error(rc)) } }

/// sealed: sample / coefficient types the GPU path accepts (there is no CPU fallback)
pub trait GpuSample: Copy { const FMT: i32; }
impl GpuSample for f32 { const FMT: i32 = SDR_FMT_F32; }
impl GpuSample for Complex<f32> { const FMT: i32 = SDR_FMT_C64; }
pub trait GpuCoef: Copy { const COMPLEX: i32; }
impl GpuCoef for f32 { const COMPLEX: i32 = 0; }
impl GpuCoef for Complex<f32> { const COMPLEX: i32 = 1; }

/// filter::Fir<C, A> backed by the GPU: same constructor, plus a block `process` mirroring
/// resample::SampleRate::process (src/resample.rs:46-67).
pub struct GpuFir<C: GpuCoef, A: GpuSample> {
    h: *mut sdr_fir_t,
    _m: std::marker::PhantomData<(C, fn(A) -> A)>,
}
unsafe impl<C: GpuCoef, A: GpuSample> Send for GpuFir<C, A> {}

impl<C: GpuCoef, A: GpuSample> GpuFir<C, A> {
    pub fn new(coef: Vec<C>) -> Result<Self, Error> { Self::with_decimation(coef, 1, false) }
    pub fn with_decimation(coef: Vec<C>, wait: usize, input_u8iq: bool) -> Result<Self, Error> {
        let cfg = sdr_fir_config_t {
            taps: coef.as_ptr() as *const f32, n_taps: coef.len(), taps_complex: C::COMPLEX,
            input_format: if input_u8iq { SDR_FMT_U8IQ } else { A::FMT },
            decimation: wait, n_channels: 1, flags: 0, device: 0, stream: std::ptr::null_mut(),
        };
        let mut err = 0;
        let h = unsafe { sdr_fir_create(&cfg, &mut err) };
        if h.is_null() { Err(Error(err)) } else { Ok(GpuFir { h, _m: std::marker::PhantomData }) }
    }
    /// output gets one element per kept input; returns inputs consumed
    pub fn process(&mut self, input: &[A], output: &mut Vec<A>) -> Result<usize, Error> {
        let n_out = unsafe { sdr_fir_output_count(self.h, input.len()) };
        output.clear();
        output.reserve(n_out);
        let (mut used, mut got) = (0usize, 0usize);
        check(unsafe {
            sdr_fir_process(self.h, input.as_ptr() as *const c_void, input.len(), input.len(),
                            output.as_mut_ptr() as *mut c_void, n_out, n_out, &mut used, &mut got)
        })?;
        unsafe { output.set_len(got) };
        Ok(used)
    }
    pub fn reset(&mut self) -> Result<(), Error> { check(unsafe { sdr_fir_reset(self.h) }) }
}
impl<C: GpuCoef, A: GpuSample> Clone for GpuFir<C, A> {
    fn clone(&self) -> Self {
        let mut err = 0;
        let h = unsafe { sdr_fir_clone(self.h, &mut err) };
        assert!(!h.is_null(), "sdr_fir_clone failed: {}", Error(err));
        GpuFir { h, _m: std::marker::PhantomData }
    }
}
impl<C: GpuCoef, A: GpuSample> Drop for GpuFir<C, A> {
    fn drop(&mut self) { unsafe { sdr_fir_destroy(self.h) } }
}

/// Block-buffered signal::Filter for a GPU FIR: pulls `block` samples upstream, one launch, serves one by one.
/// Same shape as signal::Resample::next (src/signal/adapters/resample.rs:38-82).
pub struct GpuFilter<S: Signal, C: GpuCoef> where S::Sample: GpuSample {
    signal: S,
    fir: GpuFir<C, S::Sample>,
    block: usize,
    input: Vec<S::Sample>,
    output: Vec<S::Sample>,
    next: usize,
}
impl<S: Signal, C: GpuCoef> Signal for GpuFilter<S, C> where S::Sample: GpuSample {
    type Sample = S::Sample;
    fn next(&mut self) -> Option<Self::Sample> {
        while self.next >= self.output.len() {
            self.input.clear();
            while self.input.len() < self.block {
                match self.signal.next() { Some(v) => self.input.push(v), None => break }
            }
            if self.input.is_empty() { return None; }
            self.fir.process(&self.input, &mut self.output).unwrap();
            self.next = 0;
        }
        let v = self.output[self.next];
        self.next += 1;
        Some(v)
    }
    fn rate(&self) -> f32 { self.signal.rate() }
}

/// body of fft::fft after `let mut data: Vec<_> = input.iter().collect();` (src/fft.rs:8-27)
pub fn fft_shifted(data: &[Complex<f32>], rate: f32) -> Result<Vec<(f32, Complex<f32>)>, Error> {
    if data.is_empty() { return Ok(vec![]); }
    let cfg = sdr_fft_config_t { n: data.len(), input_format: SDR_FMT_C64, flags: SDR_FFT_SHIFT | SDR_FFT_NORM,
                                 device: 0, stream: std::ptr::null_mut() };
    let mut err = 0;
    let plan = unsafe { sdr_fft_create(&cfg, &mut err) };
    if plan.is_null() { return Err(Error(err)); }
    let mut vals = vec![Complex::new(0.0f32, 0.0); data.len()];
    let mut labels = vec![0.0f32; data.len()];
    let rc = unsafe { sdr_fft_exec(plan, data.as_ptr() as *const c_void, 1, vals.as_mut_ptr() as *mut f32) };
    unsafe { sdr_fft_destroy(plan) };
    check(rc)?;
    check(unsafe { sdr_fft_labels(data.len(), rate, 0, labels.as_mut_ptr()) })?;
    Ok(labels.into_iter().zip(vals.into_iter()).collect())
}
