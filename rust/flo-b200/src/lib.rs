//! Drop-in for `libflo_audio::Encoder` (libflo/src/lossless/encoder.rs:9-45) backed by the B200 library.
//!
//! Same constructor, builder and `encode` signature as the reference; the bytes returned are
//! identical.  `FloResult<T> = Result<T, String>` as in libflo/src/core/types.rs:281.
//! There is no CPU fallback: without a usable CUDA device `encode` returns `Err`.
pub mod ffi;

use std::ffi::CStr;
use std::ptr;
use std::sync::{Arc, Mutex, OnceLock};

pub type FloResult<T> = Result<T, String>;

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::flo_last_error()).to_string_lossy().into_owned() }
}

/// One GPU context (streams + scratch).  `Send + Sync`: the C library serialises calls per context.
pub struct Context(*mut ffi::flo_ctx);
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    pub fn new(device: i32) -> FloResult<Arc<Context>> {
        let mut p = ptr::null_mut();
        let rc = unsafe { ffi::flo_ctx_create(device, &mut p) };
        if rc != 0 { return Err(last_error()); }
        Ok(Arc::new(Context(p)))
    }
    /// Process-wide context of device 0, created on first use.
    pub fn default_device() -> FloResult<Arc<Context>> {
        static CTX: OnceLock<Mutex<Option<Arc<Context>>>> = OnceLock::new();
        let cell = CTX.get_or_init(|| Mutex::new(None));
        let mut g = cell.lock().map_err(|e| e.to_string())?;
        if g.is_none() { *g = Some(Context::new(0)?); }
        Ok(g.as_ref().unwrap().clone())
    }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { ffi::flo_ctx_destroy(self.0) } }
}

/// One track of a batch (`encode_batch`).
pub struct TrackRef<'a> {
    pub samples: &'a [f32],
    pub sample_rate: u32,
    pub channels: u8,
    pub bit_depth: u8,
    pub metadata: &'a [u8],
}

/// `libflo_audio::Encoder` (encoder.rs:9-14).
pub struct Encoder {
    sample_rate: u32,
    channels: u8,
    bit_depth: u8,
    compression_level: u8,
    ctx: Option<Arc<Context>>,
}

impl Default for Encoder {
    fn default() -> Self { Encoder::new(44100, 1, 16) }
}

impl Encoder {
    /// encoder.rs:17-24
    pub fn new(sample_rate: u32, channels: u8, bit_depth: u8) -> Self {
        Encoder { sample_rate, channels, bit_depth, compression_level: 5, ctx: None }
    }
    /// encoder.rs:26-29
    pub fn with_compression(mut self, level: u8) -> Self {
        self.compression_level = level.min(9);
        self
    }
    /// Additive: pin the encoder to a specific GPU context.
    pub fn with_context(mut self, ctx: Arc<Context>) -> Self {
        self.ctx = Some(ctx);
        self
    }
    fn context(&self) -> FloResult<Arc<Context>> {
        match &self.ctx { Some(c) => Ok(c.clone()), None => Context::default_device() }
    }

    /// encoder.rs:32-45.  The reference panics on `channels == 0` (division by zero, encoder.rs:48);
    /// here that is `Err`.
    pub fn encode(&self, samples: &[f32], metadata: &[u8]) -> FloResult<Vec<u8>> {
        let ctx = self.context()?;
        let (mut out, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::flo_encode(ctx.0, samples.as_ptr(), samples.len(), self.sample_rate, self.channels,
                            self.bit_depth, self.compression_level, metadata.as_ptr(), metadata.len(),
                            &mut out, &mut len)
        };
        if rc != 0 { return Err(last_error()); }
        let v = unsafe { std::slice::from_raw_parts(out, len).to_vec() };
        unsafe { ffi::flo_free(out as *mut _) };
        Ok(v)
    }

    /// reflo's S16 ingest arm (reflo/src/audio.rs:247-254) + `encode`, fused on the device.
    pub fn encode_pcm16(&self, pcm: &[i16], metadata: &[u8]) -> FloResult<Vec<u8>> {
        let ctx = self.context()?;
        let (mut out, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::flo_encode_pcm16(ctx.0, pcm.as_ptr(), pcm.len(), self.sample_rate, self.channels,
                                  self.bit_depth, self.compression_level, metadata.as_ptr(), metadata.len(),
                                  &mut out, &mut len)
        };
        if rc != 0 { return Err(last_error()); }
        let v = unsafe { std::slice::from_raw_parts(out, len).to_vec() };
        unsafe { ffi::flo_free(out as *mut _) };
        Ok(v)
    }
}

/// Loop of `Encoder::encode` over tracks (reflo/src/main.rs:218-276) as one device pass.
pub fn encode_batch(ctx: &Arc<Context>, tracks: &[TrackRef<'_>], level: u8) -> FloResult<Vec<Vec<u8>>> {
    let raw: Vec<ffi::flo_track> = tracks.iter().map(|t| ffi::flo_track {
        samples: t.samples.as_ptr() as *const _,
        n_interleaved: t.samples.len(),
        sample_rate: t.sample_rate,
        channels: t.channels,
        bit_depth: t.bit_depth,
        meta: t.metadata.as_ptr(),
        meta_len: t.metadata.len(),
    }).collect();
    let mut outs: Vec<ffi::flo_out> = (0..tracks.len()).map(|_| ffi::flo_out { data: ptr::null_mut(), len: 0 }).collect();
    let rc = unsafe { ffi::flo_encode_batch(ctx.0, raw.as_ptr(), raw.len(), ffi::FLO_FMT_F32, level.min(9), outs.as_mut_ptr()) };
    if rc != 0 { return Err(last_error()); }
    Ok(outs.into_iter().map(|o| {
        let v = unsafe { std::slice::from_raw_parts(o.data, o.len).to_vec() };
        unsafe { ffi::flo_free(o.data as *mut _) };
        v
    }).collect())
}

/// `libflo_audio::Decoder` (libflo/src/lossless/decoder.rs:6-18) over `flo_decode`.
#[derive(Default)]
pub struct Decoder {
    ctx: Option<Arc<Context>>,
}

impl Decoder {
    pub fn new() -> Self {
        Decoder { ctx: None }
    }

    pub fn with_context(mut self, ctx: Arc<Context>) -> Self {
        self.ctx = Some(ctx);
        self
    }

    /// decoder.rs:14-18: interleaved f32 samples.
    pub fn decode(&self, data: &[u8]) -> FloResult<Vec<f32>> {
        let ctx = match &self.ctx { Some(c) => c.clone(), None => Context::default_device()? };
        let (mut out, mut n) = (ptr::null_mut::<f32>(), 0usize);
        let rc = unsafe { ffi::flo_decode(ctx.0, data.as_ptr(), data.len(), &mut out, &mut n, ptr::null_mut()) };
        if rc != 0 { return Err(last_error()); }
        let v = unsafe { std::slice::from_raw_parts(out, n).to_vec() };
        unsafe { ffi::flo_free(out as *mut _) };
        Ok(v)
    }
}

/// `core::analysis::extract_waveform_peaks` (libflo/src/core/analysis.rs:38-119): the normalised peaks of
/// `WaveformData`, bit-identical to the reference.  Same argument order as the reference.
pub fn extract_waveform_peaks(ctx: &Arc<Context>, samples: &[f32], channels: u8, sample_rate: u32, peaks_per_second: u32) -> FloResult<Vec<f32>> {
    let (mut out, mut n) = (ptr::null_mut::<f32>(), 0usize);
    let rc = unsafe { ffi::flo_waveform_peaks(ctx.0, samples.as_ptr(), samples.len(), sample_rate, channels, peaks_per_second, &mut out, &mut n) };
    if rc != 0 { return Err(last_error()); }
    if out.is_null() { return Ok(Vec::new()); }
    let v = unsafe { std::slice::from_raw_parts(out, n).to_vec() };
    unsafe { ffi::flo_free(out as *mut _) };
    Ok(v)
}

/// `core::ebu_r128::compute_ebu_r128_loudness(..).integrated_lufs` (libflo/src/core/ebu_r128.rs:182-313), the value
/// `libflo::encode()` stores as `loudness_profile[0].lufs`; agrees with the reference to ~1e-12 LU (see the header).
pub fn integrated_loudness(ctx: &Arc<Context>, samples: &[f32], channels: u8, sample_rate: u32) -> FloResult<f64> {
    let mut lufs = 0f64;
    let rc = unsafe { ffi::flo_integrated_loudness(ctx.0, samples.as_ptr(), samples.len(), sample_rate, channels, &mut lufs) };
    if rc != 0 { return Err(last_error()); }
    Ok(lufs)
}
