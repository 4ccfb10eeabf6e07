//! Raw bindings of include/flo_b200.h (one declaration per exported symbol).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct flo_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct flo_track {
    pub samples: *const c_void,
    pub n_interleaved: usize,
    pub sample_rate: u32,
    pub channels: u8,
    pub bit_depth: u8,
    pub meta: *const u8,
    pub meta_len: usize,
}

#[repr(C)]
pub struct flo_out {
    pub data: *mut u8,
    pub len: usize,
}

#[repr(C)]
pub struct flo_cand_report {
    pub k: i32,
    pub pad: i32,
    pub size: i64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct flo_info {
    pub sample_rate: u32,
    pub channels: u8,
    pub bit_depth: u8,
    pub level: u8,
    pub version_major: u8,
    pub total_samples: u64,
    pub decoded_frames: u64,
    pub n_frames: u32,
    pub data_crc32: u32,
    pub meta_offset: u64,
    pub meta_size: u64,
}

pub const FLO_FMT_F32: c_int = 0;
pub const FLO_FMT_PCM16: c_int = 1;
pub const FLO_FMT_U8: c_int = 2;
pub const FLO_FMT_S32: c_int = 3;

extern "C" {
    pub fn flo_ctx_create(device: c_int, out: *mut *mut flo_ctx) -> c_int;
    pub fn flo_ctx_destroy(ctx: *mut flo_ctx);
    pub fn flo_encode(ctx: *mut flo_ctx, samples: *const f32, n_interleaved: usize, sample_rate: u32,
                      channels: u8, bit_depth: u8, level: u8, meta: *const u8, meta_len: usize,
                      out: *mut *mut u8, out_len: *mut usize) -> c_int;
    pub fn flo_encode_pcm16(ctx: *mut flo_ctx, pcm: *const i16, n_interleaved: usize, sample_rate: u32,
                            channels: u8, bit_depth: u8, level: u8, meta: *const u8, meta_len: usize,
                            out: *mut *mut u8, out_len: *mut usize) -> c_int;
    pub fn flo_encode_batch(ctx: *mut flo_ctx, tracks: *const flo_track, n_tracks: usize, format: c_int,
                            level: u8, outs: *mut flo_out) -> c_int;
    pub fn flo_encode_batch_device(ctx: *mut flo_ctx, tracks: *const flo_track, n_tracks: usize, format: c_int,
                                   level: u8, d_out: *mut c_void, d_out_capacity: usize,
                                   offsets: *mut u64, lens: *mut u64) -> c_int;
    pub fn flo_stream_encode_frames(ctx: *mut flo_ctx, samples: *const f32, n_interleaved: usize, sample_rate: u32,
                                    channels: u8, bit_depth: u8, level: u8, out: *mut *mut u8, out_len: *mut usize,
                                    frame_off: *mut *mut u64, n_frames: *mut u32) -> c_int;
    pub fn flo_waveform_peaks(ctx: *mut flo_ctx, samples: *const f32, n_interleaved: usize, sample_rate: u32, channels: u8,
                              peaks_per_second: u32, peaks: *mut *mut f32, n_peaks: *mut usize) -> c_int;
    pub fn flo_waveform_peaks_device(ctx: *mut flo_ctx, d_samples: *const f32, n_interleaved: usize, sample_rate: u32,
                                     channels: u8, peaks_per_second: u32, d_peaks: *mut f32, capacity: usize,
                                     n_peaks: *mut usize) -> c_int;
    pub fn flo_waveform_peaks_count(n_interleaved: usize, sample_rate: u32, channels: u8, peaks_per_second: u32) -> usize;
    pub fn flo_integrated_loudness(ctx: *mut flo_ctx, samples: *const f32, n_interleaved: usize, sample_rate: u32,
                                   channels: u8, lufs: *mut f64) -> c_int;
    pub fn flo_integrated_loudness_device(ctx: *mut flo_ctx, d_samples: *const f32, n_interleaved: usize, sample_rate: u32,
                                          channels: u8, lufs: *mut f64) -> c_int;
    pub fn flo_decode(ctx: *mut flo_ctx, file: *const u8, len: usize, out: *mut *mut f32, n_interleaved: *mut usize,
                      info: *mut flo_info) -> c_int;
    pub fn flo_decode_device(ctx: *mut flo_ctx, d_file: *const c_void, len: usize, d_out: *mut f32,
                             d_out_capacity: usize, n_interleaved: *mut usize, info: *mut flo_info) -> c_int;
    pub fn flo_decode_i16(ctx: *mut flo_ctx, file: *const u8, len: usize, out: *mut *mut i16, n_interleaved: *mut usize,
                          info: *mut flo_info) -> c_int;
    pub fn flo_decode_i16_device(ctx: *mut flo_ctx, d_file: *const c_void, len: usize, d_out: *mut i16,
                                 d_out_capacity: usize, n_interleaved: *mut usize, info: *mut flo_info) -> c_int;
    pub fn flo_output_bound(tracks: *const flo_track, n_tracks: usize) -> usize;
    pub fn flo_ctx_set_stream(ctx: *mut flo_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn flo_ctx_last_timing(ctx: *mut flo_ctx, ms: *mut f32, launches: *mut u32) -> c_int;
    pub fn flo_ctx_last_counters(ctx: *mut flo_ctx, out: *mut u64) -> c_int;
    pub fn flo_ctx_enable_report(ctx: *mut flo_ctx, enable: c_int) -> c_int;
    pub fn flo_ctx_read_report(ctx: *mut flo_ctx, frame: u32, channel: u32, out: *mut flo_cand_report) -> c_int;
    pub fn flo_host_alloc(bytes: usize) -> *mut c_void;
    pub fn flo_host_free(p: *mut c_void);
    pub fn flo_free(p: *mut c_void);
    pub fn flo_last_error() -> *const c_char;
    pub fn flo_version() -> *const c_char;
    pub fn flo_device_count() -> c_int;
}
