//! Raw bindings to `include/qoa_b200.h` (hand-written; one item per C declaration).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const QOA_B200_OK: i32 = 0;
pub const QOA_B200_ENC_INVALID_CHANNELS: i32 = 1;
pub const QOA_B200_ENC_INVALID_SAMPLE_RATE: i32 = 2;
pub const QOA_B200_ENC_INVALID_SAMPLES: i32 = 3;
pub const QOA_B200_ENC_IO_ERROR: i32 = 4;
pub const QOA_B200_DEC_NOT_QOA_FILE: i32 = 11;
pub const QOA_B200_DEC_NO_SAMPLES: i32 = 12;
pub const QOA_B200_DEC_INVALID_FRAME_HEADER: i32 = 13;
pub const QOA_B200_DEC_INCOMPATIBLE_FRAME: i32 = 14;
pub const QOA_B200_DEC_IO_ERROR: i32 = 15;
pub const QOA_B200_DEC_OUTPUT_TOO_SMALL: i32 = 16;
pub const QOA_B200_DECODE_ITER: c_int = 0;
pub const QOA_B200_DECODE_ALL: c_int = 1;
pub const QOA_B200_DECODE_FRAMES: c_int = 2;

#[repr(C)] pub struct qoa_b200_ctx { _p: [u8; 0] }
#[repr(C)] pub struct qoa_b200_encoder_batch { _p: [u8; 0] }
#[repr(C)] pub struct qoa_b200_decode_plan { _p: [u8; 0] }
#[repr(C)] pub struct qoa_b200_mctx { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct qoa_b200_desc { pub channels: u32, pub sample_rate: u32, pub samples: u32 }

#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct qoa_b200_stream_info {
    pub status: i32, pub streaming: i32, pub header_samples: u32, pub channels: u32, pub sample_rate: u32,
    pub index_frames: u32, pub samples: u64, pub frames: u64, pub max_channels: u32, pub flags: u32,
}

#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct qoa_b200_frame_header { pub num_channels: u32, pub sample_rate: u32, pub num_samples_per_channel: u32 }

extern "C" {
    pub fn qoa_b200_version() -> *const c_char;
    pub fn qoa_b200_ctx_create(device: c_int, ctx: *mut *mut qoa_b200_ctx) -> c_int;
    pub fn qoa_b200_ctx_destroy(ctx: *mut qoa_b200_ctx);
    pub fn qoa_b200_last_error(ctx: *const qoa_b200_ctx) -> *const c_char;
    pub fn qoa_b200_kernel_launches(ctx: *const qoa_b200_ctx) -> u64;
    pub fn qoa_b200_encoded_size(channels: u32, samples: u32) -> usize;
    pub fn qoa_b200_frame_size(channels: u32, frame_samples: u32) -> usize;
    pub fn qoa_b200_encode_all_batch(ctx: *mut qoa_b200_ctx, descs: *const qoa_b200_desc, pcm: *const *const i16,
        pcm_len: *const usize, out: *const *mut u8, out_cap: *const usize, out_len: *mut usize, status: *mut i32,
        n: usize) -> c_int;
    pub fn qoa_b200_encode_all_batch_dev(ctx: *mut qoa_b200_ctx, descs: *const qoa_b200_desc, d_pcm: *const i16,
        pcm_off: *const u64, pcm_len: *const usize, d_out: *mut u8, out_off: *const u64, out_cap: *const usize,
        out_len: *mut usize, status: *mut i32, n: usize, stream: *mut c_void) -> c_int;
    pub fn qoa_b200_encoder_batch_new(ctx: *mut qoa_b200_ctx, descs: *const qoa_b200_desc, n: usize,
        status: *mut i32, enc: *mut *mut qoa_b200_encoder_batch) -> c_int;
    pub fn qoa_b200_encoder_batch_free(enc: *mut qoa_b200_encoder_batch);
    pub fn qoa_b200_encoder_batch_write_header(enc: *const qoa_b200_encoder_batch, out: *const *mut u8) -> c_int;
    pub fn qoa_b200_encoder_batch_encode_frame(enc: *mut qoa_b200_encoder_batch, pcm: *const *const i16,
        pcm_len: *const usize, out: *const *mut u8, out_cap: *const usize, written: *mut usize, status: *mut i32) -> c_int;
    pub fn qoa_b200_encoder_batch_encode_chunk(enc: *mut qoa_b200_encoder_batch, pcm: *const *const i16,
        pcm_len: *const usize, out: *const *mut u8, out_cap: *const usize, written: *mut usize, status: *mut i32) -> c_int;
    pub fn qoa_b200_encoder_batch_encode_chunk_dev(enc: *mut qoa_b200_encoder_batch, d_pcm: *const i16,
        pcm_off: *const u64, chunk_samples: *const u32, d_out: *mut u8, out_off: *const u64, stream: *mut c_void) -> c_int;
    pub fn qoa_b200_encoder_batch_get_lms(enc: *mut qoa_b200_encoder_batch, i: usize, c: u32, history: *mut i32,
        weights: *mut i32) -> c_int;
    pub fn qoa_b200_decode_plan_create(ctx: *mut qoa_b200_ctx, qoa: *const *const u8, len: *const usize, n: usize,
        mode: c_int, plan: *mut *mut qoa_b200_decode_plan) -> c_int;
    pub fn qoa_b200_decode_plan_create_dev(ctx: *mut qoa_b200_ctx, d_qoa: *const u8, in_off: *const u64,
        len: *const usize, n: usize, mode: c_int, stream: *mut c_void, plan: *mut *mut qoa_b200_decode_plan) -> c_int;
    pub fn qoa_b200_decode_plan_destroy(plan: *mut qoa_b200_decode_plan);
    pub fn qoa_b200_decode_plan_info(plan: *const qoa_b200_decode_plan, info: *mut qoa_b200_stream_info) -> c_int;
    pub fn qoa_b200_decode_plan_frame_headers(plan: *const qoa_b200_decode_plan, i: usize,
        hdrs: *mut qoa_b200_frame_header, cap: usize) -> usize;
    pub fn qoa_b200_decode_plan_run(plan: *mut qoa_b200_decode_plan, out: *const *mut i16, out_cap: *const usize,
        status: *mut i32) -> c_int;
    pub fn qoa_b200_decode_plan_run_dev(plan: *mut qoa_b200_decode_plan, d_out: *mut i16, out_off: *const u64,
        out_cap: *const usize, status: *mut i32, stream: *mut c_void) -> c_int;
    pub fn qoa_b200_decode_all_batch(ctx: *mut qoa_b200_ctx, qoa: *const *const u8, len: *const usize, n: usize,
        mode: c_int, out: *const *mut i16, out_cap: *const usize, info: *mut qoa_b200_stream_info, status: *mut i32) -> c_int;
    pub fn qoa_b200_gen_pcm_dev(ctx: *mut qoa_b200_ctx, d_pcm: *mut i16, n_streams: u64, first_stream: u64,
        channels: u32, samples: u64, start: u64, seed: u64, kind: c_int, stream: *mut c_void) -> c_int;
    pub fn qoa_b200_gen_qoa_dev(ctx: *mut qoa_b200_ctx, d_qoa: *mut u8, stride: u64, n_streams: u64, first_stream: u64,
        channels: u32, sample_rate: u32, samples: u32, seed: u64, stream: *mut c_void) -> c_int;
    pub fn qoa_b200_int_issue_peak(ctx: *mut qoa_b200_ctx, kind: c_int, thread_ops_per_s: *mut f64) -> c_int;
    pub fn qoa_b200_probe(qoa: *const u8, len: usize, mode: c_int, info: *mut qoa_b200_stream_info) -> c_int;
    // several GPUs from one process (batches sharded by stream, no collective)
    pub fn qoa_b200_ctx_create_multi(devices: *const c_int, n: c_int, mctx: *mut *mut qoa_b200_mctx) -> c_int;
    pub fn qoa_b200_mctx_destroy(mctx: *mut qoa_b200_mctx);
    pub fn qoa_b200_mctx_device_count(mctx: *const qoa_b200_mctx) -> c_int;
    pub fn qoa_b200_mctx_device(mctx: *const qoa_b200_mctx, i: c_int) -> c_int;
    pub fn qoa_b200_mctx_ctx(mctx: *mut qoa_b200_mctx, i: c_int) -> *mut qoa_b200_ctx;
    pub fn qoa_b200_mctx_last_error(mctx: *const qoa_b200_mctx) -> *const c_char;
    pub fn qoa_b200_mctx_kernel_launches(mctx: *const qoa_b200_mctx) -> u64;
    pub fn qoa_b200_partition(weight: *const u64, n: usize, parts: c_int, first: *mut usize) -> c_int;
    pub fn qoa_b200_mctx_partition(mctx: *const qoa_b200_mctx, weight: *const u64, n: usize, first: *mut usize) -> c_int;
    pub fn qoa_b200_mctx_encode_all_batch(mctx: *mut qoa_b200_mctx, descs: *const qoa_b200_desc, pcm: *const *const i16,
        pcm_len: *const usize, out: *const *mut u8, out_cap: *const usize, out_len: *mut usize, status: *mut i32,
        n: usize) -> c_int;
    pub fn qoa_b200_mctx_decode_all_batch(mctx: *mut qoa_b200_mctx, qoa: *const *const u8, len: *const usize, n: usize,
        mode: c_int, out: *const *mut i16, out_cap: *const usize, info: *mut qoa_b200_stream_info,
        status: *mut i32) -> c_int;
}
