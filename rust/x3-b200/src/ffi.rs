//! Raw bindings of include/x3_b200.h.  One-to-one with the C declarations.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct x3_params {
    pub block_len: u32,
    pub blocks_per_frame: u32,
    pub codes: [u32; 3],
    pub thresholds: [u32; 3],
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct x3_stats {
    pub samples_by_mode: [u64; 6],
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct x3_frame_header {
    pub source_id: u8,
    pub channels: u8,
    pub samples: u16,
    pub payload_len: u32,
    pub payload_crc: u16,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct x3_decode_result {
    pub samples: u64,
    pub frames: u64,
    pub frame_errors: u64,
    pub first_bad_frame: u64,
    pub first_bad_code: i32,
    pub used_host_walk: i32,
}

#[repr(C)]
pub struct x3_device_result { pub value: u64, pub flags: u64, pub detail: [u64; 6] }

extern "C" {
    pub fn x3_abi_version() -> c_int;
    pub fn x3_params_default(p: *mut x3_params) -> c_int;
    pub fn x3_params_validate(p: *const x3_params) -> c_int;
    pub fn x3_encode_bound(n_samples: usize, p: *const x3_params) -> usize;
    pub fn x3_encode_frame_bound(n_samples: usize, p: *const x3_params) -> usize;
    pub fn x3_strerror(code: c_int) -> *const c_char;
    pub fn x3_last_cuda_error() -> *const c_char;
    pub fn x3_write_frame_header(num_samples: usize, id: u8, payload_len: usize, payload_crc: u16, header: *mut u8) -> c_int;
    pub fn x3_read_frame_header(bytes: *const u8, len: usize, h: *mut x3_frame_header) -> c_int;
    pub fn x3_crc16(data: *const u8, len: usize) -> u16;
    pub fn x3_encode_host(pcm: *const i16, n: usize, p: *const x3_params, out: *mut u8, cap: usize,
                          out_len: *mut usize, stats: *mut x3_stats) -> c_int;
    pub fn x3_encode_device(d_pcm: *const i16, n: usize, p: *const x3_params, d_out: *mut u8, cap: usize,
                            out_len: *mut usize, stats: *mut x3_stats, stream: *mut c_void) -> c_int;
    pub fn x3_encode_frame_host(pcm: *const i16, n: usize, p: *const x3_params, out: *mut u8, cap: usize,
                                out_len: *mut usize, stats: *mut x3_stats) -> c_int;
    pub fn x3_decode_host(frames: *const u8, len: usize, p: *const x3_params, pcm: *mut i16, cap: usize,
                          n_out: *mut usize, res: *mut x3_decode_result) -> c_int;
    pub fn x3_decode_device(d_frames: *const u8, len: usize, p: *const x3_params, d_pcm: *mut i16, cap: usize,
                            n_out: *mut usize, res: *mut x3_decode_result, stream: *mut c_void) -> c_int;
    pub fn x3_decode_frame_host(payload: *const u8, len: usize, p: *const x3_params, pcm: *mut i16, cap: usize,
                                samples: usize, n_out: *mut usize) -> c_int;
    // frame-range sharding across GPUs (one process per GPU; the size all-gather is the host program's)
    pub fn x3_shard_range(n_samples: u64, p: *const x3_params, rank: u32, world: u32, s0: *mut u64, s1: *mut u64) -> c_int;
    pub fn x3_deal_files(frames_per_file: *const u64, n_files: usize, world: u32, rank_of_file: *mut u32) -> c_int;
    pub fn x3_shard_base(shard_bytes: *const u64, world: u32, rank: u32, base: *mut u64) -> c_int;
    pub fn x3_place_shard_device(d_stream: *mut u8, base: u64, d_shard: *const u8, shard_bytes: usize, stream: *mut c_void) -> c_int;
    pub fn x3_synth_device(kind: c_int, seed: u32, fs: u32, n0: u64, count: u64, d_out: *mut i16, stream: *mut c_void) -> c_int;
    pub fn x3_kernel_launch_count() -> u64;
    pub fn x3_last_kernel_ms(ms: *mut f32) -> c_int;
    pub fn x3_last_encode_kernel() -> c_int;
    // stream-ordered variants (ABI v3): results stay on the device in an x3_device_result
    pub fn x3_encode_device_async(d_pcm: *const i16, n: usize, p: *const x3_params, d_out: *mut u8, cap: usize,
                                  d_res: *mut x3_device_result, stream: *mut c_void) -> c_int;
    pub fn x3_decode_device_async(d_frames: *const u8, len_cap: usize, d_len: *const u64, p: *const x3_params,
                                  d_pcm: *mut i16, cap: usize, d_res: *mut x3_device_result, stream: *mut c_void) -> c_int;
}
