//! Raw bindings to `include/mz_b200.h` (ABI version 1).  Layout is what `bindgen` emits for the
//! header; kept by hand because this repository's build image has no Rust toolchain.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const MZ_ABI_VERSION: u32 = 1;

pub const MZ_OK: c_int = 0;
pub const MZ_ERR_BAD_ARG: c_int = 1;
pub const MZ_ERR_W_RANGE: c_int = 2;
pub const MZ_ERR_TOO_LONG: c_int = 3;
pub const MZ_ERR_EVEN_L: c_int = 4;
pub const MZ_ERR_OPEN_EVEN_W: c_int = 5;
pub const MZ_ERR_NOT_CANONICAL: c_int = 6;
pub const MZ_ERR_VALUE_WIDTH: c_int = 7;
pub const MZ_ERR_CAPACITY: c_int = 8;
pub const MZ_ERR_UNSUPPORTED: c_int = 9;
pub const MZ_ERR_NO_DEVICE: c_int = 10;
pub const MZ_ERR_CUDA: c_int = 11;
pub const MZ_ERR_NOMEM: c_int = 12;

pub const MZ_MODE_MINIMIZER: u32 = 0;
pub const MZ_MODE_CLOSED_SYNCMER: u32 = 1;
pub const MZ_MODE_OPEN_SYNCMER: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct mz_params {
    pub k: u32,
    pub w: u32,
    pub mode: u32,
    pub strand_tiebreak: u32,
    pub hash_canonical: u32,
    pub rot: u32,
    pub f: [u32; 4],
    pub c: [u32; 4],
    pub want_sk: u32,
    pub value_bits: u32,
    pub reserved: [u32; 2],
}

#[repr(C)]
#[derive(Debug)]
pub struct mz_out {
    pub pos: *mut u32,
    pub sk: *mut u32,
    pub val: *mut u64,
    pub capacity: u64,
    pub count: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct mz_timing {
    pub h2d_ms: f32,
    pub kernel_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
    pub kernel_launches: u32,
    pub reserved: u32,
}

#[repr(C)]
pub struct mz_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct mz_pcie_result {
    pub h2d_gbs: f64,
    pub d2h_gbs: f64,
    pub bidir_h2d_gbs: f64,
    pub bidir_d2h_gbs: f64,
    pub n_devices: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct mz_alu_result {
    pub lane_ops_per_s: f64,
    pub ms: f32,
    pub sm_count: u32,
}

extern "C" {
    pub fn mz_abi_version() -> u32;
    pub fn mz_strerror(code: c_int) -> *const c_char;
    pub fn mz_last_error() -> *const c_char;
    pub fn mz_device_count(n: *mut c_int) -> c_int;
    pub fn mz_params_nthash(p: *mut mz_params, k: u32, w: u32, mode: u32, canonical: u32) -> c_int;
    pub fn mz_params_mulhash(p: *mut mz_params, k: u32, w: u32, mode: u32, canonical: u32) -> c_int;
    pub fn mz_params_set_nthash(p: *mut mz_params, hash_canonical: u32) -> c_int;
    pub fn mz_params_set_mulhash(p: *mut mz_params, hash_canonical: u32) -> c_int;
    pub fn mz_params_set_tables(p: *mut mz_params, f: *const u32, c: *const u32, rot: u32,
                                hash_canonical: u32) -> c_int;
    pub fn mz_params_validate(p: *const mz_params, n_bp: u64) -> c_int;
    pub fn mz_ctx_create(device_ids: *const c_int, n_devices: c_int, ctx: *mut *mut mz_ctx) -> c_int;
    pub fn mz_ctx_destroy(ctx: *mut mz_ctx);
    pub fn mz_ctx_device_count(ctx: *const mz_ctx) -> c_int;
    pub fn mz_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn mz_host_free(p: *mut c_void);
    pub fn mz_run(ctx: *mut mz_ctx, p: *const mz_params, packed: *const u8, bp_offset: u64,
                  n_bp: u64, out: *mut mz_out) -> c_int;
    pub fn mz_run_device(ctx: *mut mz_ctx, dev_index: c_int, p: *const mz_params,
                         d_packed: *const c_void, bp_offset: u64, n_bp: u64, win_begin: u64,
                         win_end: u64, d_out: *mut mz_out) -> c_int;
    pub fn mz_run_batch(ctx: *mut mz_ctx, p: *const mz_params, packed: *const u8,
                        packed_bytes: u64, n_reads: u64, read_start_bp: *const u64,
                        read_len_bp: *const u32, stride_bytes: u64, fixed_len_bp: u32,
                        out_offsets: *mut u64, out: *mut mz_out) -> c_int;
    pub fn mz_pack_ascii(ctx: *mut mz_ctx, ascii: *const c_char, n: u64, packed_out: *mut u8) -> c_int;
    pub fn mz_run_ascii(ctx: *mut mz_ctx, p: *const mz_params, ascii: *const c_char, n: u64,
                        out: *mut mz_out) -> c_int;
    pub fn mz_last_timing(ctx: *const mz_ctx, t: *mut mz_timing) -> c_int;
    pub fn mz_run_skip_ambiguous(ctx: *mut mz_ctx, p: *const mz_params, packed: *const u8, bp_offset: u64,
                                 n_bp: u64, ambiguous: *const u8, amb_bit_offset: u64,
                                 out: *mut mz_out) -> c_int;
    pub fn mz_run_device_skip_ambiguous(ctx: *mut mz_ctx, dev_index: c_int, p: *const mz_params,
                                        d_packed: *const c_void, bp_offset: u64, n_bp: u64,
                                        d_ambiguous: *const c_void, amb_bit_offset: u64,
                                        win_begin: u64, win_end: u64, d_out: *mut mz_out) -> c_int;
    pub fn mz_pack_ascii_n(ctx: *mut mz_ctx, ascii: *const c_char, n: u64, packed_out: *mut u8,
                           ambiguous_out: *mut u8) -> c_int;
    pub fn mz_run_ascii_skip_ambiguous(ctx: *mut mz_ctx, p: *const mz_params, ascii: *const c_char,
                                       n: u64, out: *mut mz_out) -> c_int;
    pub fn mz_values(ctx: *mut mz_ctx, p: *const mz_params, packed: *const u8, bp_offset: u64, n_bp: u64,
                     pos: *const u32, n_pos: u64, value_bits: u32, val_out: *mut u64) -> c_int;
    pub fn mz_run_bucket_stats(ctx: *mut mz_ctx, p: *const mz_params, packed: *const u8, bp_offset: u64,
                               n_bp: u64, n_buckets: u32, superkmers_out: *mut u64, windows_out: *mut u64,
                               n_minimizers: *mut u64) -> c_int;
    pub fn mz_pcie_probe(ctx: *mut mz_ctx, bytes_per_device: u64, reps: u32, res: *mut mz_pcie_result) -> c_int;
    pub fn mz_alu_probe(ctx: *mut mz_ctx, dev_index: c_int, res: *mut mz_alu_result) -> c_int;
}
