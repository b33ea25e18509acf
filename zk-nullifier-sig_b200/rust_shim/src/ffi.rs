//! Raw bindings of include/plume_b200.h (hand-written; the header is the source of truth).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct plume_ctx {
    _private: [u8; 0],
}

pub const PLUME_OK: c_int = 0;
pub const PLUME_STATUS_OK: u8 = 0;
pub const PLUME_STATUS_BAD_R: u8 = 1;
pub const PLUME_STATUS_BAD_SK: u8 = 2;
pub const PLUME_STATUS_BAD_C: u8 = 3;
pub const PLUME_STATUS_ZERO_S: u8 = 4;
pub const PLUME_STATUS_H_INF: u8 = 5;

extern "C" {
    pub fn plume_version() -> c_int;
    pub fn plume_ctx_create(out: *mut *mut plume_ctx, device: c_int, fixed_window_bits: c_int) -> c_int;
    /// one context over several GPUs of this process: host-pointer batch calls range-split over them
    pub fn plume_ctx_create_multi(out: *mut *mut plume_ctx, devices: *const c_int, n_devices: c_int, fixed_window_bits: c_int) -> c_int;
    pub fn plume_ctx_device_count(ctx: *const plume_ctx) -> c_int;
    pub fn plume_ctx_sub(ctx: *mut plume_ctx, i: c_int) -> *mut plume_ctx;
    pub fn plume_ctx_destroy(ctx: *mut plume_ctx);
    pub fn plume_last_error(ctx: *const plume_ctx) -> *const c_char;
    pub fn plume_ctx_chunk_items(ctx: *const plume_ctx) -> usize;
    pub fn plume_sign_batch(
        ctx: *mut plume_ctx, version: c_int, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        sk: *const u8, r: *const u8, pk: *mut u8, nullifier: *mut u8, c: *mut u8, s: *mut u8,
        r_point: *mut u8, hashed_to_curve_r: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn plume_verify_batch(
        ctx: *mut plume_ctx, version: c_int, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        pk: *const u8, nullifier: *const u8, c: *const u8, s: *const u8,
        r_point: *const u8, hashed_to_curve_r: *const u8, ok: *mut u8,
    ) -> c_int;
    pub fn plume_hash_to_curve_batch(
        ctx: *mut plume_ctx, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize, out: *mut u8,
    ) -> c_int;
    pub fn plume_hash_to_curve_pk_batch(
        ctx: *mut plume_ctx, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize, pk33: *const u8, out: *mut u8,
    ) -> c_int;
    pub fn plume_sign_batch_device(
        ctx: *mut plume_ctx, version: c_int, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        sk: *const u8, r: *const u8, pk: *mut u8, nullifier: *mut u8, c: *mut u8, s: *mut u8,
        r_point: *mut u8, hashed_to_curve_r: *mut u8, status: *mut u8, stream: *mut c_void,
    ) -> c_int;
    // SEC1-compressed forms, the arkworks flavour and the circuit-input side (include/plume_b200.h); UNTESTED like the rest
    pub fn plume_points_compress_batch(ctx: *mut plume_ctx, n: usize, in64: *const u8, out33: *mut u8) -> c_int;
    pub fn plume_points_decompress_batch(ctx: *mut plume_ctx, n: usize, in33: *const u8, out64: *mut u8, ok: *mut u8) -> c_int;
    pub fn plume_ark_sign_batch(
        ctx: *mut plume_ctx, version: c_int, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        pk: *const u8, sk: *const u8, r: *const u8, nullifier: *mut u8, digest_private: *mut u8, s: *mut u8,
        r_point: *mut u8, hashed_to_curve_r: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn plume_ark_verify_batch(
        ctx: *mut plume_ctx, version: c_int, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        pk: *const u8, nullifier: *const u8, digest_private: *const u8, s: *const u8,
        r_point: *const u8, hashed_to_curve_r: *const u8, ok: *mut u8,
    ) -> c_int;
    pub fn plume_hash_to_curve_witness_batch(
        ctx: *mut plume_ctx, n: usize, msgs: *const u8, msg_offsets: *const u64, msg_len: usize,
        u: *mut u8, q: *mut u8, gx1_square: *mut u8, h: *mut u8, hints: *mut u8,
    ) -> c_int;
    pub fn plume_self_test(ctx: *mut plume_ctx) -> c_int;
    pub fn plume_fixed_base_mul_batch(ctx: *mut plume_ctx, n: usize, scalars: *const u8, out: *mut u8) -> c_int;
    pub fn plume_registers_batch(ctx: *mut plume_ctx, n: usize, in32: *const u8, out4: *mut u64) -> c_int;
}
