/*
 * plume_b200.h -- C ABI of the B200-native batch PLUME signer / verifier.
 *
 * This is the drop-in boundary for the per-signature hot path of the reference crate
 * `plume_rustcrypto` (plume-sig/zk-nullifier-sig, rust-k256/).  The reference has no FFI of its
 * own; each entry point below states which reference function it replaces, so that a Rust shim
 *  can keep `PlumeSignature::sign_v1 / sign_v2 / verify` and forward batches here.
 *
 * Conventions
 *   - scalars and field elements: 32 bytes, big-endian (k256 `FieldBytes` / SEC1 order);
 *   - curve points: 64 bytes `x || y` (affine, big-endian); 64 zero bytes denote the identity;
 *   - all arrays are structure-of-arrays, item i at offset i * element_size;
 *   - messages: either `msg_offsets` (n + 1 byte offsets into `msgs`) or, when it is NULL,
 *     fixed-length records of `msg_len` bytes each;
 *   - functions return 0 on success and a negative PLUME_E_* code on failure
 *     (text via plume_last_error); nothing ever aborts or unwinds across this boundary;
 *   - per-item failures are reported in `status[i]` (sign) / `ok[i]` (verify);
 *   - a context is externally synchronised: one batch call at a time per context (calls on
 *     different contexts may run concurrently; plume_last_error(NULL) is per thread).
 *   - `_device` variants take device pointers (16-byte aligned for the 32/64-byte arrays) and
 *     enqueue on the caller's CUDA stream without synchronising; the plain variants take host
 *     pointers (pinned memory is used in place, pageable memory is staged through pinned buffers
 *     by a few copy threads) and return when the results are in place.
 *   - stream ordering: the `_device` entry points share one workspace of their own (separate from
 *     the host-pointer path's), and every `_device` call is ordered after the previous one by an
 *     event, whatever streams they were given; a host-pointer call never touches that workspace,
 *     so it may follow a `_device` call without a synchronisation in between.
 *   - msg_offsets handed to a host-pointer entry point are checked (non-decreasing, every message
 *     shorter than 4 GiB; PLUME_E_ARG otherwise); offsets in device memory are the caller's
 *     responsibility.
 *   - secrets: the library's own copies of sk and r (device arena, pinned staging arena) are zeroed
 *     as soon as the chunk that used them has run, and again in plume_ctx_destroy (the reference
 *     zeroises its witness: javascript/src/lib.rs:64-71,82; rust-arkworks/src/lib.rs:202-214).
 *     The caller's own buffers are the caller's to clear.  Scalar multiplication is NOT constant
 *     time (zero windows are skipped, exceptional cases branch): this is a throughput engine for
 *     batches whose timing is not observable per item; k256's multiplication is constant time.
 *
 * There is no CPU fallback: every entry point fails with PLUME_E_NO_DEVICE / PLUME_E_CUDA when
 * the CUDA device or kernels are unavailable.
 */
#ifndef PLUME_B200_H
#define PLUME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLUME_ABI_VERSION 2   /* 2: multi-device contexts, hints argument of the witness call */

/* return codes */
#define PLUME_OK 0
#define PLUME_E_ARG (-1)        /* bad argument (null pointer, version not 1/2, ...) */
#define PLUME_E_NO_DEVICE (-2)  /* no CUDA device / device index out of range */
#define PLUME_E_CUDA (-3)       /* CUDA runtime error, see plume_last_error */
#define PLUME_E_NOMEM (-4)      /* allocation failed */
#define PLUME_E_SELFTEST (-5)   /* plume_self_test: a result differs from the reference's vectors, see plume_last_error */

/* per-item status of plume_sign_batch: the places where the reference panics
 * (rust-k256/src/randomizedsigner.rs:61, :91, :95) or rejects its inputs */
#define PLUME_STATUS_OK 0
#define PLUME_STATUS_BAD_R 1    /* r not in [1, n-1]   (SecretKey::random never yields it; :49) */
#define PLUME_STATUS_BAD_SK 2   /* sk not in [1, n-1]  (SecretKey::from_bytes rejects it) */
#define PLUME_STATUS_BAD_C 3    /* SHA-256 output c = 0 or c >= n (NonZeroScalar::from_repr, :90-91) */
#define PLUME_STATUS_ZERO_S 4   /* s = r + c*sk = 0 (:95) */
#define PLUME_STATUS_H_INF 5    /* hash_to_curve returned the identity (:61) */
#define PLUME_STATUS_BAD_PK 6   /* arkworks flavour: pk is the identity (hash_to_curve fails, rust-arkworks/src/lib.rs:97-100) or not a curve point */

typedef struct plume_ctx plume_ctx;

/* ABI / library identification. */
int plume_version(void);

/* Create a context on CUDA device `device` (ordinal as seen by the CUDA runtime in this
 * process).  Builds the fixed-base table for the generator on the device (one-off, ~0.1 s).
 * `fixed_window_bits` = 0 picks the default (22: a 3.2 GB table, 12 additions per multiplication, built in ~0.4 s);
 * otherwise 4..24 (16 = 64 MiB that stay in the L2, 16 additions; 20 = 872 MB, 13 additions, ~0.1 s).
 * A multi-GPU job is one process (context) per GPU, each given its contiguous range of the
 * batch; nothing is shared between contexts. */
int plume_ctx_create(plume_ctx** out, int device, int fixed_window_bits);
void plume_ctx_destroy(plume_ctx* ctx);

/* One context over several GPUs of this process (SURVEY.md section 8b "plume_ctx_create(devices[], n)", 8e): one
 * sub-context and one worker thread per device.  The host-pointer entry points of such a context range-split the batch
 * -- device g takes items [g n / G, (g+1) n / G) -- run the per-device pipelines concurrently and write the results
 * straight into the caller's arrays; there is no exchange step.  The `_device` entry points have no meaning on it
 * (PLUME_E_ARG): use plume_ctx_sub(ctx, g) for device-resident work on GPU g.  Every device builds its own generator
 * table concurrently unless PLUME_GTAB_BCAST=p2p|nccl asks for device 0's table to be broadcast (peer copies over
 * NVLink / ncclBroadcast; measured, not faster: DESIGN.md section 6). */
int plume_ctx_create_multi(plume_ctx** out, const int* devices, int n_devices, int fixed_window_bits);
/* Known-answer test of the context through its own batch calls: the reference's signing vector (rust-k256/tests/signing.rs:9-21:
 * message, sk, mock-RNG nonce -> V1 and V2 c, s), its intermediate points (rust-arkworks/src/tests.rs:191-262: pk, g^r, h, h^r,
 * nullifier) and hash_to_curve("abc") (rust-k256/tests/verification.rs:288-292), 300 copies per batch so that more than one
 * warp and block take part, one tampered copy that verification has to reject.  A few milliseconds; every device of a
 * multi-device context runs it.  PLUME_OK, PLUME_E_SELFTEST (plume_last_error names the field) or the error of the failing call.
 * Meant for a binding's start-up path: the product has no CPU implementation to compare itself with. */
int plume_self_test(plume_ctx* ctx);
int plume_ctx_device_count(const plume_ctx* ctx);          /* 1 for a single-device context */
/* The range split itself (no device needed): part `part` of `parts` owns items [first, first + count) of n. */
int plume_shard_range(size_t n, int part, int parts, size_t* first, size_t* count);
plume_ctx* plume_ctx_sub(plume_ctx* ctx, int i);           /* the i-th per-device context (ctx itself when single) */

/* Last error text of this context (or of context creation when ctx is NULL). */
const char* plume_last_error(const plume_ctx* ctx);

/* Largest batch one call processes in a single pass; bigger batches are cut into chunks. */
size_t plume_ctx_chunk_items(const plume_ctx* ctx);

/*
 * Batch signing.  Replaces PlumeSigner::try_sign_with_rng (rust-k256/src/randomizedsigner.rs:43-112)
 * behind PlumeSignature::sign_v1 / sign_v2 (rust-k256/src/lib.rs:149-156).  The nonce r that the
 * reference draws with SecretKey::random(rng) (:49) is an input: the RNG stays on the caller's
 * side of the boundary.
 *   version            1 or 2
 *   sk, r              n x 32
 *   pk, nullifier      n x 64   out: g^sk, hash_to_curve(m, pk)^sk
 *   c, s               n x 32   out
 *   r_point            n x 64   out: g^r           (V1 field; may be NULL)
 *   hashed_to_curve_r  n x 64   out: h^r           (V1 field; may be NULL)
 *   status             n        out: PLUME_STATUS_*; on a non-zero status every output of the item is zero
 * Any n is legal.  Batches of at most 4 096 items (environment: PLUME_TEAM_MAX) run kernels built for latency -- 2 or 4
 * lanes of a warp per item -- with bit-identical results: one signature takes ~0.76 ms, one verification ~0.64 ms on a
 * B200; the throughput figures need 10^5 items and more.
 */
int plume_sign_batch(plume_ctx* ctx, int version, size_t n,
                     const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                     const uint8_t* sk, const uint8_t* r,
                     uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s,
                     uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status);

/*
 * Batch verification.  Replaces PlumeSignature::verify (rust-k256/src/lib.rs:93-145).
 *   version 1: r_point and hashed_to_curve_r are required (v1specific = Some{..});
 *   version 2: they are ignored (v1specific = None).
 *   ok[i] = 1 iff the reference's verify() returns true.  Inputs that the reference's types
 *   cannot represent (off-curve or non-canonical points, c or s outside [1, n-1]) give 0.
 */
int plume_verify_batch(plume_ctx* ctx, int version, size_t n,
                       const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                       const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s,
                       const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok);

/*
 * Batch hash-to-curve on caller-assembled preimages: out[i] = hash_to_curve(msgs[i]) with the
 * PLUME DST.  Replaces utils::hash_to_curve (rust-k256/src/utils.rs:11-20), whose preimage is
 * m || SEC1-compressed(pk); the caller concatenates.
 *   out  n x 64
 */
int plume_hash_to_curve_batch(plume_ctx* ctx, size_t n,
                              const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                              uint8_t* out);
/* The reference's own call shape, utils::hash_to_curve(m, pk) (rust-k256/src/utils.rs:11-20): the
 * messages and the SEC1-compressed public keys as separate arrays, the library hashes m_i || pk33_i.  pk33: n x 33 slots
 * (02/03 || x; 00 followed by zeros stands for the identity and contributes its one-byte encoding, as encode_pt does). */
int plume_hash_to_curve_pk_batch(plume_ctx* ctx, size_t n,
                                 const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                 const uint8_t* pk33, uint8_t* out);

/*
 * SEC1-compressed wire form: 33-byte slots, `02/03 || x` for a finite point (what
 * `to_encoded_point(true)` yields, rust-k256/src/utils.rs:23-25; the form the JS binding exchanges,
 * javascript/src/lib.rs:97-117) and `00` followed by 32 zero bytes for the identity.
 *   plume_points_compress_batch     n x 64 -> n x 33 (no validation)
 *   plume_points_decompress_batch   n x 33 -> n x 64 and ok[i]: 1 when the slot decodes the way k256's
 *                                   AffinePoint decoding accepts it (prefix, x < p, x^3 + 7 a square), else 0 (point zeroed)
 *   plume_sign_batch_sec1           plume_sign_batch with 33-byte point outputs (r_point33 / hashed_to_curve_r33 may be NULL)
 *   plume_verify_batch_sec1         plume_verify_batch on 33-byte points; a slot that does not decode gives ok = 0
 */
int plume_points_compress_batch(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33);
int plume_points_decompress_batch(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok);
/* device-pointer forms of the two conversions (work enqueued on `stream`) */
int plume_points_compress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in64, uint8_t* out33, void* stream);
int plume_points_decompress_batch_device(plume_ctx* ctx, size_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, void* stream);
int plume_sign_batch_sec1(plume_ctx* ctx, int version, size_t n,
                          const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                          const uint8_t* sk, const uint8_t* r,
                          uint8_t* pk33, uint8_t* nullifier33, uint8_t* c, uint8_t* s,
                          uint8_t* r_point33, uint8_t* hashed_to_curve_r33, uint8_t* status);
int plume_verify_batch_sec1(plume_ctx* ctx, int version, size_t n,
                            const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                            const uint8_t* pk33, const uint8_t* nullifier33, const uint8_t* c, const uint8_t* s,
                            const uint8_t* r_point33, const uint8_t* hashed_to_curve_r33, uint8_t* ok);

/*
 * arkworks flavour: the same path with the semantics of the twin crate `plume_arkworks`.
 *   plume_ark_sign_batch    replaces sign_with_r (rust-arkworks/src/lib.rs:229-278; `sign`, :281-291, is the same after
 *                           r = Fr::rand(rng) on the caller's side).  The public key is an INPUT (the keypair's first half,
 *                           it is not recomputed from sk); r and sk are any Fr, zero included (r = 0 gives identity
 *                           r_point / hashed_to_curve_r, encoded as the single byte 00 in the challenge, :112-118);
 *                           digest_private = SHA-256(...) reduced mod n (:257); s = r + sk * c; nothing is rejected but
 *                           scalars >= n (PLUME_STATUS_BAD_R / BAD_SK) and pk = identity or off-curve (PLUME_STATUS_BAD_PK).
 *                           Outputs map onto PlumeSignaturePublic {s, nullifier} and PlumeSignaturePrivate
 *                           {hashed_to_curve_r, r_point, digest_private} (:175-201).
 *   plume_ark_verify_batch  replaces verify_non_zk (rust-arkworks/src/tests.rs:28-78): r_point and hashed_to_curve_r are
 *                           required and compared in BOTH versions; ok[i] = 1 iff it returns Ok(true); an identity pk
 *                           (Err in the reference) gives 0.
 */
int plume_ark_sign_batch(plume_ctx* ctx, int version, size_t n,
                         const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                         const uint8_t* pk, const uint8_t* sk, const uint8_t* r,
                         uint8_t* nullifier, uint8_t* digest_private, uint8_t* s,
                         uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status);
int plume_ark_verify_batch(plume_ctx* ctx, int version, size_t n,
                           const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                           const uint8_t* pk, const uint8_t* nullifier, const uint8_t* digest_private, const uint8_t* s,
                           const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok);
int plume_ark_sign_batch_device(plume_ctx* ctx, int version, size_t n,
                                const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                const uint8_t* pk, const uint8_t* sk, const uint8_t* r,
                                uint8_t* nullifier, uint8_t* digest_private, uint8_t* s,
                                uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status, void* stream);
int plume_ark_verify_batch_device(plume_ctx* ctx, int version, size_t n,
                                  const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                  const uint8_t* pk, const uint8_t* nullifier, const uint8_t* digest_private, const uint8_t* s,
                                  const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok, void* stream);

/*
 * Circuit-input side.  The circom circuit takes, next to c, s, pk and the nullifier, per-u witness hints
 * for its hash_to_curve component (circuits/circom/verify_nullifier.circom:21-31) that the reference obtains from the
 * external generate_inputs_from_array (circuits/circom/test/v1.test.ts:5,38-40; npm package
 * secp256k1_hash_to_curve_circom, not vendored, no fixture in the reference pins its conventions).  These two calls
 * produce on the device everything those inputs are made from:
 *   plume_hash_to_curve_witness_batch   for each preimage: u0, u1 = hash_to_field (u: n x 2 x 32, big-endian canonical);
 *                                       gx1_square[2i+k] = 1 when the SSWU map of u_k took x1 (g(x1) is a square), 0 when x2;
 *                                       Q0, Q1 = iso_map(map_to_curve(u_k)), the circuit's x_mapped / y_mapped
 *                                       (q: n x 2 x 64); h = Q0 + Q1 (n x 64) = plume_hash_to_curve_batch's output;
 *                                       hints (n x 2 x 3 x 32, may be NULL): per u_k the circuit's gx1_sqrt, gx2_sqrt, y_pos.
 *                                       CONVENTION UNPINNED BY THE REFERENCE, declared here: exactly one of g(x1), g(x2) is
 *                                       a square (Z = -11 is not one); that one's hint is its even square root (RFC 9380
 *                                       sgn0 = 0), the other hint is 0; y_pos is the even square root of g(x) for the x the
 *                                       map takes, so the map's y is y_pos or p - y_pos according to sgn0(u_k).
 *                                       tests/test_circuit_inputs.py checks these against the relations the circuit
 *                                       enforces (gx1_sqrt^2 = g(x1) or gx2_sqrt^2 = g(x2); y_pos^2 = g(x); the sign rule).
 *   plume_registers_batch               n 32-byte big-endian values -> n x 4 little-endian 64-bit registers, least
 *                                       significant first: scalarToCircuitValue / pointToCircuitValue of
 *                                       circuits/circom/utils.ts:11-17,32-51 (a point is its x then its y).
 */
int plume_hash_to_curve_witness_batch(plume_ctx* ctx, size_t n,
                                      const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                      uint8_t* u, uint8_t* q, uint8_t* gx1_square, uint8_t* h, uint8_t* hints);
int plume_registers_batch(plume_ctx* ctx, size_t n, const uint8_t* in32, uint64_t* out4);

/* out[i] = k_i * G for n 32-byte big-endian scalars (taken mod n; 0 gives the identity): public keys from secret keys
 * (`sk.public_key()`, rust-k256/src/randomizedsigner.rs:53), and the public-key field of the SEC1-DER scalars of the JS
 * wire form (`SecretKey::from(s).to_sec1_der()`, javascript/src/lib.rs:97-117 -- the host mirrors' sec1_der helpers). */
int plume_fixed_base_mul_batch(plume_ctx* ctx, size_t n, const uint8_t* scalars, uint8_t* out);

/* Device-pointer variants: all pointers are device memory of the context's GPU, `stream` is a
 * cudaStream_t (passed as void* to keep CUDA headers out of this file).  n must not exceed
 * plume_ctx_chunk_items().  Work is enqueued; the caller synchronises the stream. */
int plume_sign_batch_device(plume_ctx* ctx, int version, size_t n,
                            const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                            const uint8_t* sk, const uint8_t* r,
                            uint8_t* pk, uint8_t* nullifier, uint8_t* c, uint8_t* s,
                            uint8_t* r_point, uint8_t* hashed_to_curve_r, uint8_t* status, void* stream);
int plume_verify_batch_device(plume_ctx* ctx, int version, size_t n,
                              const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                              const uint8_t* pk, const uint8_t* nullifier, const uint8_t* c, const uint8_t* s,
                              const uint8_t* r_point, const uint8_t* hashed_to_curve_r, uint8_t* ok, void* stream);
int plume_hash_to_curve_batch_device(plume_ctx* ctx, size_t n,
                                     const uint8_t* msgs, const uint64_t* msg_offsets, size_t msg_len,
                                     uint8_t* out, void* stream);

/* Number of kernel launches this context has enqueued so far (bench.py's `gpu_launches`). */
uint64_t plume_ctx_launch_count(const plume_ctx* ctx);

/* Per-stage device timing, measured with CUDA events on the launching stream.  After
 * plume_ctx_set_profiling(ctx, 1) every stage launch is bracketed by an event pair;
 * plume_ctx_stage_ms returns the summed duration in milliseconds of all launches of the named
 * stage since profiling was switched on (and their number in *launches), synchronising the
 * context's streams first.  Stage names: "sign_fixed", "sign_h2c", "sign_varbase", "sign_final",
 * "verify_h2c", "verify_mul_a" (G*s - pk*c), "verify_tab_b" + "verify_mul_b" (window tables and ladder of h*s - nul*c), "verify_final", "h2c_map", "h2c_out", "binv",
 * "sec1_compress", "sec1_decompress", "h2c_witness", "registers", "fixed_mul".  On a multi-device context: summed over its devices.
 * Returns a negative value for an unknown stage.  set_profiling(ctx, 1) also resets the sums. */
int plume_ctx_set_profiling(plume_ctx* ctx, int on);
double plume_ctx_stage_ms(plume_ctx* ctx, const char* stage, uint64_t* launches);

/* Integer-pipe microbenchmark, the denominator of the integer-ALU roofline (SURVEY.md section 8d "Peak"):
 * the sustained rate of 32x32->64-bit multiply-accumulates (limb products per second) over every SM, in two
 * forms -- plain IMAD.WIDE.U32 with a 64-bit addend, and the carry-chain form (IMAD.WIDE.U32.X rows fused from
 * mad.lo.cc / madc.hi.cc) the shipped multiplier is built from.  plume_measure_imad_peak returns the faster
 * of the two; plume_measure_imad_rates returns both (either pointer may be NULL). */
int plume_measure_imad_peak(plume_ctx* ctx, int iters, double* limb_products_per_s);
int plume_measure_imad_rates(plume_ctx* ctx, int iters, double* plain_limb_products_per_s, double* carry_limb_products_per_s);

/* Test hook: element-wise base-field operation on raw limb arrays (n x 8 little-endian 32-bit limbs,
 * any representative below 2^256; host pointers).  op: 0 a*b, 1 a^2, 2 a+b, 3 a-b, 4 1/a, 5 canonical(a),
 * 6 a*b[0] (small), 7 a^((p-3)/4), 8 -a, 9 a^((p+1)/4), 10 a == 0 (mod p), 11 8a, 12 2a, 13 4a, 14 1/a by division steps (0 for 0).  Results are weakly reduced
 * (compare modulo p).  Exists so that the carry-chain assembly of the field layer can be checked in
 * isolation on the GPU (tests/test_gpu_field.py); not part of the reference's interface. */
int plume_debug_fe_op(plume_ctx* ctx, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);

/* Test hook: copies of lane `lane`'s (0 or 1) device I/O arena and pinned staging arena after the context's streams have
 * drained (host pointers of `cap` bytes each; either may be NULL), so that tests can check that no secret key or nonce
 * of a finished call survives in the library's own memory. */
int plume_debug_read_arena(plume_ctx* ctx, int lane, uint8_t* dev_copy, uint8_t* host_copy, size_t cap, size_t* dev_bytes,
                           size_t* host_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PLUME_B200_H */
