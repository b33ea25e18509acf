//! `plume_rustcrypto`-shaped API whose arithmetic runs in libplume_b200.so (hand-written sm_100a CUDA).
//!
//! UNTESTED SOURCE -- written without a Rust toolchain (none in the build image).  It keeps the public
//! names and semantics of the reference crate (PlumeSignature, PlumeSignatureV1Fields, PlumeSigner,
//! sign_v1 / sign_v2 / verify, DST, the serde derives) and adds the batch calls that make a GPU worthwhile.
//! Conversions are byte copies: the C ABI uses k256's own encodings (32-byte big-endian FieldBytes,
//! affine x || y, 64 zero bytes for the identity).
pub mod ffi;

use k256::elliptic_curve::sec1::{FromEncodedPoint, ToEncodedPoint};
use k256::{AffinePoint as KAffine, EncodedPoint, FieldBytes};
pub use k256::{AffinePoint, NonZeroScalar, SecretKey};
pub use rand_core::CryptoRngCore;
#[cfg(feature = "serde")]
pub use serde::{Deserialize, Serialize};
use std::sync::{Mutex, OnceLock};

/// Hash-to-curve domain separation tag of the PLUME suite.
pub const DST: &[u8] = b"QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_";

/// One GPU's batch engine.  Externally synchronised by the mutex below.
pub struct Engine(*mut ffi::plume_ctx);
unsafe impl Send for Engine {}

impl Engine {
    /// One context over `devices` (one entry: a single-device context; several: the batch calls range-split
    /// over them inside the library).  `window_bits` 0 = the library's default generator table (22 bits, 2 GiB).
    pub fn new(devices: &[i32], window_bits: i32) -> Result<Self, String> {
        let mut h = core::ptr::null_mut();
        let rc = unsafe {
            if devices.len() == 1 {
                ffi::plume_ctx_create(&mut h, devices[0], window_bits)
            } else {
                ffi::plume_ctx_create_multi(&mut h, devices.as_ptr(), devices.len() as i32, window_bits)
            }
        };
        if rc != ffi::PLUME_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::plume_last_error(core::ptr::null())) };
            return Err(msg.to_string_lossy().into_owned());
        }
        Ok(Engine(h))
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::plume_ctx_destroy(self.0) }
    }
}
/// The process-wide engine behind the reference-shaped calls: GPUs from `PLUME_DEVICES` ("0,1,2,3"; default "0"),
/// a 16-bit generator table (32 MiB, built in milliseconds) unless `PLUME_FIXED_WINDOW` says otherwise.
fn engine() -> &'static Mutex<Engine> {
    static E: OnceLock<Mutex<Engine>> = OnceLock::new();
    E.get_or_init(|| {
        let devices: Vec<i32> = std::env::var("PLUME_DEVICES").ok()
            .map(|v| v.split(',').filter_map(|t| t.trim().parse().ok()).collect())
            .filter(|v: &Vec<i32>| !v.is_empty())
            .unwrap_or_else(|| vec![0]);
        let window = if std::env::var_os("PLUME_FIXED_WINDOW").is_some() { 0 } else { 16 };
        let eng = Engine::new(&devices, window).expect("no B200 / libplume_b200 available (there is no CPU fallback)");
        // the reference's known-answer vectors on every device, once per process
        if unsafe { ffi::plume_self_test(eng.0) } != ffi::PLUME_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::plume_last_error(eng.0)) };
            panic!("libplume_b200 self test failed: {}", msg.to_string_lossy());
        }
        Mutex::new(eng)
    })
}

fn point_to_wire(p: &KAffine) -> [u8; 64] {
    let mut out = [0u8; 64];
    let e = p.to_encoded_point(false);
    if let (Some(x), Some(y)) = (e.x(), e.y()) {
        out[..32].copy_from_slice(x);
        out[32..].copy_from_slice(y);
    } // identity stays all-zero
    out
}
fn point_from_wire(b: &[u8; 64]) -> KAffine {
    if b.iter().all(|v| *v == 0) {
        return KAffine::IDENTITY;
    }
    let e = EncodedPoint::from_affine_coordinates(FieldBytes::from_slice(&b[..32]), FieldBytes::from_slice(&b[32..]), false);
    Option::from(KAffine::from_encoded_point(&e)).expect("library returned an off-curve point")
}
fn scalar_from_wire(b: &[u8; 32]) -> NonZeroScalar {
    Option::from(NonZeroScalar::from_repr(*FieldBytes::from_slice(b))).expect("library returned an out-of-range scalar")
}

/// Signature data of PLUME; `v1specific` selects the protocol variant.
#[cfg_attr(feature = "serde", derive(Serialize, Deserialize))]
pub struct PlumeSignature {
    pub message: Vec<u8>,
    pub pk: AffinePoint,
    pub nullifier: AffinePoint,
    pub c: NonZeroScalar,
    pub s: NonZeroScalar,
    pub v1specific: Option<PlumeSignatureV1Fields>,
}
/// The two extra points a V1 signature carries.
#[derive(Debug)]
#[cfg_attr(feature = "serde", derive(Serialize, Deserialize))]
pub struct PlumeSignatureV1Fields {
    pub r_point: AffinePoint,
    pub hashed_to_curve_r: AffinePoint,
}

/// Borrowed secret key plus the variant flag, as in the reference.
pub struct PlumeSigner<'k> {
    secret_key: &'k SecretKey,
    pub v1: bool,
}
impl<'k> PlumeSigner<'k> {
    pub fn new(secret_key: &'k SecretKey, v1: bool) -> Self {
        PlumeSigner { secret_key, v1 }
    }
}
impl<'k> signature::RandomizedSigner<PlumeSignature> for PlumeSigner<'k> {
    fn try_sign_with_rng(&self, rng: &mut impl CryptoRngCore, msg: &[u8]) -> Result<PlumeSignature, signature::Error> {
        // same RNG contract as the reference: one SecretKey::random draw per signature
        let r = SecretKey::random(rng);
        let mut out = sign_batch(self.v1, &[msg], &[self.secret_key.to_bytes().into()], &[r.to_bytes().into()]);
        Ok(out.pop().unwrap().unwrap_or_else(|st| panic!("{}", status_text(st))))
    }
}

fn status_text(st: u8) -> &'static str {
    match st {
        ffi::PLUME_STATUS_BAD_C => "it should be impossible to get the hash equal to zero",
        ffi::PLUME_STATUS_ZERO_S => "something is terribly wrong if the nonce is equal to negated product of the secret and the hash",
        ffi::PLUME_STATUS_H_INF => "something is drammatically wrong if the input hashed to the identity",
        ffi::PLUME_STATUS_BAD_R => "nonce out of range",
        _ => "secret key out of range",
    }
}

/// Batch signing: item i signs `msgs[i]` with `sk[i]` and nonce `r[i]` (32-byte big-endian each).
/// `Err(status)` marks the items where the reference would have panicked.
pub fn sign_batch(v1: bool, msgs: &[&[u8]], sk: &[[u8; 32]], r: &[[u8; 32]]) -> Vec<Result<PlumeSignature, u8>> {
    let n = msgs.len();
    assert!(sk.len() == n && r.len() == n);
    let mut blob = Vec::new();
    let mut offs = Vec::with_capacity(n + 1);
    offs.push(0u64);
    for m in msgs {
        blob.extend_from_slice(m);
        offs.push(blob.len() as u64);
    }
    let (mut pk, mut nul, mut rp, mut hr) = (vec![0u8; 64 * n], vec![0u8; 64 * n], vec![0u8; 64 * n], vec![0u8; 64 * n]);
    let (mut c, mut s, mut st) = (vec![0u8; 32 * n], vec![0u8; 32 * n], vec![0u8; n]);
    let eng = engine().lock().unwrap();
    let rc = unsafe {
        ffi::plume_sign_batch(eng.0, if v1 { 1 } else { 2 }, n, blob.as_ptr(), offs.as_ptr(), 0, sk.as_ptr() as *const u8,
            r.as_ptr() as *const u8, pk.as_mut_ptr(), nul.as_mut_ptr(), c.as_mut_ptr(), s.as_mut_ptr(), rp.as_mut_ptr(),
            hr.as_mut_ptr(), st.as_mut_ptr())
    };
    assert_eq!(rc, ffi::PLUME_OK, "plume_sign_batch failed");
    (0..n).map(|i| {
        if st[i] != ffi::PLUME_STATUS_OK {
            return Err(st[i]);
        }
        let w64 = |v: &Vec<u8>| -> [u8; 64] { v[64 * i..64 * i + 64].try_into().unwrap() };
        let w32 = |v: &Vec<u8>| -> [u8; 32] { v[32 * i..32 * i + 32].try_into().unwrap() };
        Ok(PlumeSignature {
            message: msgs[i].to_vec(),
            pk: point_from_wire(&w64(&pk)),
            nullifier: point_from_wire(&w64(&nul)),
            c: scalar_from_wire(&w32(&c)),
            s: scalar_from_wire(&w32(&s)),
            v1specific: v1.then(|| PlumeSignatureV1Fields { r_point: point_from_wire(&w64(&rp)), hashed_to_curve_r: point_from_wire(&w64(&hr)) }),
        })
    }).collect()
}

/// Batch verification of signatures of one variant (all V1 or all V2).
pub fn verify_batch(sigs: &[&PlumeSignature]) -> Vec<bool> {
    let n = sigs.len();
    if n == 0 {
        return vec![];
    }
    let v1 = sigs[0].v1specific.is_some();
    assert!(sigs.iter().all(|s| s.v1specific.is_some() == v1), "mixed variants in one batch");
    let mut blob = Vec::new();
    let mut offs = vec![0u64];
    let (mut pk, mut nul, mut rp, mut hr, mut c, mut s) = (vec![], vec![], vec![], vec![], vec![], vec![]);
    for sig in sigs {
        blob.extend_from_slice(&sig.message);
        offs.push(blob.len() as u64);
        pk.extend_from_slice(&point_to_wire(&sig.pk));
        nul.extend_from_slice(&point_to_wire(&sig.nullifier));
        c.extend_from_slice(&sig.c.to_bytes());
        s.extend_from_slice(&sig.s.to_bytes());
        if let Some(f) = &sig.v1specific {
            rp.extend_from_slice(&point_to_wire(&f.r_point));
            hr.extend_from_slice(&point_to_wire(&f.hashed_to_curve_r));
        }
    }
    let mut ok = vec![0u8; n];
    let eng = engine().lock().unwrap();
    let rc = unsafe {
        ffi::plume_verify_batch(eng.0, if v1 { 1 } else { 2 }, n, blob.as_ptr(), offs.as_ptr(), 0, pk.as_ptr(), nul.as_ptr(), c.as_ptr(),
            s.as_ptr(), if v1 { rp.as_ptr() } else { core::ptr::null() }, if v1 { hr.as_ptr() } else { core::ptr::null() }, ok.as_mut_ptr())
    };
    assert_eq!(rc, ffi::PLUME_OK, "plume_verify_batch failed");
    ok.into_iter().map(|b| b != 0).collect()
}

impl PlumeSignature {
    /// Checks the two DLEQ equations and the challenge hash; `true` iff the signature is valid.
    pub fn verify(&self) -> bool {
        verify_batch(&[self])[0]
    }
    /// Variant 1 (carries g^r and h^r).
    pub fn sign_v1(secret_key: &SecretKey, msg: &[u8], rng: &mut impl CryptoRngCore) -> Self {
        use signature::RandomizedSigner;
        PlumeSigner::new(secret_key, true).sign_with_rng(rng, msg)
    }
    /// Variant 2 (challenge over nullifier, g^r, h^r only).
    pub fn sign_v2(secret_key: &SecretKey, msg: &[u8], rng: &mut impl CryptoRngCore) -> Self {
        use signature::RandomizedSigner;
        PlumeSigner::new(secret_key, false).sign_with_rng(rng, msg)
    }
}
