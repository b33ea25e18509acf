// Known-answer test with a fixed-nonce RNG (same vector as tests/golden/reference_vectors.json "sign_kat").
// UNTESTED: no Rust toolchain in the build image.
use hex_literal::hex;
use plume_b200_shim::{PlumeSignature, SecretKey};

struct FixedNonce([u8; 32]);
impl rand_core::RngCore for FixedNonce {
    fn next_u32(&mut self) -> u32 { unreachable!() }
    fn next_u64(&mut self) -> u64 { unreachable!() }
    fn fill_bytes(&mut self, dest: &mut [u8]) { dest.copy_from_slice(&self.0) }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> { self.fill_bytes(dest); Ok(()) }
}
impl rand_core::CryptoRng for FixedNonce {}

#[test]
fn sign_then_verify_matches_the_golden_vector() {
    let sk = SecretKey::from_bytes(&hex!("519b423d715f8b581f4fa8ee59f4771a5b44c8130b4e3eacca54a56dda72b464").into()).unwrap();
    let mut rng = FixedNonce(hex!("93b9323b629f251b8f3fc2dd11f4672c5544e8230d493eceea98a90bda789808"));
    let v1 = PlumeSignature::sign_v1(&sk, b"An example app message string", &mut rng);
    assert_eq!(v1.c.to_bytes().as_slice(), hex!("c6a7fc2c926ddbaf20731a479fb6566f2daa5514baae5223fe3b32edbce83254"));
    assert_eq!(v1.s.to_bytes().as_slice(), hex!("e69f027d84cb6fe5f761e333d12e975fb190d163e8ea132d7de0bd6079ba28ca"));
    assert!(v1.verify());
    let v2 = PlumeSignature::sign_v2(&sk, b"An example app message string", &mut rng);
    assert_eq!(v2.c.to_bytes().as_slice(), hex!("3dbfb717705010d4f44a70720c95e74b475bd3a783ab0b9e8a6b3b363434eb96"));
    assert_eq!(v2.s.to_bytes().as_slice(), hex!("528e8fbb6452f82200797b1a73b2947a92524bd611085a920f1177cb8098136b"));
    assert!(v2.verify());
}
