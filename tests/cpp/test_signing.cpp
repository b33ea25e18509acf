// The reference's integration test (rust-k256/tests/signing.rs:23-64) restated against the C++ host mirror:
// a mock RNG whose fill_bytes hands out the fixed nonce, sign_v1 / sign_v2, then verify, plus a small batch.
//   usage: test_signing <msg-ascii> <sk-hex> <r-hex>   -> prints "v1 <c> <s> <verify>", "v2 ...", "batch <n_ok>"
#include <cstdio>
#include <string>
#include "plume.hpp"

static plume::Bytes32 unhex(const std::string& h) {
    plume::Bytes32 b{};
    for (int i = 0; i < 32; i++) b[i] = (uint8_t)std::stoi(h.substr(2 * i, 2), nullptr, 16);
    return b;
}
static std::string hex(const plume::Bytes32& b) {
    char buf[65];
    for (int i = 0; i < 32; i++) snprintf(buf + 2 * i, 3, "%02x", b[i]);
    return buf;
}
struct Mock {
    plume::Bytes32 r;
    void fill_bytes(uint8_t* dest, size_t n) {
        if (n != 32) throw std::runtime_error("mock rng: exactly one 32-byte fill per signature");
        for (size_t i = 0; i < n; i++) dest[i] = r[i];
    }
};

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    std::string m = argv[1];
    std::vector<uint8_t> msg(m.begin(), m.end());
    auto sk = plume::SecretKey::from_bytes(unhex(argv[2]));
    Mock rng{unhex(argv[3])};
    auto s1 = plume::PlumeSignature::sign_v1(sk, msg, rng);
    printf("v1 %s %s %d %d\n", hex(s1.c).c_str(), hex(s1.s).c_str(), (int)s1.verify(), (int)s1.v1specific.has_value());
    auto s2 = plume::PlumeSignature::sign_v2(sk, msg, rng);
    printf("v2 %s %s %d %d\n", hex(s2.c).c_str(), hex(s2.s).c_str(), (int)s2.verify(), (int)s2.v1specific.has_value());
    printf("json1 %s\n", s1.to_json().c_str());
    printf("json2 %s\n", s2.to_json().c_str());
    s2.s[31] ^= 1;
    printf("tampered %d\n", (int)s2.verify());
    // out-of-range secret key is rejected the way SecretKey::from_bytes does
    try { plume::SecretKey::from_bytes(plume::Bytes32{}); printf("zero-sk accepted\n"); } catch (const plume::Error&) { printf("zero-sk rejected\n"); }
    // a batch through the same context
    const size_t n = 1000;
    std::vector<uint8_t> msgs(n * 32), sks(n * 32), rs(n * 32);
    for (size_t i = 0; i < n * 32; i++) { msgs[i] = (uint8_t)(i * 131 + 7); sks[i] = (uint8_t)(i * 31 + 1); rs[i] = (uint8_t)(i * 17 + 3); }
    for (size_t i = 0; i < n; i++) { sks[32 * i] &= 0x7F; rs[32 * i] &= 0x7F; }
    auto cx = plume::Context::global();
    auto out = plume::sign_batch(*cx, 1, n, msgs.data(), 32, sks.data(), rs.data());
    auto ok = plume::verify_batch(*cx, 1, n, msgs.data(), 32, out);
    size_t good = 0;
    for (size_t i = 0; i < n; i++) good += (out.status[i] == 0 && ok[i] == 1);
    printf("batch %zu\n", good);
    // the arkworks twin on the same vector (rust-arkworks/src/tests.rs:267-300): pk is an input here
    for (auto ver : {plume::ark::PlumeVersion::V1, plume::ark::PlumeVersion::V2}) {
        auto sig = plume::ark::sign_with_r(s1.pk, sk.to_bytes(), msg, unhex(argv[3]), ver);
        bool good_sig = plume::ark::verify_non_zk(sig.first, sig.second, s1.pk, msg, ver);
        auto bad = sig.first; bad.s[31] ^= 1;
        printf("ark v%d %s %s %d %d\n", (int)ver, hex(sig.second.digest_private).c_str(), hex(sig.first.s).c_str(), (int)good_sig,
               (int)plume::ark::verify_non_zk(bad, sig.second, s1.pk, msg, ver));
    }
    try { plume::ark::sign_with_r(plume::AffinePoint{}, sk.to_bytes(), msg, unhex(argv[3]), plume::ark::PlumeVersion::V2); printf("identity-pk accepted\n"); }
    catch (const plume::ark::HashToCurveError&) { printf("identity-pk rejected\n"); }
    return 0;
}
