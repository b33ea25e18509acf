// plume.hpp -- header-only C++ host mirror of the reference's operator interface for the PLUME hot path,
// on top of the C ABI (include/plume_b200.h).  The reference is the Rust crate `plume_rustcrypto`
// (/root/reference/rust-k256); there is no Rust toolchain in the build image, so this is the compiled-language
// host side a C++ caller uses, with the same names, argument meaning and error behaviour:
//
//   PlumeSignature{message, pk, nullifier, c, s, v1specific}      rust-k256/src/lib.rs:67-80
//   PlumeSignatureV1Fields{r_point, hashed_to_curve_r}            rust-k256/src/lib.rs:84-89
//   PlumeSignature::verify / sign_v1 / sign_v2                    rust-k256/src/lib.rs:93,149,154
//   PlumeSigner{secret_key, v1} + try_sign_with_rng               rust-k256/src/randomizedsigner.rs:25-43
//   SecretKey::random(rng): 32 bytes from rng.fill_bytes, big-endian, rejection-sampled into [1, n-1]
//                                                                 (pinned by the mock RNG of tests/signing.rs:23-44)
//
// Where the reference panics (`expect`, randomizedsigner.rs:61,91,95) this throws plume::Panic with the same text.
// Single signatures are batches of one; sign_batch / verify_batch are the additions that make the GPU worth it.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "plume_b200.h"

namespace plume {

inline constexpr char DST[] = "QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_";  // rust-k256/src/lib.rs:61

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };

using Bytes32 = std::array<uint8_t, 32>;

// k256::AffinePoint as it crosses the ABI: x || y big-endian; all zero = identity
struct AffinePoint {
    std::array<uint8_t, 64> xy{};
    bool is_identity() const { for (uint8_t b : xy) if (b) return false; return true; }
    bool operator==(const AffinePoint& o) const { return xy == o.xy; }
};
using NonZeroScalar = Bytes32;  // big-endian, in [1, n-1]

inline bool scalar_in_range(const Bytes32& k) {
    static const uint8_t N[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE,
                                  0xBA, 0xAE, 0xDC, 0xE6, 0xAF, 0x48, 0xA0, 0x3B, 0xBF, 0xD2, 0x5E, 0x8C, 0xD0, 0x36, 0x41, 0x41};
    bool zero = true;
    for (uint8_t b : k) zero &= (b == 0);
    return !zero && std::memcmp(k.data(), N, 32) < 0;
}

// one GPU's signer/verifier; shared_ptr so signatures can keep the context they were made with
class Context {
  public:
    explicit Context(int device = 0, int fixed_window_bits = 0) {
        if (plume_ctx_create(&h_, device, fixed_window_bits) != PLUME_OK) throw Error(std::string("plume_ctx_create: ") + plume_last_error(nullptr));
    }
    // one context over several GPUs: the batch calls range-split over them (plume_ctx_create_multi)
    explicit Context(const std::vector<int>& devices, int fixed_window_bits = 0) {
        if (plume_ctx_create_multi(&h_, devices.data(), (int)devices.size(), fixed_window_bits) != PLUME_OK)
            throw Error(std::string("plume_ctx_create_multi: ") + plume_last_error(nullptr));
    }
    ~Context() { plume_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    plume_ctx* get() const { return h_; }
    void check(int rc, const char* what) const { if (rc != PLUME_OK) throw Error(std::string(what) + ": " + plume_last_error(h_)); }
    // the reference's known-answer vectors through this context (plume_self_test); throws naming the field that differs
    void self_test() const { check(plume_self_test(h_), "plume_self_test"); }
    // the context of the single-signature calls: small generator table (16-bit windows, 64 MiB), workspaces grow on demand;
    // self-tested once when it is created
    static std::shared_ptr<Context> global() {
        static std::shared_ptr<Context> g = [] { auto c = std::make_shared<Context>(0, 16); c->self_test(); return c; }();
        return g;
    }
  private:
    plume_ctx* h_ = nullptr;
};

class SecretKey {
  public:
    static SecretKey from_bytes(const Bytes32& b) {
        if (!scalar_in_range(b)) throw Error("secret key out of range");
        return SecretKey(b);
    }
    template <class Rng>
    static SecretKey random(Rng& rng) {
        for (;;) {
            Bytes32 b;
            rng.fill_bytes(b.data(), b.size());
            if (scalar_in_range(b)) return SecretKey(b);
        }
    }
    const Bytes32& to_bytes() const { return k_; }
  private:
    explicit SecretKey(const Bytes32& b) : k_(b) {}
    Bytes32 k_;
};

struct PlumeSignatureV1Fields {
    AffinePoint r_point;
    AffinePoint hashed_to_curve_r;
};

inline const char* status_text(uint8_t st) {
    switch (st) {
        case PLUME_STATUS_BAD_R: return "nonce r outside [1, n-1]";
        case PLUME_STATUS_BAD_SK: return "secret key outside [1, n-1]";
        case PLUME_STATUS_BAD_C: return "it should be impossible to get the hash equal to zero";
        case PLUME_STATUS_ZERO_S: return "something is terribly wrong if the nonce is equal to negated product of the secret and the hash";
        case PLUME_STATUS_H_INF: return "something is drammatically wrong if the input hashed to the identity";
        default: return "unknown status";
    }
}

struct PlumeSignature {
    std::vector<uint8_t> message;
    AffinePoint pk;
    AffinePoint nullifier;
    NonZeroScalar c{};
    NonZeroScalar s{};
    std::optional<PlumeSignatureV1Fields> v1specific;
    std::shared_ptr<Context> ctx;

    // rust-k256/src/lib.rs:93-145
    bool verify() const {
        auto cx = ctx ? ctx : Context::global();
        uint8_t ok = 0;
        const int version = v1specific ? 1 : 2;
        cx->check(plume_verify_batch(cx->get(), version, 1, message.data(), nullptr, message.size(), pk.xy.data(), nullifier.xy.data(), c.data(),
                                     s.data(), v1specific ? v1specific->r_point.xy.data() : nullptr,
                                     v1specific ? v1specific->hashed_to_curve_r.xy.data() : nullptr, &ok),
                  "plume_verify_batch");
        return ok != 0;
    }
    // The serde_json form of the derive at rust-k256/src/lib.rs:66,83 (field encodings per k256 0.13: upper-case hex of the
    // SEC1 compressed point / of the 32 scalar bytes; message as an array of numbers; unpinned by the reference -- the same
    // text as the Python mirror's PlumeSignature.to_json, which tests/ compare).
    std::string to_json() const {
        auto hexu = [](const uint8_t* p, size_t n) { static const char* d = "0123456789ABCDEF"; std::string o; for (size_t i = 0; i < n; i++) { o += d[p[i] >> 4]; o += d[p[i] & 15]; } return o; };
        auto pt = [&](const AffinePoint& a) {
            if (a.is_identity()) return std::string("00");
            uint8_t e[33]; e[0] = (uint8_t)(2 + (a.xy[63] & 1)); std::memcpy(e + 1, a.xy.data(), 32);
            return hexu(e, 33);
        };
        std::string o = "{\"message\":[";
        for (size_t i = 0; i < message.size(); i++) { if (i) o += ","; o += std::to_string((unsigned)message[i]); }
        o += "],\"pk\":\"" + pt(pk) + "\",\"nullifier\":\"" + pt(nullifier) + "\",\"c\":\"" + hexu(c.data(), 32) + "\",\"s\":\"" + hexu(s.data(), 32) + "\",\"v1specific\":";
        if (v1specific) o += "{\"r_point\":\"" + pt(v1specific->r_point) + "\",\"hashed_to_curve_r\":\"" + pt(v1specific->hashed_to_curve_r) + "\"}";
        else o += "null";
        return o + "}";
    }
    template <class Rng> static PlumeSignature sign_v1(const SecretKey& sk, const std::vector<uint8_t>& msg, Rng& rng, std::shared_ptr<Context> cx = nullptr);
    template <class Rng> static PlumeSignature sign_v2(const SecretKey& sk, const std::vector<uint8_t>& msg, Rng& rng, std::shared_ptr<Context> cx = nullptr);
};

// rust-k256/src/randomizedsigner.rs:25-41
class PlumeSigner {
  public:
    PlumeSigner(const SecretKey& secret_key, bool v1, std::shared_ptr<Context> cx = nullptr) : secret_key_(secret_key), v1(v1), ctx_(std::move(cx)) {}
    bool v1;
    // randomizedsigner.rs:43-112 as a batch of one
    template <class Rng>
    PlumeSignature try_sign_with_rng(Rng& rng, const std::vector<uint8_t>& msg) const {
        SecretKey r = SecretKey::random(rng);  // :49
        auto cx = ctx_ ? ctx_ : Context::global();
        PlumeSignature sig;
        PlumeSignatureV1Fields f;
        uint8_t st = 0;
        cx->check(plume_sign_batch(cx->get(), v1 ? 1 : 2, 1, msg.data(), nullptr, msg.size(), secret_key_.to_bytes().data(), r.to_bytes().data(),
                                   sig.pk.xy.data(), sig.nullifier.xy.data(), sig.c.data(), sig.s.data(), f.r_point.xy.data(),
                                   f.hashed_to_curve_r.xy.data(), &st),
                  "plume_sign_batch");
        if (st != PLUME_STATUS_OK) throw Panic(status_text(st));
        sig.message = msg;
        if (v1) sig.v1specific = f;
        sig.ctx = cx;
        return sig;
    }
    template <class Rng> PlumeSignature sign_with_rng(Rng& rng, const std::vector<uint8_t>& msg) const { return try_sign_with_rng(rng, msg); }
  private:
    const SecretKey& secret_key_;
    std::shared_ptr<Context> ctx_;
};

template <class Rng>
PlumeSignature PlumeSignature::sign_v1(const SecretKey& sk, const std::vector<uint8_t>& msg, Rng& rng, std::shared_ptr<Context> cx) {
    return PlumeSigner(sk, true, std::move(cx)).sign_with_rng(rng, msg);   // rust-k256/src/lib.rs:149-151
}
template <class Rng>
PlumeSignature PlumeSignature::sign_v2(const SecretKey& sk, const std::vector<uint8_t>& msg, Rng& rng, std::shared_ptr<Context> cx) {
    return PlumeSigner(sk, false, std::move(cx)).sign_with_rng(rng, msg);  // rust-k256/src/lib.rs:154-156
}

// ---- batch entry points (SoA, see plume_b200.h) ----------------------------------------------------------------
struct SignBatchOut {
    std::vector<uint8_t> pk, nullifier, c, s, r_point, hashed_to_curve_r, status;
};
// fixed-length messages (msg_len bytes each, n of them); sk, r: n x 32
inline SignBatchOut sign_batch(Context& cx, int version, size_t n, const uint8_t* msgs, size_t msg_len, const uint8_t* sk, const uint8_t* r) {
    SignBatchOut o;
    o.pk.resize(n * 64); o.nullifier.resize(n * 64); o.c.resize(n * 32); o.s.resize(n * 32);
    o.r_point.resize(n * 64); o.hashed_to_curve_r.resize(n * 64); o.status.resize(n);
    cx.check(plume_sign_batch(cx.get(), version, n, msgs, nullptr, msg_len, sk, r, o.pk.data(), o.nullifier.data(), o.c.data(), o.s.data(),
                              o.r_point.data(), o.hashed_to_curve_r.data(), o.status.data()),
             "plume_sign_batch");
    return o;
}
inline std::vector<uint8_t> verify_batch(Context& cx, int version, size_t n, const uint8_t* msgs, size_t msg_len, const SignBatchOut& sig) {
    std::vector<uint8_t> ok(n);
    cx.check(plume_verify_batch(cx.get(), version, n, msgs, nullptr, msg_len, sig.pk.data(), sig.nullifier.data(), sig.c.data(), sig.s.data(),
                                sig.r_point.data(), sig.hashed_to_curve_r.data(), ok.data()),
             "plume_verify_batch");
    return ok;
}

// ---- the arkworks twin (rust-arkworks/src/lib.rs), SURVEY.md 8f-3 -----------------------------------------------
namespace ark {
enum class PlumeVersion { V1 = 1, V2 = 2 };                       // lib.rs:65-69
using Fr = Bytes32;                                                // big-endian, in [0, n)
struct HashToCurveError : Error { using Error::Error; };           // hash_to_curve on the identity pk (lib.rs:97-100)
struct PlumeSignaturePublic {                                      // lib.rs:175-183
    std::vector<uint8_t> message;
    Fr s{};
    AffinePoint nullifier;
    std::optional<PlumeVersion> variant;
};
struct PlumeSignaturePrivate {                                     // lib.rs:185-194
    AffinePoint hashed_to_curve_r, r_point;
    Fr digest_private{};
    PlumeVersion variant{PlumeVersion::V1};
};
// lib.rs:229-278: keypair = (pk, sk); nothing is rejected but an identity / off-curve pk and scalars >= n
inline std::pair<PlumeSignaturePublic, PlumeSignaturePrivate> sign_with_r(const AffinePoint& pk, const Fr& sk, const std::vector<uint8_t>& message,
                                                                          const Fr& r_scalar, PlumeVersion version,
                                                                          std::shared_ptr<Context> cx = nullptr) {
    if (!cx) cx = Context::global();
    PlumeSignaturePublic pub;
    PlumeSignaturePrivate priv;
    uint8_t st = 0;
    cx->check(plume_ark_sign_batch(cx->get(), (int)version, 1, message.data(), nullptr, message.size(), pk.xy.data(), sk.data(), r_scalar.data(),
                                   pub.nullifier.xy.data(), priv.digest_private.data(), pub.s.data(), priv.r_point.xy.data(),
                                   priv.hashed_to_curve_r.xy.data(), &st),
              "plume_ark_sign_batch");
    if (st == PLUME_STATUS_BAD_PK) throw HashToCurveError("`pk` shouldn't be the identity element");
    if (st != PLUME_STATUS_OK) throw Error("scalar is not a canonical Fr");
    pub.message = message;
    pub.variant = version;
    priv.variant = version;
    return {pub, priv};
}
// rust-arkworks/src/tests.rs:28-78
inline bool verify_non_zk(const PlumeSignaturePublic& pub, const PlumeSignaturePrivate& priv, const AffinePoint& pk,
                          const std::vector<uint8_t>& message, PlumeVersion version, std::shared_ptr<Context> cx = nullptr) {
    if (!cx) cx = Context::global();
    if (pk.is_identity()) throw HashToCurveError("`pk` shouldn't be the identity element");
    uint8_t ok = 0;
    cx->check(plume_ark_verify_batch(cx->get(), (int)version, 1, message.data(), nullptr, message.size(), pk.xy.data(), pub.nullifier.xy.data(),
                                     priv.digest_private.data(), pub.s.data(), priv.r_point.xy.data(), priv.hashed_to_curve_r.xy.data(), &ok),
              "plume_ark_verify_batch");
    return ok != 0;
}
}  // namespace ark

}  // namespace plume
