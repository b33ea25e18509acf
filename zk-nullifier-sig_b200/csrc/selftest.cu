// plume_self_test: known-answer test of a context through the library's own public batch calls.
//
// The vectors are the ones the reference's tests pin (they are data, restated as hex): the message, secret key and nonce
// of rust-k256/tests/signing.rs:9-13 with the V1 / V2 `c` and `s` it asserts (:15-21), the intermediate points of
// rust-arkworks/src/tests.rs:191-262 (pk, g^r, h, h^r, h^sk) and hash_to_curve("abc") of rust-k256/tests/verification.rs:288-292.
// A maintainer's binding calls this once after plume_ctx_create: a GPU, driver or build that computes anything else than
// the reference does is caught before the first real batch, without the test-side oracle.  Host code only.
#include <cstring>
#include <string>
#include "../../include/plume_b200.h"
#include "ctx.h"

namespace {

void unhex(const char* h, uint8_t* out, size_t n) {
    auto nib = [](char ch) -> int { return ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10; };
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)(nib(h[2 * i]) << 4 | nib(h[2 * i + 1]));
}

struct Kat {
    const char* msg = "An example app message string";
    const char* sk = "519b423d715f8b581f4fa8ee59f4771a5b44c8130b4e3eacca54a56dda72b464";
    const char* r = "93b9323b629f251b8f3fc2dd11f4672c5544e8230d493eceea98a90bda789808";
    const char* c[2] = {"c6a7fc2c926ddbaf20731a479fb6566f2daa5514baae5223fe3b32edbce83254",
                        "3dbfb717705010d4f44a70720c95e74b475bd3a783ab0b9e8a6b3b363434eb96"};
    const char* s[2] = {"e69f027d84cb6fe5f761e333d12e975fb190d163e8ea132d7de0bd6079ba28ca",
                        "528e8fbb6452f82200797b1a73b2947a92524bd611085a920f1177cb8098136b"};
    const char* pk = "0cec028ee08d09e02672a68310814354f9eabfff0de6dacc1cd3a774496076ae"
                     "eff471fba0409897b6a48e8801ad12f95d0009b753cf8f51c128bf6b0bd27fbd";
    const char* g_r = "9d8ca4350e7e2ad27abc6d2a281365818076662962a28429590e2dc736fe9804"
                      "ff08c30b8afd4e854623c835d9c3aac6bcebe45112472d9b9054816a7670c5a1";
    const char* h_r = "6d017c6f63c59fa7a5b1e9a654e27d2869579f4d152131db270558fccd27b97c"
                      "586c43fb5c99818c564a8f80a88a65f83e3f44d3c6caf5a1a4e290b777ac56ed";
    const char* h_sk = "57bc3ed28172ef8adde4b9e0c2cce745fcc5a66473a45c1e626f1d0c67e55830"
                       "6a2f41488d58f33ae46edd2188e111609f9f3ae67ea38fa891d6087fe59ecb73";
    const char* h = "bcac2d0e12679f23c218889395abcdc01f2affbc49c54d1136a2190db0800b65"
                    "3bcfb339c974c0e757d348081f90a123b0a91a53e32b3752145d87f0cd70966e";
    const char* h2c_abc = "3377e01eab42db296b512293120c6cee72b6ecf9f9205760bd9ff11fb3cb2c4b"
                          "7f95890f33efebd1044d382a01b1bee0900fb6116f94688d487c6c7b9c8371f6";
};

bool same(const uint8_t* got, const char* want_hex, size_t n) {
    uint8_t want[64];
    unhex(want_hex, want, n);
    return memcmp(got, want, n) == 0;
}

// SELFTEST_N copies of the vector in one batch: the items of a warp, of several warps and of more than one block all have
// to agree, and item 1 of the verification is tampered (a flipped bit of s) and must be the only one rejected.
constexpr size_t SELFTEST_N = 300;

int self_test_single(plume_ctx* ctx) {
    const Kat k;
    const size_t n = SELFTEST_N, mlen = strlen(k.msg);
    std::string msgs;
    for (size_t i = 0; i < n; i++) msgs.append(k.msg, mlen);
    std::string sk(32 * n, 0), r(32 * n, 0);
    for (size_t i = 0; i < n; i++) {
        unhex(k.sk, (uint8_t*)&sk[32 * i], 32);
        unhex(k.r, (uint8_t*)&r[32 * i], 32);
    }
    auto u8 = [](std::string& v) { return (uint8_t*)&v[0]; };
    auto wipe = [&]() { memset(&sk[0], 0, sk.size()); memset(&r[0], 0, r.size()); };
    for (int version = 1; version <= 2; version++) {
        std::string pk(64 * n, 1), nul(64 * n, 1), c(32 * n, 1), s(32 * n, 1), rp(64 * n, 1), hr(64 * n, 1), st(n, 1), ok(n, 7);
        int rc = plume_sign_batch(ctx, version, n, (const uint8_t*)msgs.data(), nullptr, mlen, u8(sk), u8(r), u8(pk), u8(nul), u8(c),
                                  u8(s), u8(rp), u8(hr), u8(st));
        if (rc) { wipe(); return rc; }
        const char* what = nullptr;
        for (size_t i = 0; i < n && !what; i++) {
            if (st[i] != PLUME_STATUS_OK) what = "status";
            else if (!same(u8(pk) + 64 * i, k.pk, 64)) what = "pk";
            else if (!same(u8(nul) + 64 * i, k.h_sk, 64)) what = "nullifier";
            else if (!same(u8(c) + 32 * i, k.c[version - 1], 32)) what = "c";
            else if (!same(u8(s) + 32 * i, k.s[version - 1], 32)) what = "s";
            else if (!same(u8(rp) + 64 * i, k.g_r, 64)) what = "r_point";
            else if (!same(u8(hr) + 64 * i, k.h_r, 64)) what = "hashed_to_curve_r";
        }
        if (what) {
            wipe();
            return ctx_fail(ctx, PLUME_E_SELFTEST, std::string("self test: V") + char('0' + version) + " signature differs from the reference's vector in " + what);
        }
        s[32 + 31] ^= 1;
        rc = plume_verify_batch(ctx, version, n, (const uint8_t*)msgs.data(), nullptr, mlen, u8(pk), u8(nul), u8(c), u8(s),
                                version == 1 ? u8(rp) : nullptr, version == 1 ? u8(hr) : nullptr, u8(ok));
        if (rc) { wipe(); return rc; }
        for (size_t i = 0; i < n; i++)
            if ((uint8_t)ok[i] != (i == 1 ? 0 : 1)) {
                wipe();
                return ctx_fail(ctx, PLUME_E_SELFTEST, std::string("self test: V") + char('0' + version) +
                                (i == 1 ? " verification accepted a tampered signature" : " verification rejected the reference's vector"));
            }
    }
    wipe();
    // hash_to_curve: the reference's own "abc" vector and the signing vector's h = hash_to_curve(m || compress(pk))
    uint8_t out[64], pk33[33], pk64[64];
    if (int rc = plume_hash_to_curve_batch(ctx, 1, (const uint8_t*)"abc", nullptr, 3, out)) return rc;
    if (!same(out, k.h2c_abc, 64)) return ctx_fail(ctx, PLUME_E_SELFTEST, "self test: hash_to_curve(\"abc\") differs from the reference's vector");
    unhex(k.pk, pk64, 64);
    if (int rc = plume_points_compress_batch(ctx, 1, pk64, pk33)) return rc;
    std::string pre(k.msg, mlen);
    pre.append((const char*)pk33, 33);
    if (int rc = plume_hash_to_curve_batch(ctx, 1, (const uint8_t*)pre.data(), nullptr, pre.size(), out)) return rc;
    if (!same(out, k.h, 64)) return ctx_fail(ctx, PLUME_E_SELFTEST, "self test: hash_to_curve(m || pk) differs from the reference's vector");
    return PLUME_OK;
}

}  // namespace

extern "C" int plume_self_test(plume_ctx* ctx) {
    if (!ctx) return PLUME_E_ARG;
    int devices = plume_ctx_device_count(ctx);
    if (devices <= 1) return self_test_single(ctx);
    for (int g = 0; g < devices; g++) {      // every device of a multi-device context computes the vectors itself
        plume_ctx* sub = plume_ctx_sub(ctx, g);
        if (!sub) return ctx_fail(ctx, PLUME_E_ARG, "self test: missing sub-context");
        if (int rc = self_test_single(sub))
            return ctx_fail(ctx, rc, "device " + std::to_string(g) + ": " + plume_last_error(sub));
    }
    return PLUME_OK;
}
