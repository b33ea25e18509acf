// sha256.cuh -- SHA-256 compression and a small byte-stream front end.
//
// Replaces sha2::Sha256 as used by the reference for the Fiat-Shamir challenge
// (rust-k256/src/lib.rs:159-168, rust-k256/src/randomizedsigner.rs:73-89) and inside
// expand_message_xmd (k256 hash2curve; written out at rust-arkworks/src/fixed_hasher/expander.rs:89-135).
#pragma once
#include "ptx.cuh"

PLUME_DEV uint32_t sha256_k(int i) {
    const uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
        0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
        0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
        0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
        0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
        0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
        0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    return K[i];
}

PLUME_DEV void sha256_init(uint32_t* st) {
    st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
    st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}
// state after absorbing one all-zero 64-byte block (the Z_pad of expand_message_xmd)
PLUME_DEV void sha256_init_after_zero_block(uint32_t* st) {
    st[0] = 0xda5698beu; st[1] = 0x17b9b469u; st[2] = 0x62335799u; st[3] = 0x779fbecau;
    st[4] = 0x8ce5d491u; st[5] = 0xc0d26243u; st[6] = 0xbafef9eau; st[7] = 0x1837a9d8u;
}

// one compression; w = 16 big-endian message words (clobbered)
PLUME_DEV void sha256_compress(uint32_t* st, uint32_t* w) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + sha256_k(i) + w[i & 15];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// Byte-stream hasher over a small scratch buffer that lives in local memory.  Used for the
// variable-layout parts (message tails of arbitrary length, identity points that encode to one
// byte); the fixed-layout fast paths assemble their words in registers instead.
struct sha256_stream {
    uint32_t st[8];
    uint8_t buf[64];
    uint32_t fill;     // bytes in buf
    uint64_t total;    // bytes absorbed so far (including whole blocks fed directly)
};
PLUME_DEV void sha256_stream_flush(sha256_stream& s) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)s.buf[4 * i] << 24) | ((uint32_t)s.buf[4 * i + 1] << 16) | ((uint32_t)s.buf[4 * i + 2] << 8) | s.buf[4 * i + 3];
    sha256_compress(s.st, w);
    s.fill = 0;
}
PLUME_DEV void sha256_stream_byte(sha256_stream& s, uint8_t b) {
    s.buf[s.fill++] = b;
    s.total++;
    if (s.fill == 64) sha256_stream_flush(s);
}
PLUME_DEV void sha256_stream_bytes(sha256_stream& s, const uint8_t* p, uint32_t n) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) sha256_stream_byte(s, p[i]);
}
// 4 big-endian bytes of a word
PLUME_DEV void sha256_stream_word(sha256_stream& s, uint32_t w) {
    sha256_stream_byte(s, (uint8_t)(w >> 24));
    sha256_stream_byte(s, (uint8_t)(w >> 16));
    sha256_stream_byte(s, (uint8_t)(w >> 8));
    sha256_stream_byte(s, (uint8_t)w);
}
PLUME_DEV void sha256_stream_final(sha256_stream& s, uint32_t* digest_words) {
    uint64_t bits = s.total * 8;
    s.buf[s.fill++] = 0x80;
    if (s.fill > 56) {
        while (s.fill < 64) s.buf[s.fill++] = 0;
        sha256_stream_flush(s);
    }
    while (s.fill < 56) s.buf[s.fill++] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_stream_flush(s);
#pragma unroll
    for (int i = 0; i < 8; i++) digest_words[i] = s.st[i];
}
