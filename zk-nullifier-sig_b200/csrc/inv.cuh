// inv.cuh -- modular inversion in Fp by division steps (Bernstein-Yang "safegcd", variable time), for the places where ONE
// inversion is on the caller's critical path: the batched inversion of a small batch (stages.cuh binv_body).
//
// Fermat's x^(p-2) (fe.cuh fe_inv) is 255 squarings and 15 multiplications one after the other: ~62 us on one lane of a B200,
// three times per signature and twice per verification of a batch of one.  The division-step iteration works on the low 32
// bits of (f, g) = (p, x) for 30 steps at a time, which gives a 2x2 transition matrix with entries of at most 30 bits, and
// then applies that matrix to the full-size f, g (exact division by 2^30) and to the cofactors d, e (mod p); g reaches 0
// after 17-19 such rounds (at most 741 / 30 = 25 by the theorem), when f = +-1 and d = +-x^-1 2^(30 rounds).  That factor is
// removed once, at the end, with a constant.  About 9 000 instructions instead of ~35 000, and few of them multiplications.
//
// Variable time (loop counts and branches depend on x): the library's scalar multiplication is not constant time either
// (include/plume_b200.h), and the values inverted are denominators of public points.
// Anything unexpected (more than 27 rounds, f != +-1: x = 0 mod p) falls back to fe_inv.
#pragma once
#include "ec.cuh"

struct ds_mat { int32_t u, v, q, r; };   // (f, g) <- (u f + v g, q f + r g) / 2^30

PLUME_DEV int ds_ctz(uint32_t x) {
#ifdef PLUME_HOSTSIM
    return __builtin_ctz(x);
#else
    return __ffs((int)x) - 1;
#endif
}

// 30 division steps on the low words of f (odd) and g; eta = -delta of the paper
PLUME_DEV int ds_divsteps30(int eta, uint32_t f, uint32_t g, ds_mat& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;   // two's complement; |u| + |v|, |q| + |r| <= 2^30
    int i = 30;
    for (;;) {
        const int zeros = ds_ctz(g | (0xFFFFFFFFu << i));   // halvings, at most the i steps that are left
        g >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
        if (i == 0) break;
        if (eta < 0) {   // delta > 0 and g odd: (f, g) <- (g, -f)
            eta = -eta;
            uint32_t w = f; f = g; g = 0u - w;
            w = u; u = q; q = 0u - w;
            w = v; v = r; r = 0u - w;
        }
        g += f; q += u; r += v;   // g odd, f odd: the sum is even and the next pass halves it
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return eta;
}

// out = (a x + b y) >> 30 (arithmetic) for 9-word two's complement x, y and signed a, b with |a| + |b| <= 2^30;
// the low 30 bits of the sum are zero by construction
PLUME_DEV void ds_lincomb_shr30(uint32_t* out, const uint32_t* x, const uint32_t* y, int32_t a, int32_t b) {
    uint32_t T[9];
    const uint32_t ua = (uint32_t)a, ub = (uint32_t)b;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) { c += (uint64_t)x[i] * ua; T[i] = (uint32_t)c; c >>= 32; }
    c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) { c += (uint64_t)y[i] * ub + T[i]; T[i] = (uint32_t)c; c >>= 32; }
    // a negative multiplier was taken as a + 2^32: subtract x << 32 (mod 2^288)
    if (a < 0) {
        uint64_t br = 0;
#pragma unroll
        for (int i = 1; i < 9; i++) { uint64_t d = (uint64_t)T[i] - x[i - 1] - br; T[i] = (uint32_t)d; br = (d >> 32) & 1; }
    }
    if (b < 0) {
        uint64_t br = 0;
#pragma unroll
        for (int i = 1; i < 9; i++) { uint64_t d = (uint64_t)T[i] - y[i - 1] - br; T[i] = (uint32_t)d; br = (d >> 32) & 1; }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (T[i] >> 30) | (T[i + 1] << 2);
    out[8] = (uint32_t)((int32_t)T[8] >> 30);
}

// a * x mod p for a signed 31-bit a
PLUME_DEV fe ds_mul_signed(const fe& x, int32_t a) {
    fe r = fe_mul_small(x, (uint32_t)(a < 0 ? -a : a));
    return a < 0 ? fe_neg(r) : r;
}

PLUME_DEV fe ds_scale(int rounds) {   // 2^(-30 rounds) mod p
    switch (rounds) {
        case 1: return fe_lit(0x4894D4C3u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xB76B2A27u);
        case 2: return fe_lit(0x838091DDu, 0x2253530Fu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7C7F6C2Du);
        case 3: return fe_lit(0x09281676u, 0x0E024774u, 0x894D4C3Fu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xF6D7E967u);
        case 4: return fe_lit(0xB223FEDCu, 0x24A059D8u, 0x38091DD2u, 0x253530FFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x4DDBFE7Cu);
        case 5: return fe_lit(0xCADD86F2u, 0xC88FFB70u, 0x92816760u, 0xE0247748u, 0x94D4C3FFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x35227607u);
        case 6: return fe_lit(0x6C2C295Fu, 0x2B761BCBu, 0x223FEDC2u, 0x4A059D83u, 0x8091DD22u, 0x53530FFFu, 0xFFFFFFFFu, 0x93D3D504u);
        case 7: return fe_lit(0xE0E66711u, 0xB0B0A57Cu, 0xADD86F2Cu, 0x88FFB709u, 0x2816760Eu, 0x02477489u, 0x4D4C3FFFu, 0x1F199594u);
        case 8: return fe_lit(0x19051553u, 0x83999C46u, 0xC2C295F2u, 0xB761BCB2u, 0x23FEDC24u, 0xA059D838u, 0x091DD225u, 0x1C2BEA4Du);
        case 9: return fe_lit(0x0C2B26F4u, 0x6414554Eu, 0x0E66711Bu, 0x0B0A57CAu, 0xDD86F2C8u, 0x8FFB7092u, 0x816760E0u, 0x184C2172u);
        case 10: return fe_lit(0x0BE40348u, 0x30AC9BD1u, 0x90515538u, 0x3999C46Cu, 0x2C295F2Bu, 0x761BCB22u, 0x3FEDC249u, 0xF9B9800Bu);
        case 11: return fe_lit(0xEE6B246Cu, 0x2F900D20u, 0xC2B26F46u, 0x414554E0u, 0xE66711B0u, 0xB0A57CADu, 0xD86F2C88u, 0x114BE12Eu);
        case 12: return fe_lit(0x332A7F3Bu, 0xB9AC91B0u, 0xBE403483u, 0x0AC9BD19u, 0x05155383u, 0x999C46C2u, 0xC295F2B7u, 0x2E923225u);
        case 13: return fe_lit(0x1ED90854u, 0xCCA9FCEEu, 0xE6B246C2u, 0xF900D20Cu, 0x2B26F464u, 0x14554E0Eu, 0x66711B0Au, 0xEB7EC213u);
        case 14: return fe_lit(0xD0C0528Cu, 0x7B642153u, 0x32A7F3BBu, 0x9AC91B0Bu, 0xE4034830u, 0xAC9BD190u, 0x51553838u, 0xC9041683u);
        case 15: return fe_lit(0x2581B84Fu, 0x43014A31u, 0xED90854Cu, 0xCA9FCEEEu, 0x6B246C2Fu, 0x900D20C2u, 0xB26F4641u, 0x1FD32808u);
        case 16: return fe_lit(0x9F714620u, 0x9606E13Du, 0x0C0528C7u, 0xB6421533u, 0x2A7F3BB9u, 0xAC91B0BEu, 0x4034830Au, 0x2A4BD084u);
        case 17: return fe_lit(0xD708F512u, 0x7DC51882u, 0x581B84F4u, 0x3014A31Eu, 0xD90854CCu, 0xA9FCEEE6u, 0xB246C2F8u, 0x29C913E4u);
        case 18: return fe_lit(0x223C0A93u, 0x5C23D449u, 0xF7146209u, 0x606E13D0u, 0xC0528C7Bu, 0x64215332u, 0xA7F3BB9Au, 0xA6DF00CEu);
        case 19: return fe_lit(0xD67F35B8u, 0x88F02A4Du, 0x708F5127u, 0xDC518825u, 0x81B84F43u, 0x014A31EDu, 0x90854CC9u, 0xC94FB580u);
        case 20: return fe_lit(0x1154F603u, 0x59FCD6E2u, 0x23C0A935u, 0xC23D449Fu, 0x71462096u, 0x06E13D0Cu, 0x0528C7B6u, 0x30C03CE5u);
        case 21: return fe_lit(0x22004354u, 0x4553D80Du, 0x67F35B88u, 0x8F02A4D7u, 0x08F5127Du, 0xC5188258u, 0x1B84F42Fu, 0xF2A2DB03u);
        case 22: return fe_lit(0x17CA2A4Cu, 0x88010D51u, 0x154F6035u, 0x9FCD6E22u, 0x3C0A935Cu, 0x23D449F7u, 0x14620960u, 0x5649A619u);
        case 23: return fe_lit(0x1C63DF24u, 0x5F28A932u, 0x20043544u, 0x553D80D6u, 0x7F35B888u, 0xF02A4D70u, 0x8F5127DCu, 0x352445F1u);
        case 24: return fe_lit(0xF1052084u, 0x718F7C91u, 0x7CA2A4C8u, 0x8010D511u, 0x54F60359u, 0xFCD6E223u, 0xC0A935C1u, 0x4C3F7B55u);
        case 25: return fe_lit(0x27E0D117u, 0xC4148211u, 0xC63DF245u, 0xF28A9322u, 0x00435445u, 0x53D80D67u, 0xF35B888Eu, 0xDAC40559u);
        case 26: return fe_lit(0xF1F5CC24u, 0x9F83445Fu, 0x10520847u, 0x18F7C917u, 0xCA2A4C88u, 0x010D5115u, 0x4F60359Eu, 0xDB78527Cu);
        case 27: return fe_lit(0x3C1DD6F3u, 0xC7D73092u, 0x7E0D117Cu, 0x4148211Cu, 0x63DF245Fu, 0x28A93220u, 0x04354455u, 0x0162FEA6u);
        default: return fe_zero();
    }
}

// 1 / x mod p (0 for x = 0 mod p); x any representative below 2^256
PLUME_DEV fe fe_inv_var(const fe& x) {
    uint32_t f[9] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u};
    uint32_t g[9];
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] = x.v[i];
    g[8] = 0;
    fe d = fe_zero(), e = fe_one();
    int eta = -1, rounds = 0;
#pragma unroll 1
    for (;;) {
        ds_mat t;
        eta = ds_divsteps30(eta, f[0], g[0], t);
        fe nd = fe_add(ds_mul_signed(d, t.u), ds_mul_signed(e, t.v));
        e = fe_add(ds_mul_signed(d, t.q), ds_mul_signed(e, t.r));
        d = nd;
        uint32_t nf[9], ng[9];
        ds_lincomb_shr30(nf, f, g, t.u, t.v);
        ds_lincomb_shr30(ng, f, g, t.q, t.r);
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) { f[i] = nf[i]; g[i] = ng[i]; any |= ng[i]; }
        rounds++;
        if (any == 0 || rounds > 27) break;
    }
    // f = +1 or -1 when x was invertible
    bool plus = f[0] == 1u, minus = f[0] == 0xFFFFFFFFu;
#pragma unroll
    for (int i = 1; i < 9; i++) { plus = plus && f[i] == 0u; minus = minus && f[i] == 0xFFFFFFFFu; }
    if (rounds > 27 || !(plus || minus)) return fe_inv(x);
    fe r = fe_mul(d, ds_scale(rounds));
    return minus ? fe_neg(r) : r;
}
