// hostsim.cpp -- TEST-ONLY build of the kernel sources for a machine without a GPU.
//
// Compiles zk-nullifier-sig_b200/csrc/*.cuh with g++ and -DPLUME_HOSTSIM (the PTX wrappers in
// ptx.cuh become carry-flag emulations) and runs the stage bodies in plain loops, so the limb-level
// arithmetic, the GLV/Booth recoding, the SSWU map, the stage plumbing and all edge-case
// branches of the device code can be checked against the oracle by `pytest -m "not gpu"`.
// It is NOT a product path: libplume_b200.so never links or calls it, and has no CPU fallback.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "stages.cuh"

static std::vector<uint32_t> g_tab;
static int g_w = 0;

static void build_gtab(int w) {
    if (g_w == w) return;
    const int nwin = (256 + w - 1) / w;
    const size_t ne = (size_t)nwin << w;
    std::vector<uint32_t> bases(nwin * 16), zs(ne * 8), scratch(ne * 8);
    g_tab.assign(ne * 16, 0);
    gtab_bases_body(bases.data(), w);
    for (uint32_t e = 0; e < ne; e++) gtab_entry_body(e, g_tab.data(), zs.data(), bases.data(), w);
    const uint32_t T = 64;
    for (uint32_t t = 0; t < T; t++) binv_body(t, T, zs.data(), scratch.data(), (uint32_t)ne);
    for (uint32_t e = 0; e < ne; e++) gtab_norm_body(e, g_tab.data(), zs.data());
    g_w = w;
}

static void run_binv(uint32_t* ws, uint32_t n, uint32_t m, uint32_t T) {
    for (uint32_t t = 0; t < T; t++) binv_body(t, T, ws_at(ws, n, WS_Z0, 0), ws_at(ws, n, WS_P0, 0), m);
}

extern "C" {

void hs_fe_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    fe x, y, r;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    switch (op) {
        case 0: r = fe_mul(x, y); break;
        case 1: r = fe_sqr(x); break;
        case 2: r = fe_add(x, y); break;
        case 3: r = fe_sub(x, y); break;
        case 4: r = fe_inv(x); break;
        case 14: r = fe_inv_var(x); break;
        case 5: r = fe_norm(x); break;
        case 6: r = fe_mul_small(x, b[0]); break;
        case 7: r = fe_pow_pm3d4(x); break;
        case 8: r = fe_neg(x); break;
        case 9: r = fe_sqrt_cand(x); break;
        case 11: r = fe_shl<3>(x); break;
        case 12: r = fe_shl<1>(x); break;
        case 13: r = fe_shl<2>(x); break;
        default: r = fe_zero();
    }
    memcpy(out, r.v, 32);
}
int hs_fe_is_zero(const uint32_t* a) { fe x; memcpy(x.v, a, 32); return fe_is_zero(x) ? 1 : 0; }

void hs_sc_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    sc x, y, r;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    switch (op) {
        case 0: r = sc_mul(x, y); break;
        case 1: r = sc_add(x, y); break;
        case 2: r = sc_neg(x); break;
        case 3: r = sc_reduce256(x); break;
        default: memset(r.v, 0, 32);
    }
    memcpy(out, r.v, 32);
}
void hs_sc_reduce512(const uint32_t* x16, uint32_t* out) { sc r = sc_reduce512(x16); memcpy(out, r.v, 32); }

// out: mag1[5], neg1, mag2[5], neg2, then 33 digits of half 1 and 33 of half 2 (as int32)
void hs_glv(const uint32_t* k, uint32_t* out) {
    sc x; memcpy(x.v, k, 32);
    glv_half h1, h2;
    glv_split(x, h1, h2);
    memcpy(out, h1.mag, 20); out[5] = h1.neg;
    memcpy(out + 6, h2.mag, 20); out[11] = h2.neg;
    booth_reg b1 = booth_init(h1), b2 = booth_init(h2);
    for (int i = 0; i < 33; i++) out[12 + i] = (uint32_t)booth_next(b1);
    for (int i = 0; i < 33; i++) out[45 + i] = (uint32_t)booth_next(b2);
}

void hs_sha256(const uint8_t* p, uint32_t n, uint8_t* out) {
    sha256_stream s; sha256_init(s.st); s.fill = 0; s.total = 0;
    sha256_stream_bytes(s, p, n);
    uint32_t d[8]; sha256_stream_final(s, d);
    for (int i = 0; i < 8; i++) { out[4*i] = d[i] >> 24; out[4*i+1] = d[i] >> 16; out[4*i+2] = d[i] >> 8; out[4*i+3] = d[i]; }
}

// k * P (affine BE 64 bytes) -> affine BE 64 bytes, through the variable-base path
void hs_vb_mul(const uint8_t* p64, const uint8_t* k32, uint8_t* out64) {
    aff p; ld_point_be(p, p64);
    sc k = ld_sc_be(k32);
    uint32_t tabw[VB_TAB_WORDS];
    jac r = vb_mul_point(p, k, tabw);
    aff q = r.inf ? aff_infinity() : aff_from_jac_zinv(r, fe_inv(r.z));
    st_point_be(out64, q);
}
void hs_comb_mul(const uint8_t* p64, const uint8_t* k32, uint8_t* out64) {
    aff p; ld_point_be(p, p64);
    sc k = ld_sc_be(k32);
    uint32_t area[COMB_AREA_WORDS];
    fe zg = comb_build_table(p.x, p.y, area);
    fe zg2 = fe_sqr(zg);
    jac r = comb_mul_tab(k, area, zg, fe_mul(p.x, zg2), fe_mul(p.y, fe_mul(zg2, zg)));
    aff q = r.inf ? aff_infinity() : aff_from_jac_zinv(r, fe_inv(r.z));
    st_point_be(out64, q);
}
void hs_fb_mul(const uint8_t* k32, int w, uint8_t* out64) {
    build_gtab(w);
    sc k = ld_sc_be(k32);
    jac r = fb_mul(k, g_tab.data(), w);
    aff q = r.inf ? aff_infinity() : aff_from_jac_zinv(r, fe_inv(r.z));
    st_point_be(out64, q);
}

// The per-lane shares of the small-batch ("team") kernels, added up serially here: the two GLV halves of the windowed
// ladder / of the signed comb, and the lower and upper windows of the generator walk.
static void hs_store(const jac& r, uint8_t* out64) {
    aff q = r.inf ? aff_infinity() : aff_from_jac_zinv(r, fe_inv(r.z));
    st_point_be(out64, q);
}
void hs_vb_mul_halves(const uint8_t* p64, const uint8_t* k32, uint8_t* out64) {
    aff p; ld_point_be(p, p64);
    sc k = ld_sc_be(k32);
    uint32_t tabw[VB_TAB_WORDS];
    fe zg = vb_build_table(p.x, p.y, tabw, true);
    glv_half h1, h2;
    glv_split(k, h1, h2);
    jac r = jac_add(vb_ladder_half(h1, false, tabw), vb_ladder_half(h2, true, tabw));
    if (!r.inf) r.z = fe_mul(r.z, zg);
    hs_store(r, out64);
}
void hs_comb_mul_halves(const uint8_t* p64, const uint8_t* k32, uint8_t* out64) {
    aff p; ld_point_be(p, p64);
    sc k = ld_sc_be(k32);
    uint32_t area[COMB_AREA_WORDS];
    fe zg = comb_build_table(p.x, p.y, area);
    fe zg2 = fe_sqr(zg);
    fe pxs = fe_mul(p.x, zg2), pys = fe_mul(p.y, fe_mul(zg2, zg));
    glv_half h1, h2;
    glv_split(k, h1, h2);
    jac r = jac_add(comb_ladder_half(h1, false, area, pxs, pys), comb_ladder_half(h2, true, area, pxs, pys));
    if (!r.inf) r.z = fe_mul(r.z, zg);
    hs_store(r, out64);
}
void hs_fb_mul_split(const uint8_t* k32, int w, uint8_t* out64) {
    build_gtab(w);
    sc k = ld_sc_be(k32);
    const int nw = fb_windows(w), mid = nw / 2;
    hs_store(jac_add(fb_mul_windows(k, g_tab.data(), w, 0, mid), fb_mul_windows(k, g_tab.data(), w, mid, nw)), out64);
}

// flavour 0: k256 (pk is an output), 1: arkworks (pk is an input)
// comb != 0: the shipped signed-comb form of the variable-base stage; 0: the windowed ladder (both table layouts)
int hs_sign_batch(int flavour, int comb, int version, uint32_t n, const uint8_t* msgs, const uint64_t* offs, uint32_t msg_len,
                  const uint8_t* sk, const uint8_t* r, uint8_t* pk, uint8_t* nul, uint8_t* c, uint8_t* s,
                  uint8_t* r_point, uint8_t* hr, uint8_t* status, int gw, uint32_t binv_threads) {
    build_gtab(gw);
    std::vector<uint32_t> ws((size_t)WS_SLOTS * n * 8);
    sign_args a{};
    a.version = version; a.flavour = flavour; a.n = n; a.msgs.base = msgs; a.msgs.offs = offs; a.msgs.fixed_len = msg_len;
    a.sk = sk; a.r = r; a.pk = flavour ? nullptr : pk; a.pk_in = flavour ? pk : nullptr;
    a.nullifier = nul; a.c = c; a.s = s; a.r_point = r_point; a.hashed_to_curve_r = hr;
    a.status = status; a.ws = ws.data(); a.gtab = g_tab.data(); a.gw = gw; a.vbtab = nullptr;
    std::vector<uint32_t> tabv(VB_ITEM_WORDS);
    uint32_t* tabw = tabv.data();
    for (uint32_t i = 0; i < n; i++) sign_stage_fixed(i, a);
    run_binv(a.ws, n, 2 * n, binv_threads);
    for (uint32_t i = 0; i < n; i++) sign_stage_h2c(i, a);
    run_binv(a.ws, n, n, binv_threads);
    if (comb == 2) {   // the shipped split: table kernel, then one ladder thread per item and scalar
        std::vector<uint32_t> tabs((size_t)n * VB_ITEM_WORDS);
        a.vbtab = tabs.data();
        for (uint32_t i = 0; i < n; i++) sign_stage_varbase_tab(i, a, tabs.data() + (size_t)i * VB_ITEM_WORDS);
        for (uint32_t idx = 2 * n; idx-- > 0;) sign_stage_varbase_lad(idx, a, tabs.data());
        a.vbtab = nullptr;
    } else {
        for (uint32_t i = 0; i < n; i++) {
            if (comb) sign_stage_varbase_comb(i, a, tabw);
            else sign_stage_varbase(i, a, tabw);
        }
    }
    run_binv(a.ws, n, 2 * n, binv_threads);
    for (uint32_t i = 0; i < n; i++) sign_stage_final(i, a);
    return 0;
}

// the shipped form: table kernel + ladder kernel + G*s - pk*c kernel (`fused` is ignored: the one-kernel forms of
// round 1 are gone)
int hs_verify_batch(int flavour, int version, uint32_t n, const uint8_t* msgs, const uint64_t* offs, uint32_t msg_len,
                    const uint8_t* pk, const uint8_t* nul, const uint8_t* c, const uint8_t* s,
                    const uint8_t* r_point, const uint8_t* hr, uint8_t* ok, int gw, uint32_t binv_threads, int fused) {
    build_gtab(gw);
    std::vector<uint32_t> ws((size_t)WS_SLOTS * n * 8);
    verify_args a{};
    a.version = version; a.flavour = flavour; a.n = n; a.msgs.base = msgs; a.msgs.offs = offs; a.msgs.fixed_len = msg_len;
    a.pk = pk; a.nullifier = nul; a.c = c; a.s = s; a.r_point = r_point; a.hashed_to_curve_r = hr;
    a.ok = ok; a.ws = ws.data(); a.gtab = g_tab.data(); a.gw = gw; a.vbtab = nullptr;
    (void)fused;
    for (uint32_t i = 0; i < n; i++) verify_stage_h2c(i, a);
    for (uint32_t t = 0; t < binv_threads; t++) binv_body(t, binv_threads, ws_at(a.ws, n, WS_Z1, 0), ws_at(a.ws, n, WS_P0, 0), n);
    {
        // per-item table storage: the tables live from the table kernel to the ladder kernel
        std::vector<uint32_t> tabs((size_t)n * VB_ITEM_WORDS);
        auto t1 = [&](uint32_t i) { return tabs.data() + (size_t)i * VB_ITEM_WORDS; };
        auto t2 = [&](uint32_t i) { return tabs.data() + (size_t)i * VB_ITEM_WORDS + VB_TAB_WORDS; };
        // G*s - pk*c first here (the library runs it last, or concurrently for small batches): the stages are independent
        for (uint32_t i = 0; i < n; i++) verify_stage_mul_a(i, a, t2(i) + VB_TAB_WORDS);
        for (uint32_t i = 0; i < n; i++) verify_stage_mul_b1(i, a, t1(i), t2(i));
        for (uint32_t i = 0; i < n; i++) verify_stage_mul_b2(i, a, t1(i), t2(i));
    }
    run_binv(a.ws, n, 2 * n, binv_threads);
    for (uint32_t i = 0; i < n; i++) verify_stage_final(i, a);
    return 0;
}

int hs_h2c_batch(uint32_t n, const uint8_t* msgs, const uint64_t* offs, uint32_t msg_len, uint8_t* out, uint32_t binv_threads) {
    std::vector<uint32_t> ws((size_t)WS_SLOTS * n * 8);
    h2c_args a{};
    a.n = n; a.msgs.base = msgs; a.msgs.offs = offs; a.msgs.fixed_len = msg_len; a.out = out; a.ws = ws.data();
    for (uint32_t i = 0; i < n; i++) h2c_stage_map(i, a);
    run_binv(a.ws, n, n, binv_threads);
    for (uint32_t i = 0; i < n; i++) h2c_stage_out(i, a);
    return 0;
}

int hs_h2c_witness_batch(uint32_t n, const uint8_t* msgs, const uint64_t* offs, uint32_t msg_len, uint8_t* u, uint8_t* q,
                         uint8_t* gx1_square, uint8_t* h, uint32_t binv_threads, uint8_t* hints) {
    std::vector<uint32_t> ws((size_t)WS_SLOTS * n * 8);
    h2cw_args a{};
    a.n = n; a.msgs.base = msgs; a.msgs.offs = offs; a.msgs.fixed_len = msg_len; a.u = u; a.q = q; a.gx1_square = gx1_square; a.h = h; a.hints = hints;
    a.ws = ws.data();
    for (uint32_t i = 0; i < n; i++) h2cw_stage_map(i, a);
    run_binv(a.ws, n, 2 * n, binv_threads);
    for (uint32_t i = 0; i < n; i++) h2cw_stage_sum(i, a);
    run_binv(a.ws, n, n, binv_threads);
    for (uint32_t i = 0; i < n; i++) h2cw_stage_out(i, a);
    return 0;
}
void hs_registers(uint32_t n, const uint8_t* in32, uint64_t* out4) {
    for (uint32_t i = 0; i < n; i++) registers_body(i, in32, out4);
}

void hs_sec1_roundtrip(uint32_t n, const uint8_t* in33, uint8_t* out64, uint8_t* ok, uint8_t* back33) {
    for (uint32_t i = 0; i < n; i++) sec1_decompress_body(i, in33, out64, ok);
    for (uint32_t i = 0; i < n; i++) sec1_compress_body(i, out64, back33);
}

// SSWU + isogeny of one field element (canonical LE limbs) -> affine BE 64 bytes
void hs_map_to_curve(const uint32_t* u, uint8_t* out64) {
    fe x; memcpy(x.v, u, 32);
    fe xn, xd, y;
    h2c_map_sswu(xn, xd, y, x);
    jac q = h2c_iso_map(xn, xd, y);
    aff a = q.inf ? aff_infinity() : aff_from_jac_zinv(q, fe_inv(q.z));
    st_point_be(out64, a);
}

}  // extern "C"
