// emu.cpp — TEST-ONLY host build of the kernel bodies (rustcrypto-elliptic-curves_b200/csrc/*.cuh with
// ECB_EMU: PTX carry chains replaced by a carry-flag emulation).  Lets the CPU-only test tier check
// the limb-level algorithms (field reduction, formulas, GLV, windows, verify epilogue) in a container
// without a GPU.  It is NOT part of the product: the package never loads this library, and all
// parity claims are made on the B200 through libecb200.so.
#define ECB_EMU 1
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../rustcrypto-elliptic-curves_b200/csrc/kernels.cuh"

using namespace ecb;

template <class C> struct Emu {
    typedef Bodies<C> B;
    static constexpr int L = C::L;
    static constexpr int FB = C::FB;

    static void normalize(int n, const u32* proj, int mode, int compress, u8* out, u8* inf, u32* limbs, int nthreads) {
        int need = (n + B::EPT - 1) / B::EPT;
        if (nthreads < need) nthreads = need;
        if (nthreads < 1) nthreads = 1;
        for (int t = 0; t < nthreads; t++) B::body_normalize(t, nthreads, n, proj, mode, compress, out, inf, limbs);
    }
    static int ngtab() { return C::A_IS_ZERO ? 8 : 15; }
    // affine multiples 1..NG of G, via the same kernels the device init uses
    static std::vector<u32> gtab() {
        int ng = ngtab();
        std::vector<u8> pts(2 * FB * ng), ks(FB * ng, 0);
        typename EC<C>::Aff g;
        EC<C>::generator(g);
        u32 t[L];
        for (int j = 0; j < ng; j++) {
            C::F::to_limbs(t, g.x); store_be<L>(&pts[2 * FB * j], t);
            C::F::to_limbs(t, g.y); store_be<L>(&pts[2 * FB * j + FB], t);
            ks[FB * j + FB - 1] = (u8)(j + 1);
        }
        std::vector<u32> proj(3 * L * ng), out(2 * L * ng);
        for (int i = 0; i < ng; i++) B::template body_mul_var<false>(i, ng, 0, pts.data(), nullptr, ks.data(), proj.data(), nullptr);
        normalize(ng, proj.data(), NORM_AFF_LIMBS, 0, nullptr, nullptr, out.data(), 0);
        return out;
    }
    // fixed-base table: (j+1) * 16^i * G, i < 8L + 1, j < 8
    static std::vector<u32> gentab() {
        const int NW = B::GEN_WINDOWS, ne = NW * 8;
        std::vector<u8> pts(2 * FB * ne), ks(FB * ne, 0);
        typename EC<C>::Aff g;
        EC<C>::generator(g);
        u32 t[L];
        for (int e = 0; e < ne; e++) {
            C::F::to_limbs(t, g.x); store_be<L>(&pts[2 * FB * e], t);
            C::F::to_limbs(t, g.y); store_be<L>(&pts[2 * FB * e + FB], t);
            int i = e / 8, j = e % 8;
            int bit = 4 * i;
            unsigned v = (unsigned)(j + 1);
            if (bit < 8 * FB) {
                int byte = bit / 8, sh = bit % 8;
                unsigned w = v << sh;
                ks[FB * e + FB - 1 - byte] |= (u8)w;
                if (byte + 1 < FB) ks[FB * e + FB - 2 - byte] |= (u8)(w >> 8);
            } else {
                // (j+1) * (2^(8FB) mod n), j+1 <= 8: computed with the scalar field
                typename C::Fn::E a, acc;
                for (int l = 0; l < L; l++) a.v[l] = C::Fn::Params::one(l);   // R mod n (plain value)
                acc = a;
                for (unsigned m = 1; m < v; m++) C::Fn::add(acc, acc, a);
                store_be<L>(&ks[FB * e], acc.v);
            }
        }
        std::vector<u32> proj(3 * L * ne), out(2 * L * ne);
        for (int i = 0; i < ne; i++) B::template body_mul_var<false>(i, ne, 0, pts.data(), nullptr, ks.data(), proj.data(), nullptr);
        normalize(ne, proj.data(), NORM_AFF_LIMBS, 0, nullptr, nullptr, out.data(), 0);
        return out;
    }

    // big fixed-base table: nwin = ceil(8FB / gw) windows; signed gw-bit windows below the top one (entry (w << (gw-1)) + v - 1 =
    // v * 2^(gw*w) * G, v <= 2^(gw-1)), unsigned top window of tb = 8FB - gw (nwin - 1) bits with v up to 2^tb (the last entry is
    // 2^(8FB) * G) - the layout abi.cu ensure_gbig builds.  ECB_EMU_GW picks the width (default 5: windows that straddle words
    // and a short top window; 4 = the word-aligned case)
    static std::vector<u32> gbig(int gw) {
        const int nwin = (32 * L + gw - 1) / gw, tb = 32 * L - gw * (nwin - 1), per = 1 << (gw - 1), top_base = (nwin - 1) * per, ne = top_base + (1 << tb);
        std::vector<u8> pts(2 * FB * (size_t)ne), ks(FB * (size_t)ne, 0);
        typename EC<C>::Aff g;
        EC<C>::generator(g);
        u32 t[L];
        for (int e = 0; e < ne; e++) {
            C::F::to_limbs(t, g.x); store_be<L>(&pts[2 * FB * (size_t)e], t);
            C::F::to_limbs(t, g.y); store_be<L>(&pts[2 * FB * (size_t)e + FB], t);
            if (e == ne - 1) {
                u32 one[L];
                for (int l = 0; l < L; l++) one[l] = C::Fn::Params::one(l);   // R mod n = 2^(8FB) mod n
                store_be<L>(&ks[FB * (size_t)e], one);
                continue;
            }
            int w = e < top_base ? e / per : nwin - 1, v = e - w * per + 1, bit = w * gw;
            for (int b = 0; b < gw + 1; b++) if ((v >> b) & 1) { int pos = bit + b; if (pos < 8 * FB) ks[FB * (size_t)e + FB - 1 - pos / 8] |= (u8)(1u << (pos % 8)); }
        }
        std::vector<u32> proj(3 * L * (size_t)ne), out(2 * L * (size_t)ne);
        for (int i = 0; i < ne; i++) B::body_mul_var_fast(i, ne, pts.data(), nullptr, nullptr, ks.data(), proj.data(), nullptr);
        normalize(ne, proj.data(), NORM_AFF_LIMBS, 0, nullptr, nullptr, out.data(), 0);
        return out;
    }
    static int emu_gw() {
        const char* e = getenv("ECB_EMU_GW");
        const int g = e ? atoi(e) : 5;
        return g >= 2 && g <= 8 ? g : 5;
    }
    static const std::vector<u32>& gbig_cached(int gw) {
        static std::vector<u32> gt[9];
        if (gt[gw].empty()) gt[gw] = gbig(gw);
        return gt[gw];
    }
    // the two-term pipeline in every mode (VM_*): prep -> main [-> normalise -> finish]
    static void verify_mode(int mode, int n, const u8* q, const u8* z, const u8* rs, const u8* aux, u8* ok, u8* out, int compress, int nthreads) {
        const int gw = emu_gw();
        const std::vector<u32>& gt = gbig_cached(gw);
        std::vector<u32> scratch((size_t)n * B::PREP_WORDS), proj((size_t)3 * L * n);
        int need = (n + B::PREP_EPT - 1) / B::PREP_EPT;
        if (nthreads < need) nthreads = need;
        for (int t = 0; t < nthreads; t++) B::body_verify_prep(t, nthreads, n, mode, z, rs, scratch.data());
        // primeorder curves, ECDSA / SM2DSA: affine window tables for the whole batch first (as abi.cu window_tables does)
        std::vector<u32> wt;
        if (!C::A_IS_ZERO && (mode == VM_ECDSA || mode == VM_SM2DSA)) wt = wintab(n, q, nullptr, nthreads);
        for (int i = 0; i < n; i++)
            B::body_verify_main(i, n, mode, q, rs, z, aux, scratch.data(), gt.data(), gw, ok, proj.data(), nullptr, 0, wt.empty() ? nullptr : wt.data());
        if (mode == VM_SCHNORR) {
            std::vector<u8> xy((size_t)2 * FB * n), inf(n);
            normalize(n, proj.data(), NORM_XY_BYTES, 0, xy.data(), inf.data(), nullptr, 0);
            for (int i = 0; i < n; i++) B::body_finish(i, n, FIN_SCHNORR, xy.data(), 0, inf.data(), rs, ok);
        } else if (mode == VM_RECOVER) {
            normalize(n, proj.data(), NORM_SEC1, compress, out, nullptr, nullptr, 0);
            const int stride = compress ? 1 + FB : 1 + 2 * FB;
            for (int i = 0; i < n; i++) B::body_finish(i, n, FIN_RECOVER, out, stride, nullptr, nullptr, ok);
        }
    }
    static std::vector<u32> wintab(int n, const u8* pts, const u32* aff, int nthreads) {
        std::vector<u32> wt((size_t)n * 23 * L);
        int need = (n + B::WT_EPT - 1) / B::WT_EPT;
        if (nthreads < need) nthreads = need;
        for (int t = 0; t < nthreads; t++) B::body_wintab(t, nthreads, n, pts, aff, wt.data(), wt.data() + (size_t)n * 16 * L);
        return wt;
    }
    // the per-key-table verify path as abi.cu drives it: rows arrive in chunks, keys are grouped incrementally (lookup, insert,
    // number, assign), tables are built for the keys first seen in a chunk, rows are verified on the tables.  Returns the number
    // of distinct keys found.  cap = table capacity in groups (rows of overflowing groups would take the per-row path: -1 here)
    static int verify_keytab(int mode, int n, const u8* q, const u8* z, const u8* rs, u8* ok, int chunk, int cap, int nthreads) {
        const int gw = emu_gw();
        const std::vector<u32>& gt = gbig_cached(gw);
        const int KBW = B::KBW, W = B::KT_WINDOWS;
        size_t hs = 16;
        while (hs < (size_t)2 * (cap + chunk)) hs <<= 1;
        std::vector<int> htab(hs, B::KT_EMPTY), counter(4, 0);
        std::vector<u32> gkeys((size_t)cap * KBW), tab((size_t)cap * B::KT_KEY_WORDS);
        std::vector<u8> kvalid(cap);
        int built = 0;
        for (int off = 0; off < n; off += chunk) {
            const int cnt = n - off < chunk ? n - off : chunk;
            std::vector<u32> q32((size_t)cnt * KBW);
            memcpy(q32.data(), q + (size_t)off * 2 * FB, (size_t)cnt * 2 * FB);
            std::vector<int> gid(cnt, -7), rep(cnt, -7), rep_slot(cnt, -7), newgid(cnt, -7);
            for (int t = 0; t < cnt; t++) B::body_kt_lookup(t, cnt, q32.data(), htab.data(), (u32)(hs - 1), gkeys.data(), gid.data());
            for (int t = cnt - 1; t >= 0; t--) B::body_kt_insert(t, cnt, q32.data(), htab.data(), (u32)(hs - 1), gid.data(), rep.data(), rep_slot.data());   // any order
            for (int t = 0; t < cnt; t++) B::body_kt_number(t, cnt, q32.data(), htab.data(), gid.data(), rep.data(), rep_slot.data(), counter.data(), cap, gkeys.data(), newgid.data());
            for (int t = 0; t < cnt; t++) B::body_kt_assign(t, cnt, gid.data(), rep.data(), newgid.data());
            const int D = counter[0];
            if (D > cap) return -D;
            if (D > built) {
                const int c2 = D - built;
                std::vector<u32> proj((size_t)c2 * W * 3 * L);
                for (int t = 0; t < c2; t++) B::body_kt_base(t, built, c2, gkeys.data(), proj.data(), kvalid.data());
                u32* t0 = tab.data() + (size_t)built * B::KT_KEY_WORDS;
                normalize(c2 * W, proj.data(), NORM_AFF_STRIDED, B::KT_E * 2 * L, nullptr, nullptr, t0, 0);
                const int items = c2 * W;
                int th = (items + B::KT_EPT - 1) / B::KT_EPT;
                if (nthreads > th) th = nthreads;
                for (int t = 0; t < th; t++) B::template body_kt_fill<OwnInv>(t, th, items, t0);
                built = D;
            }
            std::vector<u32> scratch((size_t)cnt * B::PREP_WORDS);
            int need = (cnt + B::PREP_EPT - 1) / B::PREP_EPT;
            for (int t = 0; t < need; t++) B::body_verify_prep(t, need, cnt, mode, z + (size_t)off * FB, rs + (size_t)off * 2 * FB, scratch.data());
            for (int t = 0; t < cnt; t++)
                B::body_verify_keytab(t, cnt, mode, rs + (size_t)off * 2 * FB, z + (size_t)off * FB, scratch.data(), gid.data(), kvalid.data(), tab.data(), gt.data(), gw, ok + off);
        }
        return counter[0];
    }
    static void verify2(int n, const u8* q, const u8* z, const u8* rs, u8* ok, int nthreads) {
        verify_mode(VM_ECDSA, n, q, z, rs, nullptr, ok, nullptr, 0, nthreads);
    }
    static void decode(int n, int mode, const u8* enc, int stride, u8* xy, u8* status) {
        for (int i = 0; i < n; i++) B::body_decode(i, n, mode, enc, stride, xy, status);
    }
    static void sign(int n, const u8* d, const u8* k, const u8* z, u8* rs, u8* recid, u8* ok, int nthreads) {
        static std::vector<u32> tab;
        if (tab.empty()) tab = gentab();
        std::vector<u32> proj((size_t)3 * L * n), aff((size_t)2 * L * n);
        for (int i = 0; i < n; i++) B::template body_mul_gen<true>(i, n, k, tab.data(), proj.data());
        normalize(n, proj.data(), NORM_AFF_LIMBS, 0, nullptr, nullptr, aff.data(), 0);
        int need = (n + B::PREP_EPT - 1) / B::PREP_EPT;
        if (nthreads < need) nthreads = need;
        for (int t = 0; t < nthreads; t++) B::body_sign_finish(t, nthreads, n, d, k, z, aff.data(), rs, recid, ok);
    }
    static void mul_var_fast(int n, const u8* pts, const u8* inf, const u8* k, u8* out, int compress, u8* invalid) {
        std::vector<u32> proj((size_t)3 * L * n);
        std::vector<u32> wt;
        if (!C::A_IS_ZERO) wt = wintab(n, pts, nullptr, (n + 2) / 3);   // three rows per logical thread
        for (int i = 0; i < n; i++) B::body_mul_var_fast(i, n, pts, nullptr, inf, k, proj.data(), invalid, nullptr, 0, wt.empty() ? nullptr : wt.data());
        normalize(n, proj.data(), NORM_SEC1, compress, out, nullptr, nullptr, 0);
    }
    static void field_op(int which, int op, int n, const u8* a, const u8* b, u8* out, u8* ok) {
        for (int i = 0; i < n; i++) B::body_field_op(i, n, which, op, a, b, out, ok);
    }
    static void mul_var(int ct, int n, u32 flags, const u8* pts, const u8* inf, const u8* k, u8* out, int compress, u8* invalid, int nthreads) {
        std::vector<u32> proj((size_t)3 * L * n);
        for (int i = 0; i < n; i++) {
            if (ct) B::template body_mul_var<true>(i, n, flags, pts, inf, k, proj.data(), invalid);
            else B::template body_mul_var<false>(i, n, flags, pts, inf, k, proj.data(), invalid);
        }
        normalize(n, proj.data(), NORM_SEC1, compress, out, nullptr, nullptr, nthreads);
    }
    static void mul_gen(int ct, int n, const u8* k, u8* out, int compress) {
        static std::vector<u32> tab;
        if (tab.empty()) tab = gentab();
        std::vector<u32> proj((size_t)3 * L * n);
        for (int i = 0; i < n; i++) {
            if (ct) B::template body_mul_gen<true>(i, n, k, tab.data(), proj.data());
            else B::template body_mul_gen<false>(i, n, k, tab.data(), proj.data());
        }
        normalize(n, proj.data(), NORM_SEC1, compress, out, nullptr, nullptr, 0);
    }
    // split fixed-base table: v * 2^(W w) * G, w < G2_WINDOWS, 1 <= v <= G2_E (as abi.cu build_tables)
    static std::vector<u32> gentab2() {
        const int W = B::G2_W, NW = B::G2_WINDOWS, E = B::G2_E, ne = NW * E, top_bits = 8 * FB - W * (NW - 1);
        std::vector<u8> pts(2 * FB * (size_t)ne), ks(FB * (size_t)ne, 0);
        typename EC<C>::Aff g;
        EC<C>::generator(g);
        u32 t[L];
        for (int e = 0; e < ne; e++) {
            C::F::to_limbs(t, g.x); store_be<L>(&pts[2 * FB * (size_t)e], t);
            C::F::to_limbs(t, g.y); store_be<L>(&pts[2 * FB * (size_t)e + FB], t);
            const int w = e / E, v = e % E + 1;
            u8* s = &ks[FB * (size_t)e];
            if (w == NW - 1 && v == (1 << top_bits)) {
                u32 one[L];
                for (int l = 0; l < L; l++) one[l] = C::Fn::Params::one(l);   // R mod n = 2^(8FB) mod n
                store_be<L>(s, one);
            } else if (w == NW - 1 && v > (1 << top_bits)) {
                s[FB - 1] = 1;
            } else {
                for (int b = 0; b < W; b++) if ((v >> b) & 1) { int pos = W * w + b; s[FB - 1 - pos / 8] |= (u8)(1u << (pos % 8)); }
            }
        }
        std::vector<u32> proj(3 * L * (size_t)ne), out(2 * L * (size_t)ne);
        for (int i = 0; i < ne; i++) B::template body_mul_var<false>(i, ne, 0, pts.data(), nullptr, ks.data(), proj.data(), nullptr);
        normalize(ne, proj.data(), NORM_AFF_LIMBS, 0, nullptr, nullptr, out.data(), 0);
        return out;
    }
    // the split fixed-base path as abi.cu gen_points drives it: two half sums per scalar, complete addition inside the normalisation
    static void mul_gen2(int ct, int n, const u8* k, u8* out, int compress, int nthreads) {
        static std::vector<u32> tab;
        if (tab.empty()) tab = gentab2();
        std::vector<u32> part((size_t)2 * 3 * L * n);
        for (int h = 0; h < 2; h++) {
            const u32* th = tab.data() + (size_t)h * B::G2_NH * B::G2_E * 2 * L;
            for (int i = 0; i < n; i++) {
                if (ct) B::template body_gen_half<true>(i, n, h, k, th, part.data());
                else B::template body_gen_half<false>(i, n, h, k, th, part.data());
            }
        }
        int need = (n + B::EPT - 1) / B::EPT;
        if (nthreads < need) nthreads = need;
        for (int t = 0; t < nthreads; t++)
            B::body_normalize(t, nthreads, n, part.data(), NORM_SEC1, compress, out, nullptr, nullptr, part.data() + (size_t)3 * L * n);
    }
    static void batch_normalize(int n, const u8* xyz, u8* xy, u8* inf, int nthreads) {
        std::vector<u32> proj((size_t)3 * L * n);
        for (int i = 0; i < n; i++) B::body_load_proj(i, n, xyz, proj.data(), nullptr);
        normalize(n, proj.data(), NORM_XY_BYTES, 0, xy, inf, nullptr, nthreads);
    }
    static void verify(int n, const u8* q, const u8* z, const u8* rs, u8* ok) {
        static std::vector<u32> gt;
        if (gt.empty()) gt = gtab();
        for (int i = 0; i < n; i++) B::body_verify(i, n, q, z, rs, gt.data(), ok);
    }
};

#define DISPATCH(curve, CALL)                         \
    switch (curve) {                                  \
        case 0: Emu<CurveK256>::CALL; break;          \
        case 1: Emu<CurveP256>::CALL; break;          \
        case 2: Emu<CurveP384>::CALL; break;          \
        case 3: Emu<CurveSM2>::CALL; break;           \
        case 4: Emu<CurveP192>::CALL; break;          \
        case 5: Emu<CurveP224>::CALL; break;          \
        default: return -1;                           \
    }

extern "C" {
int emu_field_op(int curve, int which, int op, int n, const u8* a, const u8* b, u8* out, u8* ok) {
    DISPATCH(curve, field_op(which, op, n, a, b, out, ok));
    return 0;
}
int emu_mul_var(int curve, int ct, int n, unsigned flags, const u8* pts, const u8* inf, const u8* k, u8* out, int compress, u8* invalid, int nthreads) {
    DISPATCH(curve, mul_var(ct, n, flags, pts, inf, k, out, compress, invalid, nthreads));
    return 0;
}
int emu_mul_gen(int curve, int ct, int n, const u8* k, u8* out, int compress) {
    DISPATCH(curve, mul_gen(ct, n, k, out, compress));
    return 0;
}
int emu_mul_gen2(int curve, int ct, int n, const u8* k, u8* out, int compress, int nthreads) {
    DISPATCH(curve, mul_gen2(ct, n, k, out, compress, nthreads));
    return 0;
}
int emu_batch_normalize(int curve, int n, const u8* xyz, u8* xy, u8* inf, int nthreads) {
    DISPATCH(curve, batch_normalize(n, xyz, xy, inf, nthreads));
    return 0;
}
int emu_verify2(int curve, int n, const u8* q, const u8* z, const u8* rs, u8* ok, int nthreads) {
    DISPATCH(curve, verify2(n, q, z, rs, ok, nthreads));
    return 0;
}
int emu_verify_keytab(int curve, int mode, int n, const u8* q, const u8* z, const u8* rs, u8* ok, int chunk, int cap, int nthreads) {
    int r = 0;
    switch (curve) {
        case 0: r = Emu<CurveK256>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        case 1: r = Emu<CurveP256>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        case 2: r = Emu<CurveP384>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        case 3: r = Emu<CurveSM2>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        case 4: r = Emu<CurveP192>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        case 5: r = Emu<CurveP224>::verify_keytab(mode, n, q, z, rs, ok, chunk, cap, nthreads); break;
        default: return -1;
    }
    return r;
}
int emu_mul_var_fast(int curve, int n, const u8* pts, const u8* inf, const u8* k, u8* out, int compress, u8* invalid) {
    DISPATCH(curve, mul_var_fast(n, pts, inf, k, out, compress, invalid));
    return 0;
}
int emu_verify_mode(int curve, int mode, int n, const u8* q, const u8* z, const u8* rs, const u8* aux, u8* ok, u8* out, int compress, int nthreads) {
    if (mode == VM_SCHNORR && curve != 0) return -1;
    if (mode == VM_SM2DSA && curve != 3) return -1;
    DISPATCH(curve, verify_mode(mode, n, q, z, rs, aux, ok, out, compress, nthreads));
    return 0;
}
int emu_decode(int curve, int n, int mode, const u8* enc, int stride, u8* xy, u8* status) {
    DISPATCH(curve, decode(n, mode, enc, stride, xy, status));
    return 0;
}
int emu_sign(int curve, int n, const u8* d, const u8* k, const u8* z, u8* rs, u8* recid, u8* ok, int nthreads) {
    DISPATCH(curve, sign(n, d, k, z, rs, recid, ok, nthreads));
    return 0;
}
int emu_verify(int curve, int n, const u8* q, const u8* z, const u8* rs, u8* ok) {
    DISPATCH(curve, verify(n, q, z, rs, ok));
    return 0;
}
}
