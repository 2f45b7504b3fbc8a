// fp_k256.cuh — secp256k1 base field, 8x32-bit limbs, register resident, always canonical.
//
// Replaces k256::FieldElement (5x52 lazy limbs, k256/src/arithmetic/field/field_5x52.rs:288-449,
// and the 8x32 risc0 backend field_8x32_risc0.rs:139-193) for the batched path.  p = 2^256 - C with
// C = 2^32 + 977 (field_5x52.rs:136), so a 512-bit product folds as lo + hi*C with one row of
// IMAD.WIDE by 977 plus a shifted add; a second tiny fold and one conditional subtraction give the
// canonical residue.  All values are kept fully reduced (like FieldElement8x32R0), so there is no
// magnitude bookkeeping (field_impl.rs) anywhere on the device.
#pragma once
#include "bigint.cuh"
#include "safegcd.cuh"

namespace ecb {

template <int L_> struct Fe { u32 v[L_]; };

#ifndef ECB_FIELD_FN   // (bench/pointloop.cu overrides this to compare call strategies)
#if defined(__CUDACC__) && !defined(ECB_EMU)
#define ECB_FIELD_FN __device__ __noinline__
#else
#define ECB_FIELD_FN inline
#endif
#endif

struct FpK256 {
    static constexpr int L = 8;
    typedef Fe<8> E;
    static constexpr bool MONT = false;

    ECB_HD static constexpr u32 p(int i) {
        constexpr u32 t[8] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        return t[i];
    }

    ECB_DEV static void set_zero(E& r) { zero_n<8>(r.v); }
    ECB_DEV static void set_one(E& r) { zero_n<8>(r.v); r.v[0] = 1; }
    ECB_DEV static void set_u32(E& r, u32 k) { zero_n<8>(r.v); r.v[0] = k; }
    ECB_DEV static bool is_zero(const E& a) { return is_zero_n<8>(a.v); }
    ECB_DEV static bool eq(const E& a, const E& b) { return eq_n<8>(a.v, b.v); }
    ECB_DEV static bool is_odd(const E& a) { return a.v[0] & 1; }
    ECB_DEV static void cmov(E& r, const E& a, u32 mask) { cmov_n<8>(r.v, a.v, mask); }
    ECB_DEV static void select(E& r, bool c, const E& a, const E& b) { select_n<8>(r.v, c, a.v, b.v); }

    // v (8 limbs) + carry*2^256, known < 2^256 + 2^68  ->  canonical
    ECB_DEV static void final_sub(u32* r, const u32* v, u32 carry) {
        u32 u[8];
        u[0] = add_cc(v[0], 977u);
        u[1] = addc_cc(v[1], 1u);
        ECB_UNROLL
        for (int i = 2; i < 8; i++) u[i] = addc_cc(v[i], 0u);
        u32 c2 = addc(0u, 0u);
        select_n<8>(r, (carry | c2) != 0, u, v);
    }

    // r = (acc[0..8) + top * 2^256) mod p  with top < 2^32 (+ top2 * 2^288, top2 in {0,1})
    ECB_DEV static void fold_top(u32* r, const u32* acc, u32 top, u32 top2) {
        u32 g0 = top * 977u;
        u32 g1 = madhi_cc(top, 977u, top);
        u32 g2 = addc(0u, 0u);
        g1 = add_cc(g1, top2 * 977u);
        g2 = addc(g2, top2);
        u32 v[8];
        v[0] = add_cc(acc[0], g0);
        v[1] = addc_cc(acc[1], g1);
        v[2] = addc_cc(acc[2], g2);
        ECB_UNROLL
        for (int i = 3; i < 8; i++) v[i] = addc_cc(acc[i], 0u);
        u32 carry = addc(0u, 0u);
        final_sub(r, v, carry);
    }

    // r = t[0..16) mod p
    ECB_DEV static void reduce512(u32* r, const u32* t) {
        const u32* lo = t;
        const u32* hi = t + 8;
        u32 e[9], o[9];
        // e = lo + sum_{j even} hi[j]*977*2^(32j)
        ECB_UNROLL
        for (int j = 0; j < 8; j += 2) {
            e[j] = (j == 0) ? madlo_cc(hi[j], 977u, lo[j]) : madloc_cc(hi[j], 977u, lo[j]);
            e[j + 1] = madhic_cc(hi[j], 977u, lo[j + 1]);
        }
        e[8] = addc(0u, 0u);
        // o (weight +1) = sum_{j odd} hi[j]*977*2^(32(j-1)) + hi      (the hi << 32 term)
        ECB_UNROLL
        for (int j = 1; j < 8; j += 2) {
            o[j - 1] = (j == 1) ? madlo_cc(hi[j], 977u, hi[j - 1]) : madloc_cc(hi[j], 977u, hi[j - 1]);
            o[j] = madhic_cc(hi[j], 977u, hi[j]);
        }
        o[8] = addc(0u, 0u);
        u32 acc[10];
        acc[0] = e[0];
        acc[1] = add_cc(e[1], o[0]);
        ECB_UNROLL
        for (int i = 2; i < 9; i++) acc[i] = addc_cc(e[i], o[i - 1]);
        acc[9] = addc(0u, o[8]);
        fold_top(r, acc, acc[8], acc[9]);
    }

    // The multiplier and squarer are real functions with by-value (register) arguments: one ~150
    // instruction copy each per kernel instead of one per call site, so the window loops of the
    // point kernels stay inside the instruction cache (ncu showed "no instruction" as the top stall
    // of the fully inlined build, profiles/).  ptxas passes the 8-limb structs in registers: no
    // local-memory traffic at the call.
    ECB_DEV static void mul_body(E& r, const E& a, const E& b) {
        u32 t[16];
        mul_wide<8>(t, a.v, b.v);
        reduce512(r.v, t);
    }
    ECB_DEV static void sqr_body(E& r, const E& a) {
        u32 t[16];
        sqr_wide<8>(t, a.v);
        reduce512(r.v, t);
    }
    ECB_FIELD_FN static E mul_fn(E a, E b) { E r; mul_body(r, a, b); return r; }
    ECB_FIELD_FN static E sqr_fn(E a) { E r; sqr_body(r, a); return r; }
#ifdef ECB_FIELD_OUT_PTR   // experiment (bench/pointloop.cu): operands by value, result stored through a pointer by the callee
    ECB_FIELD_FN static void mul_ofn(E* r, E a, E b) { E t; mul_body(t, a, b); *r = t; }
    ECB_FIELD_FN static void sqr_ofn(E* r, E a) { E t; sqr_body(t, a); *r = t; }
    ECB_DEV static void mul(E& r, const E& a, const E& b) { mul_ofn(&r, a, b); }
    ECB_DEV static void sqr(E& r, const E& a) { sqr_ofn(&r, a); }
#else
    ECB_DEV static void mul(E& r, const E& a, const E& b) { r = mul_fn(a, b); }
    ECB_DEV static void sqr(E& r, const E& a) { r = sqr_fn(a); }
#endif
    ECB_DEV static void add(E& r, const E& a, const E& b) {
        u32 v[8];
        u32 c = add_n<8>(v, a.v, b.v);
        final_sub(r.v, v, c);
    }
    ECB_DEV static void sub(E& r, const E& a, const E& b) {
        // a - b, then + p (= - C mod 2^256) when it borrowed: the correction operands are the borrow mask
        // itself, so there is no select pass (19 instructions instead of 25)
        u32 v[8];
        v[0] = sub_cc(a.v[0], b.v[0]);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) v[i] = subc_cc(a.v[i], b.v[i]);
        const u32 m = subc(0u, 0u);                 // 0xFFFFFFFF on borrow, else 0
        r.v[0] = sub_cc(v[0], m & 977u);
        r.v[1] = subc_cc(v[1], m >> 31);
        ECB_UNROLL
        for (int i = 2; i < 8; i++) r.v[i] = subc_cc(v[i], 0u);
    }
    ECB_DEV static void neg(E& r, const E& a) {
        E z;
        set_zero(z);
        sub(r, z, a);
    }
    ECB_DEV static void dbl(E& r, const E& a) { add(r, a, a); }
    // r = a / 2 = (a + (a odd ? p : 0)) >> 1: the masked addition leaves a 257-bit value whose bit 256 is the carry
    ECB_DEV static void half(E& r, const E& a) {
        const u32 m = (u32)0 - (a.v[0] & 1u);
        u32 t[8];
        t[0] = add_cc(a.v[0], m & p(0));
        t[1] = addc_cc(a.v[1], m & p(1));
        ECB_UNROLL
        for (int i = 2; i < 8; i++) t[i] = addc_cc(a.v[i], m);
        const u32 c = addc(0u, 0u);
        ECB_UNROLL
        for (int i = 0; i < 7; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
        r.v[7] = (t[7] >> 1) | (c << 31);
    }
    // r = a * k for a small constant k (k*2^256 must stay far below 2^320: k < 2^16 here)
    ECB_DEV static void mul_small(E& r, const E& a, u32 k) {
        u32 e[9], o[9];
        ECB_UNROLL
        for (int j = 0; j < 8; j += 2) {
            e[j] = (j == 0) ? madlo_cc(a.v[j], k, 0u) : madloc_cc(a.v[j], k, 0u);
            e[j + 1] = madhic_cc(a.v[j], k, 0u);
        }
        e[8] = addc(0u, 0u);
        ECB_UNROLL
        for (int j = 1; j < 8; j += 2) {
            o[j - 1] = (j == 1) ? madlo_cc(a.v[j], k, 0u) : madloc_cc(a.v[j], k, 0u);
            o[j] = madhic_cc(a.v[j], k, 0u);
        }
        u32 acc[9];
        acc[0] = e[0];
        acc[1] = add_cc(e[1], o[0]);
        ECB_UNROLL
        for (int i = 2; i < 8; i++) acc[i] = addc_cc(e[i], o[i - 1]);
        acc[8] = addc(e[8], o[7]);
        fold_top(r.v, acc, acc[8], 0u);
    }

    ECB_DEV static void sqr_n(E& r, const E& a, int n) {
        r = a;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < n; i++) sqr(r, r);
    }

    // the same inverse by Bernstein-Yang divsteps (safegcd.cuh; values are canonical integers here, no Montgomery factor);
    // inv_trick is what the one-chain-per-CTA Montgomery-trick kernels call (-DECB_SAFEGCD=0: the Fermat chain, for A/B)
    ECB_DEV static void inv_gcd(E& r, const E& a) {
        u32 pp[8];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) pp[i] = p(i);
        SafeGcd<8>::inv(r.v, a.v, pp);
    }
    ECB_DEV static void inv_trick(E& r, const E& a) {
#if ECB_SAFEGCD
        inv_gcd(r, a);
#else
        inv(r, a);
#endif
    }
    // a^(p-2): the addition chain of k256/src/arithmetic/field.rs:187-216 (255 S + 15 M).  0 -> 0.
    ECB_DEV static void inv(E& r, const E& a) {
        E x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
        sqr(t, a); mul(x2, t, a);
        sqr(t, x2); mul(x3, t, a);
        sqr_n(t, x3, 3); mul(x6, t, x3);
        sqr_n(t, x6, 3); mul(x9, t, x3);
        sqr_n(t, x9, 2); mul(x11, t, x2);
        sqr_n(t, x11, 11); mul(x22, t, x11);
        sqr_n(t, x22, 22); mul(x44, t, x22);
        sqr_n(t, x44, 44); mul(x88, t, x44);
        sqr_n(t, x88, 88); mul(x176, t, x88);
        sqr_n(t, x176, 44); mul(x220, t, x44);
        sqr_n(t, x220, 3); mul(x223, t, x3);
        sqr_n(t, x223, 23); mul(t, t, x22);
        sqr_n(t, t, 5); mul(t, t, a);
        sqr_n(t, t, 3); mul(t, t, x2);
        sqr_n(t, t, 2); mul(r, t, a);
    }

    // a^((p+1)/4) — k256/src/arithmetic/field.rs:220-255.  Caller checks r^2 == a.
    ECB_DEV static void sqrt_candidate(E& r, const E& a) {
        E x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
        sqr(t, a); mul(x2, t, a);
        sqr(t, x2); mul(x3, t, a);
        sqr_n(t, x3, 3); mul(x6, t, x3);
        sqr_n(t, x6, 3); mul(x9, t, x3);
        sqr_n(t, x9, 2); mul(x11, t, x2);
        sqr_n(t, x11, 11); mul(x22, t, x11);
        sqr_n(t, x22, 22); mul(x44, t, x22);
        sqr_n(t, x44, 44); mul(x88, t, x44);
        sqr_n(t, x88, 88); mul(x176, t, x88);
        sqr_n(t, x176, 44); mul(x220, t, x44);
        sqr_n(t, x220, 3); mul(x223, t, x3);
        sqr_n(t, x223, 23); mul(t, t, x22);
        sqr_n(t, t, 6); mul(t, t, x2);
        sqr_n(r, t, 2);
    }

    // canonical big-endian bytes -> element; false if the value is >= p (field_5x52.rs:75-79)
    ECB_DEV static bool from_limbs(E& r, const u32* v) {
        copy_n<8>(r.v, v);
        u32 pp[8];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) pp[i] = p(i);
        return !geq_n<8>(v, pp);
    }
    ECB_DEV static void to_limbs(u32* v, const E& a) { copy_n<8>(v, a.v); }
};

}  // namespace ecb
