// mont.cuh — word-by-word Montgomery fields on 32-bit limbs (L = 8 or 12), always reduced.
//
// One template serves every modulus the primeorder path needs:
//   base fields  P-256 (p256/src/arithmetic/field.rs:240-319, hand-written 4x64 Montgomery),
//                P-384 / SM2 (fiat-crypto word-by-word Montgomery, p384/src/arithmetic/field/p384_64.rs:146,
//                sm2/src/arithmetic/field/sm2_64.rs via primeorder/src/field.rs:48-268);
//   scalar fields of all four curves (the reference uses Barrett for p256, p256/src/arithmetic/scalar/
//                scalar64.rs:39-62, and wide reduction for k256, k256/src/arithmetic/scalar/wide32.rs — results
//                are canonical integers either way, so a Montgomery multiplier is observably identical).
// Parameters come from a generated struct P (curve_consts.cuh): L, p(i), n0 = -p^-1 mod 2^32,
// one(i) = R mod p, r2(i) = R^2 mod p.
#pragma once
#include "bigint.cuh"
#include "safegcd.cuh"
#include "fp_k256.cuh"   // Fe<L>

namespace ecb {

// true for the sparse primes whose reduction runs column by column (not ASC_OK): P-384, SM2
template <class P, bool S = P::SPARSE> struct SparseColumnwise { static constexpr bool value = false; };
template <class P> struct SparseColumnwise<P, true> { static constexpr bool value = !P::ASC_OK; };

// SQSPLIT: the squaring is issued as two calls (product, reduction) like the multiplication of the column-wise curves.
// A per-kernel choice (kernels_impl.cuh SplitSqr): see the note at sqr().
template <class P, bool SQSPLIT = false> struct Mont {
    static constexpr int L = P::L;
    typedef Fe<P::L> E;
    static constexpr bool MONT = true;
    typedef P Params;

    ECB_DEV static void set_zero(E& r) { zero_n<L>(r.v); }
    ECB_DEV static void set_one(E& r) {
        ECB_UNROLL
        for (int i = 0; i < L; i++) r.v[i] = P::one(i);
    }
    ECB_DEV static bool is_zero(const E& a) { return is_zero_n<L>(a.v); }
    ECB_DEV static bool eq(const E& a, const E& b) { return eq_n<L>(a.v, b.v); }
    ECB_DEV static void cmov(E& r, const E& a, u32 mask) { cmov_n<L>(r.v, a.v, mask); }
    ECB_DEV static void select(E& r, bool c, const E& a, const E& b) { select_n<L>(r.v, c, a.v, b.v); }

    // v + carry*2^(32L) in [0, 2p) -> [0, p)
    ECB_DEV static void final_sub(u32* r, const u32* v, u32 carry) {
        u32 u[L];
        u[0] = sub_cc(v[0], P::p(0));
        ECB_UNROLL
        for (int i = 1; i < L; i++) u[i] = subc_cc(v[i], P::p(i));
        u32 bw = subc(0u, 0u) & 1u;
        select_n<L>(r, (carry != 0) || (bw == 0), u, v);
    }

    // Montgomery reduction for the NIST-style primes p = 2^(32L) + sum_k s_k 2^(32 e_k) - 1 (P-256, P-384, SM2):
    // n0 = 1, so the row multipliers are the running low limbs themselves and Q*p is a handful of
    // limb-shifted copies of Q = sum_i m_i 2^(32 i).  No multiplier-pipe instruction is issued at all:
    //   low half:   Q = t_lo + sum_k s_k (Q << 32 e_k)  (mod R), discovered limb by limb;
    //   high half:  r = t_hi + Q + sum_k s_k (Q >> 32 (L - e_k)) + (carry out of the low half).
    // ASC_OK primes (every pair of middle exponents sums to >= L: P-256) run the low half as whole carry
    // chains, one per term in ascending exponent order, and continue each chain into the high half; the
    // others (P-384, SM2) find Q column by column with a signed two-word accumulator.
    ECB_DEV static void redc_sparse(u32* r, u32* t) {
        u32 h[L], top = 0;
        ECB_UNROLL
        for (int i = 0; i < L; i++) h[i] = t[L + i];
        if constexpr (P::ASC_OK) {
            u32 cs[P::NT];
            ECB_UNROLL
            for (int k = 0; k < P::NT; k++) {
                const int e = P::te(k);
                if (P::ts(k) > 0) {
                    t[e] = add_cc(t[e], t[0]);
                    ECB_UNROLL
                    for (int i = e + 1; i < L; i++) t[i] = addc_cc(t[i], t[i - e]);
                    cs[k] = addc(0u, 0u);
                } else {
                    t[e] = sub_cc(t[e], t[0]);
                    ECB_UNROLL
                    for (int i = e + 1; i < L; i++) t[i] = subc_cc(t[i], t[i - e]);
                    cs[k] = subc(0u, 0u) & 1u;
                }
            }
            // t[0..L) is Q now; continue every chain through the high half.  The first chain also carries -p (its addends
            // above limb e are otherwise zero; P::foldc holds the constants, the missing +1 at limb 0 enters as the carry-in
            // of the "+ Q" chain below): the value left in (top : h) is (t + Q p) / R - p, in [-p, p), and the final
            // correction is a masked addition of p driven by the sign word instead of a comparison against p.
            ECB_UNROLL
            for (int k = 0; k < P::NT; k++) {
                const int e = P::te(k);
                if (P::ts(k) > 0) {
                    add_cc(cs[k], 0xFFFFFFFFu);                      // carry flag := saved carry
                    ECB_UNROLL
                    for (int j = 0; j < L; j++) h[j] = addc_cc(h[j], j < e ? t[L - e + j] : (k == 0 ? P::foldc(j) : 0u));
                    top = addc(top, k == 0 ? P::foldc_top : 0u);
                } else {
                    sub_cc(0u, cs[k]);                               // borrow flag := saved borrow
                    ECB_UNROLL
                    for (int j = 0; j < L; j++) h[j] = subc_cc(h[j], j < e ? t[L - e + j] : (k == 0 ? P::foldc(j) : 0u));
                    top = subc(top, k == 0 ? P::foldc_top : 0u);
                }
            }
            // + Q * 2^(32L) + 1
            add_cc(1u, 0xFFFFFFFFu);                                 // carry flag := 1
            h[0] = addc_cc(h[0], t[0]);
            ECB_UNROLL
            for (int j = 1; j < L; j++) h[j] = addc_cc(h[j], t[j]);
            top = addc(top, 0u);                                     // 0 or all ones
            r[0] = add_cc(h[0], masked_p(0, top));
            ECB_UNROLL
            for (int j = 1; j < L; j++) r[j] = addc_cc(h[j], masked_p(j, top));
            return;
        } else {
            // Column by column with a signed 64-bit accumulator in plain integer arithmetic (no carry-flag chains: the
            // compiler is free to use three-input adds, and no long-lived carry predicates compete with the multiplier's -
            // with pass-wise chains here ptxas ran out of predicate registers for L = 12 and spilled them into LOP3 bit
            // twiddling, 109 extra instructions per P-384 multiplication).
            //   column i <  L:  acc += t[i] + sum_k s_k Q[i - e_k];             Q[i]   = low word, acc >>= 32
            //   column i >= L:  acc += t[i] + Q[i - L] + sum_k s_k Q[i - e_k];  r[i-L] = low word, acc >>= 32   (Q indices < L only)
            constexpr bool SIGN_FIX = L <= 8;
            long long acc = 0;
            ECB_UNROLL
            for (int i = 0; i < 2 * L; i++) {
                if (i < P::te(0)) continue;                          // no term reaches these limbs: Q[i] = t[i], carry 0
                acc += (long long)(u64)t[i];
                if (i >= L) acc += (long long)(u64)t[i - L];
                ECB_UNROLL
                for (int k = 0; k < P::NT; k++) {
                    const int src = i - P::te(k);
                    if (src < 0 || src >= L) continue;
                    if (P::ts(k) > 0) acc += (long long)(u64)t[src]; else acc -= (long long)(u64)t[src];
                }
                // - p folded into the columns: the result below is (t + Q p) / R - p, in [-p, p).  Measured: +2 % for SM2,
                // -1..2 % for P-384 (twelve more three-input adds than the select pass saves), so only the 8-limb curves do it.
                if (SIGN_FIX && i >= L) acc -= (long long)(u64)P::p(i - L);
                if (i < L) t[i] = (u32)acc; else h[i - L] = (u32)acc;
                acc >>= 32;
            }
            if constexpr (SIGN_FIX) {
                // sign word 0 or all ones: add p back when negative (no comparison against p, no select pass)
                top = (u32)acc;
                r[0] = add_cc(h[0], masked_p(0, top));
                ECB_UNROLL
                for (int j = 1; j < L; j++) r[j] = addc_cc(h[j], masked_p(j, top));
            } else {
                final_sub(r, h, (u32)acc);
            }
        }
    }

    // p = R - 2^(32e) + 1 (P-224, R = 2^224, e = 3; n0 = -1).  With Q' = t_lo * p^-1 mod R the reduced value is
    // (t - Q' p) / R, in (-p, p).  p = 1 - 2^(32e) (mod R) makes Q' the fixed point of Q' = t_lo + (Q' << 32e) (mod R),
    // found limb by limb with one ascending carry chain; then t - Q' p = t - Q' R + Q' 2^(32e) - Q', whose low half is
    // (carry of that chain) * R, so  r = t_hi + (Q' >> 32(L-e)) + carry - Q',  plus p when negative.  No multiplier-pipe
    // instruction and no final comparison against p.
    template <class PP = P> ECB_DEV static void redc_negsparse(u32* r, u32* t) {
        constexpr int e = PP::NE;
        t[e] = add_cc(t[e], t[0]);
        ECB_UNROLL
        for (int i = e + 1; i < L; i++) t[i] = addc_cc(t[i], t[i - e]);
        const u32 cs = addc(0u, 0u);                                   // t[0..L) is Q' now
        u32 s[L], d[L];
        add_cc(cs, 0xFFFFFFFFu);                                       // carry flag := cs
        ECB_UNROLL
        for (int j = 0; j < L; j++) s[j] = addc_cc(t[L + j], j < e ? t[L - e + j] : 0u);
        const u32 sc = addc(0u, 0u);
        d[0] = sub_cc(s[0], t[0]);
        ECB_UNROLL
        for (int j = 1; j < L; j++) d[j] = subc_cc(s[j], t[j]);
        const u32 bw = subc(0u, 0u);                                   // all ones on borrow
        const u32 m = bw & ~((u32)0 - sc);                             // negative: borrow not covered by the carry of s
        r[0] = add_cc(d[0], PP::p(0) & m);
        ECB_UNROLL
        for (int j = 1; j < L; j++) r[j] = addc_cc(d[j], PP::p(j) & m);
    }
    // Montgomery reduction of t[0..2L): r = t * R^-1 mod p   (t < p * R)
    ECB_DEV static void redc(u32* r, u32* t) {
        if constexpr (P::SPARSE) redc_sparse(r, t);
        else if constexpr (P::NEGSPARSE) redc_negsparse(r, t);
        else redc_generic(r, t);
    }
    // odd limb counts (P-224): row by row with a 64-bit running carry; a row's carry-out (weight L + i) never feeds a later
    // multiplier m_i, so it is parked in cy[] and added once at the end
    ECB_DEV static void redc_rows(u32* r, u32* t) {
        u32 cy[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) {
            const u32 m = t[i] * P::n0;
            u64 c = 0;
            ECB_UNROLL
            for (int j = 0; j < L; j++) {
                c += (u64)m * P::p(j) + t[i + j];
                t[i + j] = (u32)c;
                c >>= 32;
            }
            cy[i] = (u32)c;
        }
        u32 v[L];
        v[0] = add_cc(t[L], cy[0]);
        ECB_UNROLL
        for (int i = 1; i < L; i++) v[i] = addc_cc(t[L + i], cy[i]);
        const u32 top = addc(0u, 0u);
        final_sub(r, v, top);
    }
    ECB_DEV static void redc_generic(u32* r, u32* t) {
        if constexpr (L % 2 != 0) { redc_rows(r, t); return; }
        // Row i adds m_i * p at limb i as two aligned-pair carry chains (even j, odd j).  The chain
        // carry-outs (weight i+L and i+L+1) never feed a later m_i, so they are collected in cy[]
        // and added once at the end instead of being rippled to the top in every row.
        u32 cy[L + 1];
        ECB_UNROLL
        for (int i = 0; i <= L; i++) cy[i] = 0;
        ECB_UNROLL
        for (int i = 0; i < L; i++) {
            const u32 m = t[i] * P::n0;
            ECB_UNROLL
            for (int j = 0; j < L; j += 2) {
                t[i + j] = (j == 0) ? madlo_cc(m, P::p(j), t[i + j]) : madloc_cc(m, P::p(j), t[i + j]);
                t[i + j + 1] = madhic_cc(m, P::p(j), t[i + j + 1]);
            }
            cy[i] = addc(cy[i], 0u);            // weight L + i
            ECB_UNROLL
            for (int j = 1; j < L; j += 2) {
                t[i + j] = (j == 1) ? madlo_cc(m, P::p(j), t[i + j]) : madloc_cc(m, P::p(j), t[i + j]);
                t[i + j + 1] = madhic_cc(m, P::p(j), t[i + j + 1]);
            }
            cy[i + 1] = addc(cy[i + 1], 0u);    // weight L + i + 1
        }
        u32 v[L];
        v[0] = add_cc(t[L], cy[0]);
        ECB_UNROLL
        for (int i = 1; i < L; i++) v[i] = addc_cc(t[L + i], cy[i]);
        u32 top = addc(cy[L], 0u);
        final_sub(r, v, top);
    }

    // real functions with register arguments (see fp_k256.cuh for why)
    ECB_DEV static void mul_body(E& r, const E& a, const E& b) {
        u32 t[2 * L];
        mul_wide<L>(t, a.v, b.v);
        redc(r.v, t);
    }
    ECB_DEV static void sqr_body(E& r, const E& a) {
        u32 t[2 * L];
        sqr_wide<L>(t, a.v);
        redc(r.v, t);
    }
    ECB_FIELD_FN static E mul_fn(E a, E b) { E r; mul_body(r, a, b); return r; }
    ECB_FIELD_FN static E sqr_fn(E a) { E r; sqr_body(r, a); return r; }
    // Two-call form for the column-wise sparse reduction (P-384, SM2): product and reduction as separate functions.
    // Inside one function ptxas interleaves the reduction's columns with the multiplier's row chains, runs out of the
    // seven predicate registers for the carries and spills them into LOP3 bit twiddling (442 of 757 instructions of the
    // P-384 multiplication); across a call boundary it cannot.  The 2L-limb product travels in registers.
    struct Wide { u32 v[2 * L]; };
    static constexpr bool SPLIT = SparseColumnwise<P>::value;
    ECB_FIELD_FN static Wide mulw_fn(E a, E b) { Wide t; mul_wide<L>(t.v, a.v, b.v); return t; }
    ECB_FIELD_FN static E redc_fn(Wide t) { E r; redc(r.v, t.v); return r; }
#ifdef ECB_FIELD_OUT_PTR   // experiment (bench/pointloop.cu): operands by value, result stored through a pointer by the callee
    ECB_FIELD_FN static void mul_ofn(E* r, E a, E b) { E t; mul_body(t, a, b); *r = t; }
    ECB_FIELD_FN static void sqr_ofn(E* r, E a) { E t; sqr_body(t, a); *r = t; }
    ECB_DEV static void mul(E& r, const E& a, const E& b) { mul_ofn(&r, a, b); }
    ECB_DEV static void sqr(E& r, const E& a) { sqr_ofn(&r, a); }
#else
    ECB_DEV static void mul(E& r, const E& a, const E& b) {
        if constexpr (SPLIT) { Wide t = mulw_fn(a, b); r = redc_fn(t); }
        else r = mul_fn(a, b);
    }
    // The squaring stays one function by default.  Splitting it like mul costs P-384 3-6 % where the one-function form
    // compiles well (the public-input kernels); whether it hits the predicate spills depends on the kernel it is compiled
    // into (ptxas allocates registers across the call graph) - the P-384 secret-scalar kernels do (367 instead of 272
    // instructions) and take the split form, 7 % faster there.  Check `tools/sass_funcs.py` for P2R / LOP3 in the
    // squarer after touching a kernel that calls it.
    ECB_FIELD_FN static Wide sqrw_fn(E a) { Wide t; sqr_wide<L>(t.v, a.v); return t; }
    ECB_DEV static void sqr(E& r, const E& a) {
        if constexpr (SQSPLIT) { Wide t = sqrw_fn(a); r = redc_fn(t); }
        else r = sqr_fn(a);
    }
#endif
    ECB_DEV static void add(E& r, const E& a, const E& b) {
        u32 v[L];
        u32 c = add_n<L>(v, a.v, b.v);
        final_sub(r.v, v, c);
    }
    // p(i) & m for an all-ones / zero mask m; the sparse primes only have limbs 0, 1, 2^32-2, 2^32-1
    ECB_DEV static u32 masked_p(int i, u32 m) {
        return P::p(i) == 0u ? 0u : P::p(i) == 0xFFFFFFFFu ? m : P::p(i) == 1u ? (m >> 31) : P::p(i) == 0xFFFFFFFEu ? (m << 1) : (P::p(i) & m);
    }
    ECB_DEV static void sub(E& r, const E& a, const E& b) {
        u32 v[L];
        v[0] = sub_cc(a.v[0], b.v[0]);
        ECB_UNROLL
        for (int i = 1; i < L; i++) v[i] = subc_cc(a.v[i], b.v[i]);
        const u32 m = subc(0u, 0u);                 // 0xFFFFFFFF on borrow, else 0
        if constexpr (P::SPARSE) {
            // + p on borrow with mask-derived operands: no select pass
            r.v[0] = add_cc(v[0], masked_p(0, m));
            ECB_UNROLL
            for (int i = 1; i < L; i++) r.v[i] = addc_cc(v[i], masked_p(i, m));
        } else {
            u32 u[L];
            u[0] = add_cc(v[0], P::p(0));
            ECB_UNROLL
            for (int i = 1; i < L; i++) u[i] = addc_cc(v[i], P::p(i));
            select_n<L>(r.v, m != 0, u, v);
        }
    }
    ECB_DEV static void neg(E& r, const E& a) {
        E z;
        set_zero(z);
        sub(r, z, a);
    }
    ECB_DEV static void dbl(E& r, const E& a) { add(r, a, a); }
    // r = a / 2 (Montgomery form is linear, so this is the plain halving): (a + (a odd ? p : 0)) >> 1
    ECB_DEV static void half(E& r, const E& a) {
        const u32 m = (u32)0 - (a.v[0] & 1u);
        u32 t[L];
        t[0] = add_cc(a.v[0], masked_p(0, m));
        ECB_UNROLL
        for (int i = 1; i < L; i++) t[i] = addc_cc(a.v[i], masked_p(i, m));
        const u32 c = addc(0u, 0u);
        ECB_UNROLL
        for (int i = 0; i < L - 1; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
        r.v[L - 1] = (t[L - 1] >> 1) | (c << 31);
    }
    // small constant multiples by addition chains (only 2,3,4,8 are needed by the formulas)
    ECB_DEV static void mul_small(E& r, const E& a, u32 k) {
        E t;
        if (k == 2) { dbl(r, a); }
        else if (k == 3) { dbl(t, a); add(r, t, a); }
        else if (k == 4) { dbl(t, a); dbl(r, t); }
        else if (k == 8) { dbl(t, a); dbl(t, t); dbl(r, t); }
        else {   // generic double-and-add, k >= 1
            E acc = a;
            int top = 31;
            while (!((k >> top) & 1)) top--;
            for (int i = top - 1; i >= 0; i--) { dbl(acc, acc); if ((k >> i) & 1) add(acc, acc, a); }
            r = acc;
        }
    }

    // plain integer limbs (< p) -> Montgomery form; false when the input is >= p
    ECB_DEV static bool from_limbs(E& r, const u32* v) {
        u32 pp[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) pp[i] = P::p(i);
        bool ok = !geq_n<L>(v, pp);
        E a, r2;
        copy_n<L>(a.v, v);
        ECB_UNROLL
        for (int i = 0; i < L; i++) r2.v[i] = P::r2(i);
        mul(r, a, r2);
        return ok;
    }
    // Montgomery form -> canonical integer limbs
    ECB_DEV static void to_limbs(u32* v, const E& a) {
        u32 t[2 * L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) { t[i] = a.v[i]; t[L + i] = 0; }
        redc(v, t);
    }
    ECB_DEV static bool is_odd(const E& a) {   // canonical LSB (primeorder/src/field.rs:169-172)
        u32 v[L];
        to_limbs(v, a);
        return v[0] & 1;
    }
    // mixed-domain product: plain x times Montgomery-form y gives plain x*y (mod p)
    ECB_DEV static void mul_plain(u32* r, const u32* x_plain, const E& y_mont) {
        E a, o;
        copy_n<L>(a.v, x_plain);
        mul(o, a, y_mont);
        copy_n<L>(r, o.v);
    }

    ECB_DEV static void sqr_n(E& r, const E& a, int n) {
        r = a;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < n; i++) sqr(r, r);
    }

    // r = a^e for the public exponent e = p - 2 - (extra) given as limbs; 4-bit fixed window.
    ECB_DEV static void pow_limbs(E& r, const E& a, const u32* e) {
        E tab[16];
        set_one(tab[0]);
        tab[1] = a;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 2; i < 16; i++) mul(tab[i], tab[i - 1], a);
        E acc;
        set_one(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int w = 8 * L - 1; w >= 0; w--) {
            u32 nib = (e[w >> 3] >> ((w & 7) * 4)) & 15u;
            sqr(acc, acc); sqr(acc, acc); sqr(acc, acc); sqr(acc, acc);
            if (nib) mul(acc, acc, tab[nib]);   // exponent is public and identical across the warp
        }
        r = acc;
    }
    // a^(p-2) (Fermat).  0 -> 0.  The reference uses Fermat for p256 (field.rs:364-382) and
    // Bernstein-Yang for p384/sm2 (primeorder/src/field.rs:506-559); the result is the same element.
    ECB_DEV static void inv(E& r, const E& a) {
        u32 e[L];
        e[0] = sub_cc(P::p(0), 2u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) e[i] = subc_cc(P::p(i), 0u);
        pow_limbs(r, a, e);
    }
    // The same inverse by Bernstein-Yang divsteps (safegcd.cuh): a quarter of Fermat's dependent path, constant time.  The
    // stored value aR is inverted as a plain integer, (aR)^-1 = a^-1 R^-1, and two Montgomery products with R^2 bring it
    // to a^-1 R.  inv_trick is what the one-chain-per-CTA Montgomery-trick kernels call (-DECB_SAFEGCD=0: Fermat, for A/B).
    ECB_DEV static void inv_gcd(E& r, const E& a) {
        u32 pp[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) pp[i] = P::p(i);
        E t, r2;
        SafeGcd<L>::inv(t.v, a.v, pp);
        ECB_UNROLL
        for (int i = 0; i < L; i++) r2.v[i] = P::r2(i);
        mul(t, t, r2);
        mul(r, t, r2);
    }
    ECB_DEV static void inv_trick(E& r, const E& a) {
#if ECB_SAFEGCD
        inv_gcd(r, a);
#else
        inv(r, a);
#endif
    }
    // p = 1 (mod 4) (P-224: p - 1 = 2^96 (2^128 - 1)): Tonelli-Shanks, variable time - square roots are only taken of public
    // data (point decompression).  The reference runs the constant-time form of the same algorithm
    // (p224/src/arithmetic/field.rs:103-233, S = 96, ROOT_OF_UNITY = 22^t); the root it returns is +-this one and the caller
    // picks the sign by parity, so the decoded point is the same.  A non-residue returns garbage (callers check r^2 == a).
    template <class PP = P> ECB_DEV static void sqrt_tonelli_shanks(E& r, const E& a) {
        if (is_zero(a)) { r = a; return; }
        u32 e[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) e[i] = PP::ts_half_t(i);     // (t - 1) / 2, t = (p - 1) / 2^S odd
        E w, x, b, z;
        pow_limbs(w, a, e);
        mul(x, a, w);                                            // a^((t+1)/2)
        mul(b, x, w);                                            // a^t
        ECB_UNROLL
        for (int i = 0; i < L; i++) z.v[i] = PP::ts_root(i);     // g^t for a non-residue g: order 2^S
        int v = PP::TS_S;
        E one;
        set_one(one);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        while (!eq(b, one)) {
            int k = 0;
            E t2 = b;
            while (k < v && !eq(t2, one)) { sqr(t2, t2); k++; }  // least k with b^(2^k) = 1
            if (k >= v) { r = x; return; }                       // no such k below v: a is not a square
            E zz = z;
            for (int i = 0; i < v - k - 1; i++) sqr(zz, zz);
            mul(x, x, zz);
            sqr(z, zz);
            mul(b, b, z);
            v = k;
        }
        r = x;
    }
    // a^((p+1)/4) for p = 3 mod 4 (p256 field.rs:385-411, p384 field.rs:95-117, sm2 field.rs sqrt)
    ECB_DEV static void sqrt_candidate(E& r, const E& a) {
        if constexpr (P::SQRT_TS) { sqrt_tonelli_shanks(r, a); return; }
        u32 e[L];
        e[0] = add_cc(P::p(0), 1u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) e[i] = addc_cc(P::p(i), 0u);
        u32 c = addc(0u, 0u);
        ECB_UNROLL
        for (int i = 0; i < L - 1; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
        e[L - 1] = (e[L - 1] >> 2) | (c << 30);
        pow_limbs(r, a, e);
    }
};

}  // namespace ecb
