// ec.cuh — group law and per-element algorithms, generic over a curve descriptor C
// (curve_consts.cuh): C::F base field, C::Fn scalar field, C::A_IS_ZERO (secp256k1) or a = -3.
//
// Points are homogeneous projective (x = X/Z, y = Y/Z), identity = (0 : 1 : 0), exactly the
// representation of k256::ProjectivePoint (k256/src/arithmetic/projective.rs:38-50) and
// primeorder::ProjectivePoint (primeorder/src/projective.rs:37-41), and the formulas are the
// Renes-Costello-Batina complete ones the reference uses (projective.rs:96-274 for a = 0,
// primeorder/src/point_arithmetic.rs:209-317 for a = -3), written from the closed forms
// (SURVEY.md App. B.1), so P+P, P+(-P) and the identity need no branches.
#pragma once
#include "curve_consts.cuh"

namespace ecb {

// Point-level operations are real (non-inlined) device functions: one copy of add / add_mixed / dbl per
// kernel keeps the instruction footprint of the window loops small (the bodies are 2-6 k SASS
// instructions each) and the build time bounded; the ~100 local-memory moves per call are noise next
// to the ~14 field multiplications inside.
#ifndef ECB_POINT_FN
#if defined(__CUDACC__) && !defined(ECB_EMU)
#define ECB_POINT_FN __device__ __noinline__
#else
#define ECB_POINT_FN inline
#endif
#endif

template <class C> struct EC {
    typedef typename C::F F;
    typedef typename F::E E;
    typedef typename C::Fn Fn;
    typedef typename Fn::E S;
    static constexpr int L = C::L;
    static constexpr int FB = C::FB;

    struct Proj { E X, Y, Z; };
    struct Aff { E x, y; };

    // ---------------------------------------------------------------- constants as elements
    ECB_DEV static void const_b3(E& r) {
        ECB_UNROLL
        for (int i = 0; i < L; i++) r.v[i] = C::b3(i);
    }
    ECB_DEV static void const_b(E& r) {
        ECB_UNROLL
        for (int i = 0; i < L; i++) r.v[i] = C::b(i);
    }
    ECB_DEV static void generator(Aff& g) {
        ECB_UNROLL
        for (int i = 0; i < L; i++) { g.x.v[i] = C::gx(i); g.y.v[i] = C::gy(i); }
    }
    // r = 3b * a
    ECB_DEV static void mul_b3(E& r, const E& a) {
        if constexpr (C::A_IS_ZERO) {
            F::mul_small(r, a, C::B3_SMALL);
        } else {
            E b3;
            const_b3(b3);
            F::mul(r, a, b3);
        }
    }

    // ---------------------------------------------------------------- basic point helpers
    ECB_DEV static void set_identity(Proj& p) { F::set_zero(p.X); F::set_one(p.Y); F::set_zero(p.Z); }
    ECB_DEV static void from_affine(Proj& p, const Aff& a) { p.X = a.x; p.Y = a.y; F::set_one(p.Z); }
    ECB_DEV static bool is_identity(const Proj& p) { return F::is_zero(p.Z); }
    ECB_DEV static void cmov(Proj& r, const Proj& a, u32 mask) {
        F::cmov(r.X, a.X, mask); F::cmov(r.Y, a.Y, mask); F::cmov(r.Z, a.Z, mask);
    }
    // negate when mask is all-ones (branch free)
    ECB_DEV static void cneg(Proj& p, u32 mask) {
        E ny;
        F::neg(ny, p.Y);
        F::cmov(p.Y, ny, mask);
    }

    // ---------------------------------------------------------------- complete addition
    ECB_POINT_FN static void add(Proj& r, const Proj& p, const Proj& q) {
        E xx, yy, zz, xy, yz, xz, t0, t1;
        F::mul(xx, p.X, q.X);
        F::mul(yy, p.Y, q.Y);
        F::mul(zz, p.Z, q.Z);
        F::add(t0, p.X, p.Y); F::add(t1, q.X, q.Y); F::mul(xy, t0, t1); F::sub(xy, xy, xx); F::sub(xy, xy, yy);
        F::add(t0, p.Y, p.Z); F::add(t1, q.Y, q.Z); F::mul(yz, t0, t1); F::sub(yz, yz, yy); F::sub(yz, yz, zz);
        F::add(t0, p.X, p.Z); F::add(t1, q.X, q.Z); F::mul(xz, t0, t1); F::sub(xz, xz, xx); F::sub(xz, xz, zz);
        finish_add(r, xx, yy, zz, xy, yz, xz);
    }
    // p + (affine q); q must not be the identity
    ECB_POINT_FN static void add_mixed(Proj& r, const Proj& p, const Aff& q) {
        E xx, yy, zz, xy, yz, xz, t0, t1;
        F::mul(xx, p.X, q.x);
        F::mul(yy, p.Y, q.y);
        zz = p.Z;
        F::add(t0, p.X, p.Y); F::add(t1, q.x, q.y); F::mul(xy, t0, t1); F::sub(xy, xy, xx); F::sub(xy, xy, yy);
        F::mul(yz, q.y, p.Z); F::add(yz, yz, p.Y);
        F::mul(xz, q.x, p.Z); F::add(xz, xz, p.X);
        finish_add(r, xx, yy, zz, xy, yz, xz);
    }
    ECB_DEV static void finish_add(Proj& r, const E& xx, const E& yy, const E& zz, const E& xy, const E& yz, const E& xz) {
        E t0, t1, t2, t3;
        if constexpr (C::A_IS_ZERO) {
            // X3 = xy(yy - b3 zz) - b3 yz xz ; Y3 = (yy + b3 zz)(yy - b3 zz) + 3 xx b3 xz ; Z3 = yz(yy + b3 zz) + 3 xx xy
            E bzz, yym, yyp, byz, xx3, bxz;
            mul_b3(bzz, zz);
            F::sub(yym, yy, bzz);
            F::add(yyp, yy, bzz);
            mul_b3(byz, yz);
            mul_b3(bxz, xz);
            F::dbl(xx3, xx); F::add(xx3, xx3, xx);
            F::mul(t0, xy, yym); F::mul(t1, byz, xz); F::sub(t2, t0, t1);
            F::mul(t0, yyp, yym); F::mul(t1, xx3, bxz); F::add(t3, t0, t1);
            F::mul(t0, yz, yyp); F::mul(t1, xx3, xy);
            r.X = t2; r.Y = t3; F::add(r.Z, t0, t1);
        } else {
            // a = -3:  A = yy + 3xz - b3 zz ; B = b3 xz - 3xx - 9zz ; Cc = 3(xx - zz) ; D = yy - 3xz + b3 zz
            E bzz, bxz, xz3, A, B, Cc, D;
            mul_b3(bzz, zz);
            mul_b3(bxz, xz);
            F::dbl(xz3, xz); F::add(xz3, xz3, xz);
            F::add(A, yy, xz3); F::sub(A, A, bzz);
            F::sub(D, yy, xz3); F::add(D, D, bzz);
            F::sub(t0, xx, zz); F::dbl(Cc, t0); F::add(Cc, Cc, t0);
            F::dbl(t0, zz); F::add(t0, t0, zz);       // 3zz
            F::dbl(t1, t0); F::add(t1, t1, t0);       // 9zz
            F::dbl(t0, xx); F::add(t0, t0, xx);       // 3xx
            F::sub(B, bxz, t0); F::sub(B, B, t1);
            F::mul(t0, xy, A); F::mul(t1, yz, B); F::sub(t2, t0, t1);
            F::mul(t0, Cc, B); F::mul(t1, D, A); F::add(t3, t0, t1);
            F::mul(t0, yz, D); F::mul(t1, xy, Cc);
            r.X = t2; r.Y = t3; F::add(r.Z, t0, t1);
        }
    }
    ECB_POINT_FN static void dbl(Proj& r, const Proj& p) {
        if constexpr (C::A_IS_ZERO) {
            // X3 = 2XY(Y^2 - 9bZ^2) ; Y3 = (Y^2 - 9bZ^2)(Y^2 + 3bZ^2) + 24b Y^2 Z^2 ; Z3 = 8 Y^3 Z
            E yy, zz, xy, yz, bzz, bzz3, m, pp, t0, t1;
            F::sqr(yy, p.Y);
            F::sqr(zz, p.Z);
            F::mul(xy, p.X, p.Y);
            F::mul(yz, p.Y, p.Z);
            mul_b3(bzz, zz);
            F::dbl(bzz3, bzz); F::add(bzz3, bzz3, bzz);
            F::sub(m, yy, bzz3);
            F::add(pp, yy, bzz);
            F::mul(t0, xy, m); F::dbl(r.X, t0);
            F::mul(t0, m, pp); F::mul(t1, bzz, yy); F::mul_small(t1, t1, 8u); F::add(r.Y, t0, t1);
            F::mul(t0, yy, yz); F::mul_small(r.Z, t0, 8u);
        } else {
            E xx, yy, zz, xy, yz, xz;
            F::sqr(xx, p.X);
            F::sqr(yy, p.Y);
            F::sqr(zz, p.Z);
            F::mul(xy, p.X, p.Y); F::dbl(xy, xy);
            F::mul(yz, p.Y, p.Z); F::dbl(yz, yz);
            F::mul(xz, p.X, p.Z); F::dbl(xz, xz);
            finish_add(r, xx, yy, zz, xy, yz, xz);
        }
    }

    // ---------------------------------------------------------------- byte boundary
    // x||y big-endian -> affine element; false unless both coordinates < p and on the curve
    // (k256/src/arithmetic/affine.rs:241-270, primeorder/src/affine.rs:164-195)
    ECB_DEV static bool load_affine(Aff& a, const u8* xy) {
        u32 t[L];
        load_be<L>(t, xy);
        bool ok = F::from_limbs(a.x, t);
        load_be<L>(t, xy + FB);
        ok = F::from_limbs(a.y, t) && ok;
        return on_curve(a) && ok;
    }
    ECB_DEV static bool on_curve(const Aff& a) {
        E lhs, rhs, t, b;
        F::sqr(lhs, a.y);
        F::sqr(t, a.x);
        F::mul(rhs, t, a.x);
        if constexpr (!C::A_IS_ZERO) {
            F::dbl(t, a.x); F::add(t, t, a.x);
            F::sub(rhs, rhs, t);
        }
        const_b(b);
        F::add(rhs, rhs, b);
        return F::eq(lhs, rhs);
    }
    // X||Y||Z big-endian -> projective; false when a coordinate is >= p (no curve check: projective
    // coordinates are private in the reference and always produced by the group law)
    ECB_DEV static bool load_proj(Proj& p, const u8* xyz) {
        u32 t[L];
        load_be<L>(t, xyz); bool ok = F::from_limbs(p.X, t);
        load_be<L>(t, xyz + FB); ok = F::from_limbs(p.Y, t) && ok;
        load_be<L>(t, xyz + 2 * FB); ok = F::from_limbs(p.Z, t) && ok;
        return ok;
    }
    // scalar bytes -> plain limbs reduced once mod n (Reduce<Uint>::reduce_bytes, k256 scalar.rs:700-713)
    ECB_DEV static void load_scalar(u32* k, const u8* bytes) {
        u32 t[L], u[L], nn[L];
        load_be<L>(t, bytes);
        ECB_UNROLL
        for (int i = 0; i < L; i++) nn[i] = C::n(i);
        u32 bw = sub_n<L>(u, t, nn);
        select_n<L>(k, bw == 0, u, t);
    }
    // SEC1 slot: identity = all zero; compressed 02/03||x; uncompressed 04||x||y
    // (k256 affine.rs:272-284, primeorder affine.rs:340-358)
    ECB_DEV static void encode(u8* out, bool inf, const E& x, const E& y, bool compress) {
        const int n = compress ? 1 + FB : 1 + 2 * FB;
        if (inf) {
            for (int i = 0; i < n; i++) out[i] = 0;
            return;
        }
        u32 t[L];
        F::to_limbs(t, x);
        store_be<L>(out + 1, t);
        F::to_limbs(t, y);
        if (compress) {
            out[0] = (u8)(2u + (t[0] & 1u));
        } else {
            out[0] = 4;
            store_be<L>(out + 1 + FB, t);
        }
    }

    // ---------------------------------------------------------------- window tables (per thread, local memory)
    template <int N> struct Table { Proj e[N]; };   // e[j] = j * P, e[0] = identity

    template <int N> ECB_DEV static void build_table(Table<N>& t, const Proj& p) {
        set_identity(t.e[0]);
        t.e[1] = p;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 2; j < N; j++) {
            // Table elements are passed to the non-inlined point functions by reference.  Round 2 tried the "copy the
            // dynamically indexed element to a named temporary first" form here (the pattern kernels.cuh lincomb_g_q needs
            // for its accumulator operand, nvcc 12.9 / sm_100a): with it k_verify<CurveK256> (the complete-formula verify
            // kernel) returned wrong results on the B200 while every k_mul_var instance stayed correct, so the direct form,
            // which passes every test in every kernel that builds a table, stays.  Neither form has a defect visible in the
            // source or in the host emulation; the root cause is not isolated.  Regression tests: every table index on the
            // secret path (tests/test_gpu_round2.py), agreement of the two verify kernels on all Wycheproof rows
            // (tests/test_gpu_parity.py::test_verify_kernels_agree), the MUL vectors in CT mode.
            if ((j & 1) == 0) dbl(t.e[j], t.e[j >> 1]);
            else add(t.e[j], t.e[j - 1], p);
        }
    }
    // constant-time |idx| lookup: scan every entry, masked move (k256 mul.rs:92-127,
    // primeorder/src/projective.rs:130-137)
    template <int N, bool CT> ECB_DEV static void table_get(Proj& r, const Table<N>& t, u32 idx) {
        if constexpr (CT) {
            r = t.e[0];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (u32 j = 1; j < (u32)N; j++) {
                u32 mask = (u32)0 - (u32)(j == idx);
                cmov(r, t.e[j], mask);
            }
        } else {
            r = t.e[idx];
        }
    }

    // ---------------------------------------------------------------- primeorder scalar mul
    // k*P with a fixed 4-bit window, MSB first, one complete addition per window (also for a zero digit) and 4 doublings
    // between windows - the schedule of primeorder/src/projective.rs:106-150.  The reference scans a 16-entry table with
    // unsigned digits; here the digits are SIGNED (k + 0x88..8, digit = nibble - 8, the carry becomes a 65th digit 0/1):
    // a 9-entry table {0..8}P (7 instead of 15 point operations to build, half the scan) and a masked negation of Y.
    // The digits, the scan and the negation are branch- and address-independent of k; the result point is the same.
    template <bool CT> ECB_DEV static void mul_window4(Proj& r, const Proj& p, const u32* k) {
        Table<9> tab;
        build_table<9>(tab, p);
        u32 kb[L + 1];
        kb[0] = add_cc(k[0], 0x88888888u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) kb[i] = addc_cc(k[i], 0x88888888u);
        kb[L] = addc(0u, 0u);
        Proj acc, e;
        set_identity(acc);
        const int top = 8 * L;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int w = top; w >= 0; w--) {
            if (w != top) { dbl(acc, acc); dbl(acc, acc); dbl(acc, acc); dbl(acc, acc); }
            u32 mag, neg;
            if (w == top) { mag = kb[L]; neg = 0; }
            else {
                const int d = (int)((kb[w >> 3] >> ((w & 7) * 4)) & 15u) - 8;
                neg = (u32)(d >> 31);
                mag = (u32)((d ^ (int)neg) - (int)neg);
            }
            table_get<9, CT>(e, tab, mag);
            cneg(e, neg);
            add(acc, acc, e);
        }
        r = acc;
    }
};

// ----------------------------------------------------------------------------------------------
// secp256k1 specifics: GLV endomorphism + signed radix-16 (k256/src/arithmetic/mul.rs)

struct K256Glv {
    typedef EC<CurveK256> G;
    typedef G::Proj Proj;
    typedef G::E E;
    typedef CurveK256 C;
    typedef C::Fn Fn;

    struct Split {
        u32 a1[5], a2[5];   // |r1| + 0x88..8, |r2| + 0x88..8 (33 nibbles: signed digit i = nibble i - 8, digit 32 = nibble 32)
        u32 neg1, neg2;     // all-ones when r1 (r2) was negated
    };

    // (k * g) >> 384 rounded (WideScalar::mul_shift_vartime, k256/src/arithmetic/scalar/wide64.rs:64-119)
    template <int WHICH> ECB_DEV static void mul_shift_384(u32* q, const u32* k) {
        u32 gg[8], t[16];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) gg[i] = (WHICH == 1) ? C::g1(i) : C::g2(i);
        mul_wide<8>(t, k, gg);
        u32 rnd = t[11] >> 31;
        q[0] = add_cc(t[12], rnd);
        q[1] = addc_cc(t[13], 0u);
        q[2] = addc_cc(t[14], 0u);
        q[3] = addc_cc(t[15], 0u);
        q[4] = addc(0u, 0u);
        q[5] = q[6] = q[7] = 0;
    }

    // decompose_scalar (mul.rs:260-268) + sign fix (mul.rs:350-362) + radix-16 bias
    ECB_DEV static void decompose(Split& s, const u32* k) {
        u32 q1[8], q2[8], c1[8], c2[8], r1[8], r2[8], t[8], nn[8], hn[8];
        Fn::E cm;
        mul_shift_384<1>(q1, k);
        mul_shift_384<2>(q2, k);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) cm.v[i] = C::minus_b1_R(i);
        Fn::mul_plain(c1, q1, cm);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) cm.v[i] = C::minus_b2_R(i);
        Fn::mul_plain(c2, q2, cm);
        Fn::E a, b, o;
        copy_n<8>(a.v, c1); copy_n<8>(b.v, c2);
        Fn::add(o, a, b);
        copy_n<8>(r2, o.v);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) cm.v[i] = C::minus_lambda_R(i);
        Fn::mul_plain(t, r2, cm);
        copy_n<8>(a.v, k); copy_n<8>(b.v, t);
        Fn::add(o, a, b);
        copy_n<8>(r1, o.v);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) { nn[i] = C::n(i); hn[i] = C::half_n(i); }
        bool h1 = !geq_n<8>(hn, r1);   // r1 > n/2  (scalar.rs:519-523)
        bool h2 = !geq_n<8>(hn, r2);
        u32 m1[8], m2[8];
        sub_n<8>(m1, nn, r1);
        sub_n<8>(m2, nn, r2);
        select_n<8>(r1, h1, m1, r1);
        select_n<8>(r2, h2, m2, r2);
        s.neg1 = (u32)0 - (u32)h1;
        s.neg2 = (u32)0 - (u32)h2;
        bias(s.a1, r1);
        bias(s.a2, r2);
    }
    // a (< 2^128) + 0x8888...8 (32 nibbles)
    ECB_DEV static void bias(u32* o, const u32* a) {
        o[0] = add_cc(a[0], 0x88888888u);
        o[1] = addc_cc(a[1], 0x88888888u);
        o[2] = addc_cc(a[2], 0x88888888u);
        o[3] = addc_cc(a[3], 0x88888888u);
        o[4] = addc(0u, 0u);
    }
    // signed digit i of the biased value: returns magnitude 0..8 and a negate mask
    ECB_DEV static void digit(const u32* a, int i, u32& mag, u32& neg) {
        if (i == 32) { mag = a[4]; neg = 0; return; }
        int d = (int)((a[i >> 3] >> ((i & 7) * 4)) & 15u) - 8;
        neg = (u32)(d >> 31);
        mag = (u32)((d ^ (int)neg) - (int)neg);
    }
    ECB_DEV static void endo(Proj& p) {   // (x, y) -> (beta x, y)   projective.rs:287-293
        E beta;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) beta.v[i] = C::beta(i);
        FpK256::mul(p.X, p.X, beta);
    }

    // k*P: one 9-entry table of P, beta applied on the fly for the lambda half, 33 rounds of
    // (4 doublings, 2 additions) — the N = 1 case of lincomb(), mul.rs:342-393.
    template <bool CT> ECB_DEV static void mul(Proj& r, const Proj& p, const u32* k) {
        Split s;
        decompose(s, k);
        G::Table<9> tab;
        G::build_table<9>(tab, p);
        Proj acc, e;
        G::set_identity(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 32; i >= 0; i--) {
            if (i != 32) { G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); }
            u32 mag, neg;
            digit(s.a1, i, mag, neg);
            G::table_get<9, CT>(e, tab, mag);
            G::cneg(e, neg ^ s.neg1);
            G::add(acc, acc, e);
            digit(s.a2, i, mag, neg);
            G::table_get<9, CT>(e, tab, mag);
            endo(e);
            G::cneg(e, neg ^ s.neg2);
            G::add(acc, acc, e);
        }
        r = acc;
    }
};

// curve-dispatching variable-base multiplication
template <class C, bool CT> struct VarMul {
    ECB_DEV static void run(typename EC<C>::Proj& r, const typename EC<C>::Proj& p, const u32* k) {
        EC<C>::template mul_window4<CT>(r, p, k);
    }
};
template <bool CT> struct VarMul<CurveK256, CT> {
    ECB_DEV static void run(EC<CurveK256>::Proj& r, const EC<CurveK256>::Proj& p, const u32* k) {
        K256Glv::mul<CT>(r, p, k);
    }
};

}  // namespace ecb
