// jac.cuh — variable-time fast path for PUBLIC inputs (ECDSA verification, vartime scalar mul).
//
// The reference computes everything with the complete projective formulas and constant-time scans
// (k256/src/arithmetic/mul.rs:342-393, primeorder/src/projective.rs:106-150) because its API serves
// secret scalars too.  Outputs are canonical (SEC1 bytes / booleans), so for public inputs the device
// is free to use cheaper arithmetic (north_star: "variable-time windows are allowed
// only for public-input verification"):
//   * Jacobian coordinates (x = X/Z^2, y = Y/Z^3, Z = 0 = identity): doubling 2M+5S (a = 0) or 3M+5S
//     (a = -3), mixed addition 7M+4S, against 6M+2S+3m / 8M+3S+2m_b and 11M+3m / 13M for RCB;
//   * exceptional cases (P = +-Q, identity operands) are handled by explicit, warp-rarely-taken branches,
//     so results stay exact for adversarial inputs (Wycheproof "edge case" rows);
//   * u1*G comes from a precomputed affine table of all v * 2^(gw*w) * G (gw-bit windows, no doublings);
//   * secp256k1: the per-element table {1..8}Q is brought to ONE common Z without any inversion by
//     working on the isomorphic curve y^2 = x^3 + 7 Z^6 (the a = 0 doubling does not involve b), so all
//     66 GLV additions are mixed additions; the true Z is restored with one multiplication at the end.
//   * primeorder curves: the per-element table {1..8}Q is made AFFINE for the whole batch by a separate kernel (k_wintab:
//     Jacobian multiples, one inversion per thread shared by 16 rows x 7 entries through Montgomery's trick), so the window
//     loop runs on mixed additions too (mul_window_affine); recovery, whose point is derived in-kernel, keeps mul_window_signed.
// Secret-scalar entry points (ECB200_FLAG_CT) never come here.
#pragma once
#include "ec.cuh"

namespace ecb {

template <class C> struct Jac {
    typedef typename C::F F;
    typedef typename F::E E;
    static constexpr int L = C::L;
    struct J { E X, Y, Z; };
    struct A { E x, y; };

    ECB_DEV static void set_inf(J& p) { F::set_one(p.X); F::set_one(p.Y); F::set_zero(p.Z); }
    ECB_DEV static bool is_inf(const J& p) { return F::is_zero(p.Z); }
    ECB_DEV static void from_affine(J& p, const A& a) { p.X = a.x; p.Y = a.y; F::set_one(p.Z); }

    // ---- doubling
    // Both forms are the textbook Jacobian doubling divided through by powers of two (Z3 = Y Z instead of 2 Y Z, so
    // X3 and Y3 carry 1/4 and 1/8): with L = M/2 the formulas need 6 (a = 0) / 8 (a = -3) add-type field operations
    // instead of 10 / 12 - every one of them is 19-26 instructions of carry chain plus canonicalisation.
    ECB_DEV static void dbl_body(J& r, const J& p) {
        if constexpr (C::A_IS_ZERO) {
            // a = 0:  S = Y^2, L = 3 X^2 / 2, T = X S, X3 = L^2 - 2T, Y3 = L (T - X3) - S^2, Z3 = Y Z     (3M + 4S)
            E s_, l, t, u;
            F::mul(u, p.Y, p.Z);                      // Z3 (written last: r may alias p)
            F::sqr(s_, p.Y);
            F::sqr(l, p.X);
            F::half(t, l); F::add(l, l, t);           // L = X^2 + X^2/2
            F::mul(t, p.X, s_);                       // T
            F::sqr(r.X, l);
            F::sub(r.X, r.X, t); F::sub(r.X, r.X, t); // X3 = L^2 - 2T
            F::sqr(s_, s_);                           // S^2
            F::sub(t, t, r.X);
            F::mul(t, l, t);
            F::sub(r.Y, t, s_);
            r.Z = u;
        } else {
            // a = -3: delta = Z^2, S = Y^2, T = X S, m = (X - delta)(X + delta), L = 3m/2,
            //         X3 = L^2 - 2T, Y3 = L (T - X3) - S^2, Z3 = Y Z                                        (4M + 4S)
            E delta, s_, t, t0, t1, l, u;
            F::mul(u, p.Y, p.Z);                      // Z3
            F::sqr(delta, p.Z);
            F::sqr(s_, p.Y);
            F::mul(t, p.X, s_);                       // T
            F::sub(t0, p.X, delta); F::add(t1, p.X, delta);
            F::mul(l, t0, t1);                        // m
            F::half(t0, l); F::add(l, l, t0);         // L = m + m/2
            F::sqr(r.X, l);
            F::sub(r.X, r.X, t); F::sub(r.X, r.X, t); // X3
            F::sqr(s_, s_);
            F::sub(t, t, r.X);
            F::mul(t, l, t);
            F::sub(r.Y, t, s_);
            r.Z = u;
        }
    }
    ECB_POINT_FN static void dbl(J& r, const J& p) { dbl_body(r, p); }
    // n doublings in one call: the point is read from and written back to local memory once instead of n times
    ECB_POINT_FN static void dbl_n(J& r, int n) {
        J t = r;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < n; i++) dbl_body(t, t);
        r = t;
    }
    // doubling of an affine point (Z = 1): used for the exceptional branch and table starts
    ECB_DEV static void dbl_affine(J& r, const A& q) {
        J t;
        from_affine(t, q);
        dbl(r, t);
    }

#ifndef ECB_MADD_EARLY_STORE
#define ECB_MADD_EARLY_STORE 1
#endif
    // ---- mixed addition r = p + q, q affine and not the identity.  zr (optional) receives Z3 / Z1 (= H)
    // (valid only on the generic branch; callers that need it guarantee no exceptional case).
    ECB_POINT_FN static void madd(J& r, const J& p, const A& q, E* zr) {
        if (is_inf(p)) { from_affine(r, q); return; }
        // madd-2004-hmv (8M + 3S, 7 add-type ops): Z1Z1 = Z1^2, U2 = x2*Z1Z1, S2 = y2*Z1*Z1Z1, H = U2 - X1, R = S2 - Y1,
        // Z3 = Z1*H, HH = H^2, HHH = H*HH, V = X1*HH, X3 = R^2 - HHH - 2V, Y3 = R(V - X3) - Y1*HHH
#if ECB_MADD_EARLY_STORE
        // Same formulas, ordered for register pressure: every value dies as early as the data flow allows and each output
        // coordinate is stored as soon as its input counterpart has been read for the last time (r may alias p), so at most
        // four field elements are live across a multiplier call instead of five
        E z1z1, u2, s2, h, rr, hh, hhh, v, t;
        F::sqr(z1z1, p.Z);
        F::mul(u2, q.x, z1z1);
        F::sub(h, u2, p.X);
        F::mul(s2, q.y, p.Z); F::mul(s2, s2, z1z1);
        F::sub(rr, s2, p.Y);
        if (F::is_zero(h)) {
            if (F::is_zero(rr)) dbl_affine(r, q);   // p == q
            else set_inf(r);                         // p == -q
            return;
        }
        if (zr) *zr = h;
        F::mul(t, p.Z, h);
        r.Z = t;                                     // Z1 is not read again
        F::sqr(hh, h);
        F::mul(hhh, hh, h);
        F::mul(v, p.X, hh);
        F::sqr(t, rr); F::sub(t, t, hhh); F::sub(t, t, v); F::sub(t, t, v);   // two subtractions are cheaper than dbl + sub
        r.X = t;                                     // X1 is not read again
        F::sub(v, v, t); F::mul(v, rr, v);
        F::mul(t, p.Y, hhh);
        F::sub(v, v, t);
        r.Y = v;
#else
        E z1z1, u2, s2, h, rr, hh, hhh, v, t;
        F::sqr(z1z1, p.Z);
        F::mul(u2, q.x, z1z1);
        F::mul(s2, q.y, p.Z); F::mul(s2, s2, z1z1);
        F::sub(h, u2, p.X);
        F::sub(rr, s2, p.Y);
        if (F::is_zero(h)) {
            if (F::is_zero(rr)) dbl_affine(r, q);   // p == q
            else set_inf(r);                         // p == -q
            return;
        }
        if (zr) *zr = h;
        F::sqr(hh, h);
        F::mul(hhh, hh, h);
        F::mul(v, p.X, hh);
        E x3, y3, z3;
        F::mul(z3, p.Z, h);
        F::sqr(x3, rr); F::sub(x3, x3, hhh); F::sub(x3, x3, v); F::sub(x3, x3, v);   // two subtractions are cheaper than dbl + sub
        F::sub(y3, v, x3); F::mul(y3, rr, y3);
        F::mul(t, p.Y, hhh);
        F::sub(y3, y3, t);
        r.X = x3; r.Y = y3; r.Z = z3;
#endif
    }

    // ---- branch-free mixed addition for the secret-scalar FIXED-BASE path (kernels.cuh body_gen_half): p += q in place.
    // `inf` (all-ones / zero) says p is still the identity, `take` (all-ones / zero) says the digit is non-zero; both are
    // secret-derived and only ever used as masks.  The generic madd-2004-hmv result is always computed; it is replaced by
    // q when p was the identity and dropped when the digit is zero.  The caller guarantees p != +-q (h != 0) whenever
    // both are real points: inside one half of the split fixed-base schedule the running sum is smaller in absolute value
    // than every entry of the window being added (see body_gen_half), so the exceptional cases of the Jacobian formulas
    // cannot occur and no complete formula is needed until the halves are combined.  Returns the new `inf`.
    ECB_POINT_FN static u32 madd_ct(J& p, u32 inf, const A& q, u32 take) {
#if ECB_MADD_EARLY_STORE
        // ordered like madd above: each coordinate is committed (by mask) as soon as its old value has been read for the last time
        E z1z1, u2, s2, h, rr, hh, hhh, v, t;
        F::sqr(z1z1, p.Z);
        F::mul(u2, q.x, z1z1);
        F::sub(h, u2, p.X);
        F::mul(s2, q.y, p.Z); F::mul(s2, s2, z1z1);
        F::sub(rr, s2, p.Y);
        F::mul(t, p.Z, h);
        F::set_one(u2);
        F::cmov(t, u2, inf);
        F::cmov(p.Z, t, take);
        F::sqr(hh, h);
        F::mul(hhh, hh, h);
        F::mul(v, p.X, hh);
        F::sqr(t, rr); F::sub(t, t, hhh); F::sub(t, t, v); F::sub(t, t, v);
        F::sub(v, v, t); F::mul(v, rr, v);                 // rr (V - X3), with the true X3
        F::cmov(t, q.x, inf);
        F::cmov(p.X, t, take);
        F::mul(t, p.Y, hhh);
        F::sub(v, v, t);
        F::cmov(v, q.y, inf);
        F::cmov(p.Y, v, take);
        return inf & ~take;
#else
        E z1z1, u2, s2, h, rr, hh, hhh, v, t, x3, y3, z3, one;
        F::sqr(z1z1, p.Z);
        F::mul(u2, q.x, z1z1);
        F::mul(s2, q.y, p.Z); F::mul(s2, s2, z1z1);
        F::sub(h, u2, p.X);
        F::sub(rr, s2, p.Y);
        F::sqr(hh, h);
        F::mul(hhh, hh, h);
        F::mul(v, p.X, hh);
        F::mul(z3, p.Z, h);
        F::sqr(x3, rr); F::sub(x3, x3, hhh); F::sub(x3, x3, v); F::sub(x3, x3, v);
        F::sub(y3, v, x3); F::mul(y3, rr, y3);
        F::mul(t, p.Y, hhh);
        F::sub(y3, y3, t);
        F::set_one(one);
        F::cmov(x3, q.x, inf); F::cmov(y3, q.y, inf); F::cmov(z3, one, inf);
        F::cmov(p.X, x3, take); F::cmov(p.Y, y3, take); F::cmov(p.Z, z3, take);
        return inf & ~take;
#endif
    }

    // ---- full Jacobian addition r = p + q
    ECB_POINT_FN static void add(J& r, const J& p, const J& q) {
        if (is_inf(q)) { r = p; return; }
        if (is_inf(p)) { r = q; return; }
        // add-2007-bl (11M + 5S, 12 add-type ops).  add-1998-cmo-2 (12M + 4S, 7 add-type ops) was measured too: +0.6 % on
        // P-256 verify but -9 % on P-384 P*k, where a multiplication costs 67 IMAD.WIDE more than a squaring.
        E z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
        F::sqr(z1z1, p.Z);
        F::sqr(z2z2, q.Z);
        F::mul(u1, p.X, z2z2);
        F::mul(u2, q.X, z1z1);
        F::mul(s1, p.Y, q.Z); F::mul(s1, s1, z2z2);
        F::mul(s2, q.Y, p.Z); F::mul(s2, s2, z1z1);
        F::sub(h, u2, u1);
        F::sub(rr, s2, s1);
        if (F::is_zero(h)) {
            if (F::is_zero(rr)) dbl(r, p);
            else set_inf(r);
            return;
        }
        F::dbl(rr, rr);
        F::dbl(i, h); F::sqr(i, i);
        F::mul(j, h, i);
        F::mul(v, u1, i);
        F::add(t, p.Z, q.Z); F::sqr(t, t); F::sub(t, t, z1z1); F::sub(t, t, z2z2); F::mul(t, t, h);
        E x3, y3;
        F::sqr(x3, rr); F::sub(x3, x3, j); F::sub(x3, x3, v); F::sub(x3, x3, v);
        F::sub(y3, v, x3); F::mul(y3, rr, y3);
        F::mul(j, s1, j); F::dbl(j, j);
        F::sub(y3, y3, j);
        r.X = x3; r.Y = y3; r.Z = t;
    }

    // Jacobian -> homogeneous projective (x = X'/Z', y = Y'/Z') for the shared normalisation kernel:
    // (X*Z : Y : Z^3); the identity maps to (0 : 1 : 0).
    ECB_DEV static void to_proj(typename EC<C>::Proj& o, const J& p) {
        if (is_inf(p)) { EC<C>::set_identity(o); return; }
        E zz;
        F::sqr(zz, p.Z);
        F::mul(o.X, p.X, p.Z);
        o.Y = p.Y;
        F::mul(o.Z, zz, p.Z);
    }

    // signed radix-16 digit i of a value biased by 0x88..8 (see K256Glv::bias): magnitude 0..8, negate mask
    ECB_DEV static void digit16(const u32* a, int i, int top, u32& mag, u32& neg) {
        if (i == top) { mag = a[top >> 3]; neg = 0; return; }
        int d = (int)((a[i >> 3] >> ((i & 7) * 4)) & 15u) - 8;
        neg = (u32)(d >> 31);
        mag = (u32)((d ^ (int)neg) - (int)neg);
    }
    ECB_DEV static void cneg_y(A& a, u32 mask) {
        E ny;
        F::neg(ny, a.y);
        F::cmov(a.y, ny, mask);
    }

    // ---- u1 * G from the big fixed-base table: nwin = ceil(32L / gw) windows of gw bits (any 2 <= gw <= 24; windows may
    // straddle words), the top one holds the tb = 32L - gw (nwin - 1) bits that are left.  All windows but the top one are
    // SIGNED: with the bias 2^(gw-1) added to each of them, window w holds d_w + 2^(gw-1), d_w in [-2^(gw-1), 2^(gw-1));
    // the top window absorbs the carry and stays unsigned, v in [0, 2^tb] (so no extra addition for a carry window).
    //   tab[(w << (gw-1)) + v - 1]        = v * 2^(gw*w) * G, 1 <= v <= 2^(gw-1), w < nwin-1   (negative digits negate y)
    //   tab[((nwin-1) << (gw-1)) + v - 1] = v * 2^(gw*(nwin-1)) * G, 1 <= v <= 2^tb
    // 34 MiB instead of the 64 MiB of an all-unsigned table for a 256-bit curve at gw = 16 (16 additions per row); the table is
    // built once per context and curve, so wider windows trade HBM for additions: gw = 20 -> 13 additions, 410 MB; 22 -> 12, 1.5 GB.
    ECB_DEV static void add_fixed_base(J& acc, const u32* u1, const u32* tab, int gw) {
        const int nwin = (32 * L + gw - 1) / gw;
        const int top_bit = gw * (nwin - 1);
        u32 kb[L + 2];
        ECB_UNROLL
        for (int i = 0; i < L + 2; i++) kb[i] = 0;
        for (int b = gw - 1; b < top_bit; b += gw) kb[b >> 5] |= 1u << (b & 31);   // the bias, one bit per signed window
        kb[0] = add_cc(u1[0], kb[0]);
        ECB_UNROLL
        for (int i = 1; i < L; i++) kb[i] = addc_cc(u1[i], kb[i]);
        kb[L] = addc(0u, 0u);
        const u32 vmask = (1u << gw) - 1u;
        const u32 half = 1u << (gw - 1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int w = 0; w < nwin; w++) {
            const int bit = w * gw;
            const u64 two = ((u64)kb[(bit >> 5) + 1] << 32) | kb[bit >> 5];
            const u32 raw = (u32)(two >> (bit & 31));
            u32 mag, neg = 0;
            if (w == nwin - 1) {
                mag = raw;                                       // unsigned: everything from top_bit up, carry included (<= 2^tb)
            } else {
                const int d = (int)(raw & vmask) - (int)half;
                neg = (u32)(d >> 31);
                mag = (u32)((d ^ (int)neg) - (int)neg);
            }
            if (mag) {
                A g;
                load_entry(g, tab + (((size_t)w << (gw - 1)) + mag - 1) * 2 * L);   // 128-bit gathers: every lane reads its own line
                cneg_y(g, neg);
                madd(acc, acc, g, nullptr);
            }
        }
    }

    // ---- generic (a = -3) variable-base part: acc = k * Q, signed radix-16, Jacobian table {1..8}Q
    ECB_DEV static void mul_window_signed(J& acc, const A& Q, const u32* k) {
        J tab[8];                                  // tab[j-1] = j*Q
        from_affine(tab[0], Q);
        dbl(tab[1], tab[0]);
        madd(tab[2], tab[1], Q, nullptr);
        dbl(tab[3], tab[1]);
        madd(tab[4], tab[3], Q, nullptr);
        dbl(tab[5], tab[2]);
        madd(tab[6], tab[5], Q, nullptr);
        dbl(tab[7], tab[3]);
        u32 kb[L + 1];
        kb[0] = add_cc(k[0], 0x88888888u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) kb[i] = addc_cc(k[i], 0x88888888u);
        kb[L] = addc(0u, 0u);
        set_inf(acc);
        const int top = 8 * L;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = top; i >= 0; i--) {
            if (i != top) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                dbl_n(acc, 4);
            }
            u32 mag, neg;
            digit16(kb, i, top, mag, neg);
            if (mag) {
                J e = tab[mag - 1];
                E ny;
                F::neg(ny, e.Y);
                F::cmov(e.Y, ny, neg);
                add(acc, acc, e);
            }
        }
    }

    // ---- the same loop over an AFFINE table {1..8}Q that lives in global memory (2L limbs per entry, built for the whole
    // batch by k_wintab: Jacobian multiples, then ONE field inversion per thread shared by up to 16 rows x 7 entries through
    // Montgomery's trick).  Every window addition becomes a mixed addition: 8M+3S instead of 11M+5S, 60 times per row.
    // Measured and not kept: prefetch.global.L1 / .L2 of the entry before the four doublings (+-0.1 %: six resident CTAs
    // already hide the gather), the Jacobian loop compiled out of the kernel (see kernels.cuh body_verify_main).
    ECB_DEV static void load_entry(A& e, const u32* p) {
#if defined(__CUDA_ARCH__)
        if constexpr (L % 4 == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
            ECB_UNROLL
            for (int k = 0; k < L / 4; k++) {
                uint4 a = __ldg(q + k), b = __ldg(q + L / 4 + k);
                e.x.v[4 * k] = a.x; e.x.v[4 * k + 1] = a.y; e.x.v[4 * k + 2] = a.z; e.x.v[4 * k + 3] = a.w;
                e.y.v[4 * k] = b.x; e.y.v[4 * k + 1] = b.y; e.y.v[4 * k + 2] = b.z; e.y.v[4 * k + 3] = b.w;
            }
        } else if constexpr (L % 2 != 0) {   // L = 7: 28-byte coordinates, word loads
            ECB_UNROLL
            for (int l = 0; l < L; l++) { e.x.v[l] = __ldg(p + l); e.y.v[l] = __ldg(p + L + l); }
        } else {   // L = 6: 24-byte coordinates, 8-byte aligned
            const uint2* q = reinterpret_cast<const uint2*>(p);
            ECB_UNROLL
            for (int k = 0; k < L / 2; k++) {
                uint2 a = __ldg(q + k), b = __ldg(q + L / 2 + k);
                e.x.v[2 * k] = a.x; e.x.v[2 * k + 1] = a.y;
                e.y.v[2 * k] = b.x; e.y.v[2 * k + 1] = b.y;
            }
        }
#else
        ECB_UNROLL
        for (int l = 0; l < L; l++) { e.x.v[l] = p[l]; e.y.v[l] = p[L + l]; }
#endif
    }
    ECB_DEV static void mul_window_affine(J& acc, const u32* tab, const u32* k) {
        u32 kb[L + 1];
        kb[0] = add_cc(k[0], 0x88888888u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) kb[i] = addc_cc(k[i], 0x88888888u);
        kb[L] = addc(0u, 0u);
        set_inf(acc);
        const int top = 8 * L;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = top; i >= 0; i--) {
            u32 mag, neg;
            digit16(kb, i, top, mag, neg);
            const u32* ent = tab + (size_t)(mag ? mag - 1 : 0) * 2 * L;
            if (i != top) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                dbl_n(acc, 4);
            }
            if (mag) {
                A e;
                load_entry(e, ent);
                cneg_y(e, neg);
                madd(acc, acc, e, nullptr);   // handles acc = identity and acc = +-e
            }
        }
    }
};

// ----------------------------------------------------------------------------------------------
// secp256k1: GLV + shared-Z table on the isomorphic curve
struct K256Fast {
    typedef CurveK256 C;
    typedef Jac<C> JJ;
    typedef JJ::J J;
    typedef JJ::A A;
    typedef FpK256 F;
    typedef F::E E;

    // Per-thread window table {1..8}Q.  The first NS entries live in shared memory (word w of entry j at
    // s[(16 j + w) * stride], s already offset by the thread index: conflict-free whatever entry each thread picks), the
    // rest in local memory.  NS = 7 (56 KB per 128-thread CTA, four CTAs per SM still fit) shrinks the local frame by 448
    // bytes per thread and halves the DRAM write-back traffic, but measured 4 % slower on the B200 (kernels_impl.cuh), so the
    // shipped kernels use NS = 0 (all local), which is also what the host emulation runs.
    template <int NS> struct WinTab {
        u32* s;
        int stride;
        A loc[8 - NS];
        ECB_DEV void store(int j, const E& x, const E& y) {
            if (j < NS) {
                ECB_UNROLL
                for (int w = 0; w < 8; w++) { s[(16 * j + w) * stride] = x.v[w]; s[(16 * j + 8 + w) * stride] = y.v[w]; }
            } else { loc[j - NS].x = x; loc[j - NS].y = y; }
        }
        ECB_DEV void load(A& a, int j) const {
            if (j < NS) {
                ECB_UNROLL
                for (int w = 0; w < 8; w++) { a.x.v[w] = s[(16 * j + w) * stride]; a.y.v[w] = s[(16 * j + 8 + w) * stride]; }
            } else { a = loc[j - NS]; }
        }
    };

    // acc = (r1 + r2*lambda) * Q with the split s (|r1|, |r2| < 2^128, signs in s.neg*): 128 doublings,
    // <= 66 mixed additions.  Q must be a valid affine point (not the identity).
    template <int NS> ECB_DEV static void mul_glv(J& acc, const A& Q, const K256Glv::Split& s, u32* stab, int sstride) {
        // 1) multiples 1..8 of Q in Jacobian form; only (X_j, Y_j) and zr_j = Z_j / Z_{j-1} are kept
        WinTab<NS> tab;
        tab.s = stab; tab.stride = sstride;
        E zr[7];
        J& cur = acc;                               // the accumulator's storage doubles as the running multiple (smaller frame)
        tab.store(0, Q.x, Q.y);
        JJ::from_affine(cur, Q);
        JJ::dbl(cur, cur);                          // Z2 = 2 y1  (Z1 = 1)
        tab.store(1, cur.X, cur.Y); zr[0] = cur.Z;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 3; j <= 8; j++) {   // (j-1)Q + Q never hits an exceptional case (the order of Q is a large prime)
            E z;
            JJ::madd(cur, cur, Q, &z);   // in place: madd reads p completely before it writes r
            tab.store(j - 1, cur.X, cur.Y);
            zr[j - 2] = z;
        }
        const E z8 = cur.Z;
        // 2) rescale entries 1..7 to Z8: entry j gets (x * zs^2, y * zs^3) with zs = Z8 / Zj = prod_{i>j} zr_i
        E zs = zr[6];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 7; j >= 1; j--) {
            A e;
            tab.load(e, j - 1);
            E zs2, zs3;
            F::sqr(zs2, zs);
            F::mul(zs3, zs2, zs);
            F::mul(e.x, e.x, zs2);
            F::mul(e.y, e.y, zs3);
            tab.store(j - 1, e.x, e.y);
            if (j > 1) { E f = zr[j - 2]; F::mul(zs, zs, f); }
        }
        // tab[0..7] = {1..8}Q are now affine points of the isomorphic curve y^2 = x^3 + 7*Z8^6 (global Z = Z8)
        E beta;
        ECB_UNROLL
        for (int l = 0; l < 8; l++) beta.v[l] = C::beta(l);
        JJ::set_inf(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 32; i >= 0; i--) {
            if (i != 32) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                JJ::dbl_n(acc, 4);
            }
            u32 mag, neg;
            K256Glv::digit(s.a1, i, mag, neg);
            if (mag) {
                A e;
                tab.load(e, (int)mag - 1);
                JJ::cneg_y(e, neg ^ s.neg1);
                JJ::madd(acc, acc, e, nullptr);
            }
            K256Glv::digit(s.a2, i, mag, neg);
            if (mag) {
                A e;
                tab.load(e, (int)mag - 1);
                F::mul(e.x, e.x, beta);
                JJ::cneg_y(e, neg ^ s.neg2);
                JJ::madd(acc, acc, e, nullptr);
            }
        }
        // 3) back to the real curve: Z *= Z8
        if (!JJ::is_inf(acc)) F::mul(acc.Z, acc.Z, z8);
    }
};

}  // namespace ecb
