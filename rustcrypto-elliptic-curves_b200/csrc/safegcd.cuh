// safegcd.cuh — constant-time modular inversion by Bernstein-Yang divsteps ("safegcd", eprint 2019/266), on signed 30-bit
// limbs, for the ONE inversion chain a CTA runs in the Montgomery-trick kernels (kernels_impl.cuh BlockInv).
//
// Why: that chain is latency, not throughput - one thread inverts the product of the CTA's rows while the other 127 wait at a
// barrier.  Fermat's a^(p-2) is ~256 squarings + 15-78 multiplications, each a ~150-instruction web of dependent carry chains
// (~700 cycles for a lone warp on the B200: ~95 us per inversion; ncu shows `barrier` as the top stall of k_verify_prep,
// k_kt_fill, k_wintab and k_normalize).  divsteps needs 20 rounds of (30 cheap word-level steps + two 2x2-matrix updates of
// 9-limb integers) for a 256-bit modulus: about a quarter of the dependent path.  The reference itself inverts its p384 /
// sm2 field elements this way (primeorder/src/field.rs:506-559 wraps crypto-bigint's Bernstein-Yang inverter) and uses Fermat
// addition chains for k256 / p256 (k256/src/arithmetic/field.rs:187-218, p256/src/arithmetic/field.rs:364-382); both give
// THE inverse, a canonical residue, so outputs are bit-identical whichever runs.
//
// Constant time: a fixed number of divsteps (the "half-delta" variant needs at most 590 for 256-bit inputs - Wuille's
// convex-hull bound, as used by libsecp256k1's modinv32 - so 20 rounds of 30; other sizes take the original paper's bound
// floor((49 d + 57) / 17) for the slower plain variant, which is conservative for this one), masks instead of branches,
// no data-dependent address.  Extra divsteps after g has reached 0 leave (f, d) unchanged, so over-estimating is harmless.
// 0 maps to 0, as with Fermat.
#pragma once
#include "bigint.cuh"

#ifndef ECB_SAFEGCD
#define ECB_SAFEGCD 1      // the Montgomery-trick kernels invert with divsteps; 0 = Fermat chains (A/B measurements)
#endif

namespace ecb {

template <int L> struct SafeGcd {
    static constexpr int BITS = 32 * L;
    static constexpr int NL = (BITS + 2 + 29) / 30;                 // 30-bit limbs incl. head-room for values in (-2p, p)
    static constexpr int ROUNDS = BITS == 256 ? 20 : ((49 * BITS + 57) / 17 + 29) / 30;
    static constexpr int M30 = 0x3FFFFFFF;
    typedef long long i64;
    struct S30 { int v[NL]; };
    struct T2 { int u, v, q, r; };

    ECB_DEV static void to_s30(S30& r, const u32* a) {
        ECB_UNROLL
        for (int i = 0; i < NL; i++) {
            const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
            u64 two = a[w];
            if (w + 1 < L) two |= (u64)a[w + 1] << 32;
            r.v[i] = (int)((u32)(two >> sh) & (u32)M30);
        }
    }
    ECB_DEV static void from_s30(u32* a, const S30& s) {               // s in [0, 2^BITS), limbs in [0, 2^30)
        u64 acc = 0;
        int nb = 0, j = 0;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) {
            acc |= (u64)(u32)s.v[i] << nb;
            nb += 30;
            if (nb >= 32 && j < L) { a[j++] = (u32)acc; acc >>= 32; nb -= 32; }
        }
        if (j < L) a[j] = (u32)acc;
    }
    // 30 divsteps on the low words; the transition matrix t (scaled by 2^30) maps (f, g) -> (f', g') = t (f, g) / 2^30
    ECB_DEV static int divsteps_30(int zeta, u32 f0, u32 g0, T2& t) {
        u32 u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#if defined(__CUDA_ARCH__)
#pragma unroll 6
#endif
        for (int i = 0; i < 30; i++) {
            u32 c1 = (u32)(zeta >> 31);
            const u32 c2 = (u32)0 - (g & 1u);
            const u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
            g += x & c2; q += y & c2; r += z & c2;
            c1 &= c2;
            zeta = (int)((u32)zeta ^ c1) - 1;
            f += g & c1; u += q & c1; v += r & c1;
            g >>= 1; u <<= 1; v <<= 1;
        }
        t.u = (int)u; t.v = (int)v; t.q = (int)q; t.r = (int)r;
        return zeta;
    }
    // (d, e) <- t (d, e) / 2^30 mod p, results in (-2p, p)
    ECB_DEV static void update_de(S30& d, S30& e, const T2& t, const S30& mod, u32 mod_inv30) {
        const int u = t.u, v = t.v, q = t.q, r = t.r;
        const int sd = d.v[NL - 1] >> 31, se = e.v[NL - 1] >> 31;
        int md = (u & sd) + (v & se), me = (q & sd) + (r & se);
        int di = d.v[0], ei = e.v[0];
        i64 cd = (i64)u * di + (i64)v * ei, ce = (i64)q * di + (i64)r * ei;
        md -= (int)((mod_inv30 * (u32)cd + (u32)md) & (u32)M30);
        me -= (int)((mod_inv30 * (u32)ce + (u32)me) & (u32)M30);
        cd += (i64)mod.v[0] * md; ce += (i64)mod.v[0] * me;
        cd >>= 30; ce >>= 30;
        ECB_UNROLL
        for (int i = 1; i < NL; i++) {
            di = d.v[i]; ei = e.v[i];
            cd += (i64)u * di + (i64)v * ei; ce += (i64)q * di + (i64)r * ei;
            cd += (i64)mod.v[i] * md; ce += (i64)mod.v[i] * me;
            d.v[i - 1] = (int)cd & M30; cd >>= 30;
            e.v[i - 1] = (int)ce & M30; ce >>= 30;
        }
        d.v[NL - 1] = (int)cd; e.v[NL - 1] = (int)ce;
    }
    // (f, g) <- t (f, g) / 2^30 (exact)
    ECB_DEV static void update_fg(S30& f, S30& g, const T2& t) {
        const int u = t.u, v = t.v, q = t.q, r = t.r;
        int fi = f.v[0], gi = g.v[0];
        i64 cf = (i64)u * fi + (i64)v * gi, cg = (i64)q * fi + (i64)r * gi;
        cf >>= 30; cg >>= 30;
        ECB_UNROLL
        for (int i = 1; i < NL; i++) {
            fi = f.v[i]; gi = g.v[i];
            cf += (i64)u * fi + (i64)v * gi; cg += (i64)q * fi + (i64)r * gi;
            f.v[i - 1] = (int)cf & M30; cf >>= 30;
            g.v[i - 1] = (int)cg & M30; cg >>= 30;
        }
        f.v[NL - 1] = (int)cf; g.v[NL - 1] = (int)cg;
    }
    ECB_DEV static void carry(S30& r) {
        ECB_UNROLL
        for (int i = 1; i < NL; i++) { r.v[i] += r.v[i - 1] >> 30; r.v[i - 1] &= M30; }
    }
    // r in (-2p, p), negated when sign < 0, brought to [0, p)
    ECB_DEV static void normalize(S30& r, int sign, const S30& mod) {
        int cond_add = r.v[NL - 1] >> 31;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) r.v[i] += mod.v[i] & cond_add;
        const int cond_neg = sign >> 31;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) r.v[i] = (r.v[i] ^ cond_neg) - cond_neg;
        carry(r);
        cond_add = r.v[NL - 1] >> 31;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) r.v[i] += mod.v[i] & cond_add;
        carry(r);
    }
    // r = x^-1 mod p for plain little-endian limbs, x < p, p odd; 0 -> 0
    ECB_DEV static void inv(u32* r, const u32* x, const u32* p) {
        S30 d, e, f, g, mod;
        to_s30(mod, p);
        to_s30(g, x);
        f = mod;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) { d.v[i] = 0; e.v[i] = 0; }
        e.v[0] = 1;
        u32 pi = p[0];                                                 // p^-1 mod 2^32 by Newton: 3 -> 6 -> 12 -> 24 -> 48 bits
        pi *= 2u - p[0] * pi; pi *= 2u - p[0] * pi; pi *= 2u - p[0] * pi; pi *= 2u - p[0] * pi;
        const u32 mod_inv30 = pi & (u32)M30;
        int zeta = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < ROUNDS; i++) {
            T2 t;
            zeta = divsteps_30(zeta, (u32)f.v[0], (u32)g.v[0], t);
            update_de(d, e, t, mod, mod_inv30);
            update_fg(f, g, t);
        }
        normalize(d, f.v[NL - 1], mod);                                // g = 0, f = +-1: x^-1 = sign(f) d
        from_s30(r, d);
    }
};

}  // namespace ecb
