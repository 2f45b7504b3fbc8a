// ecport.cpp — C++ restatement ("port") of the reference's CPU algorithms for the hot path, with the
// reference's own limb layouts, driven by OpenMP over the host cores.
//
// TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline and `--impl reference` legs
// and by tests as a second checker at sizes the Python oracle cannot reach.  The product
// (rustcrypto-elliptic-curves_b200/) never links or loads it.  Parity status: pinned —
// tests/test_port.py checks it against oracle/ecoracle.py (itself pinned to the reference's golden
// vectors) on the reference's MUL vectors, the Wycheproof rows and random inputs.
//
// The reference cannot be compiled here (no Rust toolchain), so this is a port, not the reference:
//   k256 field       5x52 limbs, u128 accumulators, lazy reduction with magnitudes
//                    (k256/src/arithmetic/field/field_5x52.rs:133-285 helpers, :288-449 mul_inner)
//   k256 group       RCB complete add / double with the reference's normalize_weak / negate schedule
//                    (k256/src/arithmetic/projective.rs:96-161, :225-274)
//   k256 scalar mul  GLV decomposition, signed radix-16, interleaved lincomb, 33x8 fixed-base tables
//                    (k256/src/arithmetic/mul.rs:260-304, :342-393, :397-439)
//   primeorder       4x64 / 6x64 word-by-word Montgomery (p256 field.rs:240-319 is a hand-tuned
//                    equivalent; p384/sm2 are fiat-crypto CIOS), RCB a=-3 Alg 4/6
//                    (primeorder/src/point_arithmetic.rs:209-238, :286-317), 4-bit window mul
//                    (primeorder/src/projective.rs:106-150), lincomb = x*k + y*l (:415-420)
//   ecdsa            hazmat::verify_prehashed semantics, k256 low-s rule
//                    (k256/src/ecdsa.rs:201-208)
// Deviations (cost-neutral or stated): scalar-field arithmetic uses a generic Montgomery multiplier
// for every curve; s^-1 in verification is the variable-time Stein inversion the reference calls
// (Scalar::invert_vartime, k256 scalar.rs:467-516, p256 scalar.rs:365-409); to_affine uses the Fermat
// chain as the reference does.
//
//   g++ -O3 -march=x86-64-v3 -fopenmp -shared -fPIC -o oracle/libecport.so oracle/ecport.cpp      (oracle/Makefile)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint8_t u8;
typedef unsigned __int128 u128;

// =============================================================================================
// generic 64-bit Montgomery field (CIOS), runtime modulus
template <int N> struct MontCtx {
    u64 p[N], one[N], r2[N], n0;
    void init(const u64* mod) {
        for (int i = 0; i < N; i++) p[i] = mod[i];
        u64 inv = 1;
        for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;   // p^-1 mod 2^64
        n0 = (u64)0 - inv;
        // one = 2^(64N) mod p by 64N modular doublings of 1; r2 by another 64N
        u64 t[N] = {1};
        for (int i = 0; i < 128 * N; i++) {
            u64 c = 0;
            for (int j = 0; j < N; j++) { u64 nc = t[j] >> 63; t[j] = (t[j] << 1) | c; c = nc; }
            u64 d[N], bw = 0;
            for (int j = 0; j < N; j++) { u128 s = (u128)t[j] - p[j] - bw; d[j] = (u64)s; bw = (u64)(s >> 64) & 1; }
            if (c || !bw) for (int j = 0; j < N; j++) t[j] = d[j];
            if (i == 64 * N - 1) for (int j = 0; j < N; j++) one[j] = t[j];
        }
        for (int j = 0; j < N; j++) r2[j] = t[j];
    }
};

template <int N> static inline void mont_mul(u64* r, const u64* a, const u64* b, const MontCtx<N>& C) {
    u64 t[N + 2] = {0};
    for (int i = 0; i < N; i++) {
        u128 c = 0;
        for (int j = 0; j < N; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[N]; t[N] = (u64)c; t[N + 1] = (u64)(c >> 64);
        u64 m = t[0] * C.n0;
        c = (u128)m * C.p[0] + t[0]; c >>= 64;
        for (int j = 1; j < N; j++) { c += (u128)m * C.p[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[N]; t[N - 1] = (u64)c; t[N] = t[N + 1] + (u64)(c >> 64);
    }
    u64 d[N], bw = 0;
    for (int j = 0; j < N; j++) { u128 s = (u128)t[j] - C.p[j] - bw; d[j] = (u64)s; bw = (u64)(s >> 64) & 1; }
    bool take = t[N] || !bw;
    for (int j = 0; j < N; j++) r[j] = take ? d[j] : t[j];
}
template <int N> static inline void mod_add(u64* r, const u64* a, const u64* b, const u64* p) {
    u64 t[N], c = 0;
    for (int j = 0; j < N; j++) { u128 s = (u128)a[j] + b[j] + c; t[j] = (u64)s; c = (u64)(s >> 64); }
    u64 d[N], bw = 0;
    for (int j = 0; j < N; j++) { u128 s = (u128)t[j] - p[j] - bw; d[j] = (u64)s; bw = (u64)(s >> 64) & 1; }
    bool take = c || !bw;
    for (int j = 0; j < N; j++) r[j] = take ? d[j] : t[j];
}
template <int N> static inline void mod_sub(u64* r, const u64* a, const u64* b, const u64* p) {
    u64 t[N], bw = 0;
    for (int j = 0; j < N; j++) { u128 s = (u128)a[j] - b[j] - bw; t[j] = (u64)s; bw = (u64)(s >> 64) & 1; }
    u64 c = 0;
    for (int j = 0; j < N; j++) { u128 s = (u128)t[j] + (bw ? p[j] : 0) + c; r[j] = (u64)s; c = (u64)(s >> 64); }
}
template <int N> static inline bool geq(const u64* a, const u64* b) {
    for (int j = N - 1; j >= 0; j--) { if (a[j] != b[j]) return a[j] > b[j]; }
    return true;
}
template <int N> static inline bool is_zero(const u64* a) { u64 t = 0; for (int j = 0; j < N; j++) t |= a[j]; return t == 0; }
template <int N> static inline void load_be(u64* v, const u8* b) {
    for (int i = 0; i < N; i++) { u64 w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[8 * (N - 1 - i) + k]; v[i] = w; }
}
template <int N> static inline void store_be(u8* b, const u64* v) {
    for (int i = 0; i < N; i++) for (int k = 0; k < 8; k++) b[8 * (N - 1 - i) + k] = (u8)(v[i] >> (56 - 8 * k));
}

// field element type bound to a global context
template <int N, const MontCtx<N>* CTX> struct MF {
    u64 v[N];
    static const MontCtx<N>& C() { return *CTX; }
    static MF zero() { MF r; memset(r.v, 0, sizeof r.v); return r; }
    static MF one() { MF r; memcpy(r.v, C().one, sizeof r.v); return r; }
    static MF from_plain(const u64* x) { MF r; mont_mul<N>(r.v, x, C().r2, C()); return r; }
    void to_plain(u64* x) const { u64 o[N] = {1}; mont_mul<N>(x, v, o, C()); }
    MF operator*(const MF& b) const { MF r; mont_mul<N>(r.v, v, b.v, C()); return r; }
    MF operator+(const MF& b) const { MF r; mod_add<N>(r.v, v, b.v, C().p); return r; }
    MF operator-(const MF& b) const { MF r; mod_sub<N>(r.v, v, b.v, C().p); return r; }
    MF sqr() const { return *this * *this; }
    MF dbl() const { return *this + *this; }
    MF neg() const { return zero() - *this; }
    bool isz() const { return is_zero<N>(v); }
    bool operator==(const MF& b) const { return memcmp(v, b.v, sizeof v) == 0; }
    MF pow(const u64* e) const {   // square-and-multiply, 4-bit window
        MF tab[16]; tab[0] = one(); for (int i = 1; i < 16; i++) tab[i] = tab[i - 1] * *this;
        MF acc = one();
        for (int w = 16 * N - 1; w >= 0; w--) {
            acc = acc.sqr().sqr().sqr().sqr();
            unsigned nib = (e[w / 16] >> ((w % 16) * 4)) & 15;
            if (nib) acc = acc * tab[nib];
        }
        return acc;
    }
    MF inv() const {   // Fermat: a^(p-2)
        u64 e[N]; memcpy(e, C().p, sizeof e);
        e[0] -= 2;   // p is odd and > 2: no borrow
        return pow(e);
    }
    // Scalar::invert_vartime - Stein's binary extended GCD on plain values, the schedule of
    // k256/src/arithmetic/scalar.rs:467-516 and p256/src/arithmetic/scalar.rs:365-409 (u-loop, v-loop, sub-step; halving
    // an odd A as (A >> 1) + (m >> 1) + 1).  This is what hazmat::verify_prehashed calls for s^-1.  0 -> 0.
    MF inv_vartime() const {
        const u64* m = C().p;
        u64 u[N], v[N], A[N] = {1}, Cc[N] = {0}, half[N];
        to_plain(u);
        memcpy(v, m, sizeof v);
        for (int j = 0; j < N; j++) half[j] = (m[j] >> 1) | (j + 1 < N ? m[j + 1] << 63 : 0);   // FRAC_MODULUS_2
        auto shr1 = [](u64* x) { for (int j = 0; j < N; j++) x[j] = (x[j] >> 1) | (j + 1 < N ? x[j + 1] << 63 : 0); };
        auto halve_mod = [&](u64* x) {
            const bool odd = x[0] & 1;
            shr1(x);
            if (odd) {   // + (m - 1) / 2 + 1: stays below m
                u64 c = 1;
                for (int j = 0; j < N; j++) { u128 t = (u128)x[j] + half[j] + c; x[j] = (u64)t; c = (u64)(t >> 64); }
            }
        };
        if (is_zero<N>(u)) return zero();
        while (!is_zero<N>(u)) {
            while (!(u[0] & 1)) { shr1(u); halve_mod(A); }
            while (!(v[0] & 1)) { shr1(v); halve_mod(Cc); }
            if (geq<N>(u, v)) {
                u64 bw = 0;
                for (int j = 0; j < N; j++) { u128 t = (u128)u[j] - v[j] - bw; u[j] = (u64)t; bw = (u64)(t >> 64) & 1; }
                mod_sub<N>(A, A, Cc, m);
            } else {
                u64 bw = 0;
                for (int j = 0; j < N; j++) { u128 t = (u128)v[j] - u[j] - bw; v[j] = (u64)t; bw = (u64)(t >> 64) & 1; }
                mod_sub<N>(Cc, Cc, A, m);
            }
        }
        return from_plain(Cc);
    }
};

// =============================================================================================
// contexts
static MontCtx<4> K256N, P256P, P256N, SM2P, SM2N;
static MontCtx<6> P384P, P384N;
static const u64 k256n_[4] = {0xBFD25E8CD0364141ull, 0xBAAEDCE6AF48A03Bull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull};
static const u64 p256p_[4] = {0xFFFFFFFFFFFFFFFFull, 0x00000000FFFFFFFFull, 0x0000000000000000ull, 0xFFFFFFFF00000001ull};
static const u64 p256n_[4] = {0xF3B9CAC2FC632551ull, 0xBCE6FAADA7179E84ull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFF00000000ull};
static const u64 sm2p_[4] = {0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFF00000000ull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFEFFFFFFFFull};
static const u64 sm2n_[4] = {0x53BBF40939D54123ull, 0x7203DF6B21C6052Bull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFEFFFFFFFFull};
static const u64 p384p_[6] = {0x00000000FFFFFFFFull, 0xFFFFFFFF00000000ull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull};
static const u64 p384n_[6] = {0xECEC196ACCC52973ull, 0x581A0DB248B0A77Aull, 0xC7634D81F4372DDFull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull};

// =============================================================================================
// primeorder curves (a = -3)
template <int N, const MontCtx<N>* FP, const MontCtx<N>* FN> struct PrimeOrder {
    typedef MF<N, FP> F;
    typedef MF<N, FN> S;
    struct Pt { F x, y, z; };
    F b, gx, gy;
    u64 n[N];   // group order
    static Pt identity() { return Pt{F::zero(), F::one(), F::zero()}; }

    // RCB Alg 4 (primeorder/src/point_arithmetic.rs:209-238)
    Pt add(const Pt& l, const Pt& r) const {
        F xx = l.x * r.x, yy = l.y * r.y, zz = l.z * r.z;
        F xy = ((l.x + l.y) * (r.x + r.y)) - (xx + yy);
        F yz = ((l.y + l.z) * (r.y + r.z)) - (yy + zz);
        F xz = ((l.x + l.z) * (r.x + r.z)) - (xx + zz);
        F bzz = xz - (b * zz);
        F bzz3 = bzz.dbl() + bzz;
        F yym = yy - bzz3, yyp = yy + bzz3;
        F zz3 = zz.dbl() + zz;
        F bxz = (b * xz) - (zz3 + xx);
        F bxz3 = bxz.dbl() + bxz;
        F xx3m = xx.dbl() + xx - zz3;
        return Pt{(yyp * xy) - (yz * bxz3), (yyp * yym) + (xx3m * bxz3), (yym * yz) + (xy * xx3m)};
    }
    // RCB Alg 6 (primeorder/src/point_arithmetic.rs:286-317)
    Pt dbl(const Pt& p) const {
        F xx = p.x.sqr(), yy = p.y.sqr(), zz = p.z.sqr();
        F xy2 = (p.x * p.y).dbl(), xz2 = (p.x * p.z).dbl();
        F bzz = (b * zz) - xz2;
        F bzz3 = bzz.dbl() + bzz;
        F yym = yy - bzz3, yyp = yy + bzz3;
        F yf = yyp * yym, xf = yym * xy2;
        F zz3 = zz.dbl() + zz;
        F bxz2 = (b * xz2) - (zz3 + xx);
        F bxz6 = bxz2.dbl() + bxz2;
        F xx3m = xx.dbl() + xx - zz3;
        F y = yf + (xx3m * bxz6);
        F yz2 = (p.y * p.z).dbl();
        F x = xf - (bxz6 * yz2);
        F z = (yz2 * yy).dbl().dbl();
        return Pt{x, y, z};
    }
    // 4-bit fixed window, 16-entry table, add every window (primeorder/src/projective.rs:106-150);
    // the constant-time table scan is a plain index here (timing baseline only)
    Pt mul(const Pt& p, const u64* k) const {
        Pt pc[16];
        pc[0] = identity(); pc[1] = p;
        for (int i = 2; i < 16; i++) pc[i] = (i % 2 == 0) ? dbl(pc[i / 2]) : add(pc[i - 1], p);
        Pt q = identity();
        for (int pos = 64 * N - 4;; pos -= 4) {
            unsigned slot = (k[pos / 64] >> (pos % 64)) & 15;
            Pt t = identity();
            for (int i = 1; i < 16; i++) if ((unsigned)i == slot) t = pc[i];   // scan kept for cost parity
            q = add(q, t);
            if (pos == 0) break;
            q = dbl(dbl(dbl(dbl(q))));
        }
        return q;
    }
    bool on_curve(const F& x, const F& y) const {
        F three = F::one() + F::one() + F::one();
        return y.sqr() == (x.sqr() * x) - (three * x) + b;
    }
    // to_affine + SEC1 (primeorder/src/projective.rs:62-74, affine.rs:340-358)
    void encode(u8* out, const Pt& p, bool compress) const {
        int fb = 8 * N, sz = 1 + (compress ? fb : 2 * fb);
        if (p.z.isz()) { memset(out, 0, sz); return; }
        F zi = p.z.inv();
        u64 t[N];
        (p.x * zi).to_plain(t); store_be<N>(out + 1, t);
        (p.y * zi).to_plain(t);
        if (compress) out[0] = 2 + (t[0] & 1);
        else { out[0] = 4; store_be<N>(out + 1 + fb, t); }
    }
    void reduce_scalar(u64* k) const {   // Reduce<Uint>: one conditional subtraction
        if (geq<N>(k, n)) { u64 bw = 0; for (int j = 0; j < N; j++) { u128 s = (u128)k[j] - n[j] - bw; k[j] = (u64)s; bw = (u64)(s >> 64) & 1; } }
    }
    bool verify(const u8* q, const u8* zb, const u8* rs) const {
        u64 r[N], s[N], z[N], qx[N], qy[N];
        load_be<N>(r, rs); load_be<N>(s, rs + 8 * N); load_be<N>(z, zb); load_be<N>(qx, q); load_be<N>(qy, q + 8 * N);
        if (is_zero<N>(r) || is_zero<N>(s) || geq<N>(r, n) || geq<N>(s, n)) return false;
        if (geq<N>(qx, F::C().p) || geq<N>(qy, F::C().p)) return false;
        F X = F::from_plain(qx), Y = F::from_plain(qy);
        if (!on_curve(X, Y)) return false;
        reduce_scalar(z);
        S w = S::from_plain(s).inv_vartime();   // hazmat::verify_prehashed: s.invert_vartime()
        u64 u1[N], u2[N];
        (S::from_plain(z) * w).to_plain(u1);
        (S::from_plain(r) * w).to_plain(u2);
        Pt G{gx, gy, F::one()}, Q{X, Y, F::one()};
        Pt R = add(mul(G, u1), mul(Q, u2));   // lincomb default body (primeorder/src/projective.rs:415-420)
        if (R.z.isz()) return false;
        u64 xa[N];
        (R.x * R.z.inv()).to_plain(xa);
        reduce_scalar(xa);
        return memcmp(xa, r, sizeof xa) == 0;
    }
};

static PrimeOrder<4, &P256P, &P256N> CP256;
static PrimeOrder<4, &SM2P, &SM2N> CSM2;
static PrimeOrder<6, &P384P, &P384N> CP384;

// =============================================================================================
// secp256k1: 5x52 field
struct Fe {
    u64 n[5];
};
static const u64 M52 = 0xFFFFFFFFFFFFFull;

// mul_inner (field_5x52.rs:288-449): column sums in two u128 accumulators, high half folded with R = 2^256 mod p << 4
static inline Fe fe_mul(const Fe& A, const Fe& B) {
    const u64 *a = A.n, *b = B.n;
    const u128 R = 0x1000003D10ull;
    u128 c, d;
    u64 t3, t4, tx, u0, r0, r1, r2;
    d = (u128)a[0] * b[3] + (u128)a[1] * b[2] + (u128)a[2] * b[1] + (u128)a[3] * b[0];
    c = (u128)a[4] * b[4];
    d += (c & M52) * R; c >>= 52;
    t3 = (u64)d & M52; d >>= 52;
    d += (u128)a[0] * b[4] + (u128)a[1] * b[3] + (u128)a[2] * b[2] + (u128)a[3] * b[1] + (u128)a[4] * b[0];
    d += (u128)(u64)c * R;
    t4 = (u64)d & M52; d >>= 52;
    tx = t4 >> 48; t4 &= (M52 >> 4);
    c = (u128)a[0] * b[0];
    d += (u128)a[1] * b[4] + (u128)a[2] * b[3] + (u128)a[3] * b[2] + (u128)a[4] * b[1];
    u0 = (u64)d & M52; d >>= 52;
    u0 = (u0 << 4) | tx;
    c += (u128)u0 * (u64)(R >> 4);
    r0 = (u64)c & M52; c >>= 52;
    c += (u128)a[0] * b[1] + (u128)a[1] * b[0];
    d += (u128)a[2] * b[4] + (u128)a[3] * b[3] + (u128)a[4] * b[2];
    c += (d & M52) * R; d >>= 52;
    r1 = (u64)c & M52; c >>= 52;
    c += (u128)a[0] * b[2] + (u128)a[1] * b[1] + (u128)a[2] * b[0];
    d += (u128)a[3] * b[4] + (u128)a[4] * b[3];
    c += (d & M52) * R; d >>= 52;
    r2 = (u64)c & M52; c >>= 52;
    c += (u128)(u64)d * R + t3;
    Fe r;
    r.n[0] = r0; r.n[1] = r1; r.n[2] = r2;
    r.n[3] = (u64)c & M52; c >>= 52;
    r.n[4] = (u64)c + t4;
    return r;
}
static inline Fe fe_sqr(const Fe& a) { return fe_mul(a, a); }   // field_5x52.rs:462-464
static inline Fe fe_add(const Fe& a, const Fe& b) { Fe r; for (int i = 0; i < 5; i++) r.n[i] = a.n[i] + b.n[i]; return r; }
static inline Fe fe_dbl(const Fe& a) { return fe_add(a, a); }
static inline Fe fe_muls(const Fe& a, u64 k) { Fe r; for (int i = 0; i < 5; i++) r.n[i] = a.n[i] * k; return r; }
static inline Fe fe_neg(const Fe& a, u64 mag) {   // field_5x52.rs:252-260
    u64 m = mag + 1;
    Fe r;
    r.n[0] = 0xFFFFEFFFFFC2Full * 2 * m - a.n[0];
    r.n[1] = M52 * 2 * m - a.n[1]; r.n[2] = M52 * 2 * m - a.n[2]; r.n[3] = M52 * 2 * m - a.n[3];
    r.n[4] = 0x0FFFFFFFFFFFFull * 2 * m - a.n[4];
    return r;
}
static inline Fe fe_addcorr(const Fe& a, u64 x) {   // add_modulus_correction, field_5x52.rs:133-152
    Fe r;
    u64 t0 = a.n[0] + x * 0x1000003D1ull;
    u64 t1 = a.n[1] + (t0 >> 52); t0 &= M52;
    u64 t2 = a.n[2] + (t1 >> 52); t1 &= M52;
    u64 t3 = a.n[3] + (t2 >> 52); t2 &= M52;
    u64 t4 = a.n[4] + (t3 >> 52); t3 &= M52;
    r.n[0] = t0; r.n[1] = t1; r.n[2] = t2; r.n[3] = t3; r.n[4] = t4;
    return r;
}
static inline Fe fe_nw(const Fe& a) {   // normalize_weak, field_5x52.rs:173-184
    Fe t = a;
    u64 x = t.n[4] >> 48;
    t.n[4] &= 0x0FFFFFFFFFFFFull;
    return fe_addcorr(t, x);
}
static inline Fe fe_norm(const Fe& a) {   // normalize, field_5x52.rs:189-206
    Fe r = fe_nw(a);
    u64 m = r.n[1] & r.n[2] & r.n[3];
    bool over = (r.n[4] >> 48) || ((r.n[4] == 0x0FFFFFFFFFFFFull) && (m == M52) && (r.n[0] >= 0xFFFFEFFFFFC2Full));
    if (over) { r = fe_addcorr(r, 1); r.n[4] &= 0x0FFFFFFFFFFFFull; }
    return r;
}
static inline bool fe_is_zero_n(const Fe& a) { Fe r = fe_norm(a); return (r.n[0] | r.n[1] | r.n[2] | r.n[3] | r.n[4]) == 0; }
static inline Fe fe_from_u64x4(const u64* w) {
    Fe r;
    r.n[0] = w[0] & M52;
    r.n[1] = ((w[0] >> 52) | (w[1] << 12)) & M52;
    r.n[2] = ((w[1] >> 40) | (w[2] << 24)) & M52;
    r.n[3] = ((w[2] >> 28) | (w[3] << 36)) & M52;
    r.n[4] = w[3] >> 16;
    return r;
}
static inline void fe_to_u64x4(u64* w, const Fe& a) {   // a normalized
    w[0] = a.n[0] | (a.n[1] << 52);
    w[1] = (a.n[1] >> 12) | (a.n[2] << 40);
    w[2] = (a.n[2] >> 24) | (a.n[3] << 28);
    w[3] = (a.n[3] >> 36) | (a.n[4] << 16);
}
static inline Fe fe_sqrn(Fe a, int n) { for (int i = 0; i < n; i++) a = fe_sqr(a); return a; }
static Fe fe_inv(const Fe& a) {   // k256/src/arithmetic/field.rs:187-216
    Fe x2 = fe_mul(fe_sqr(a), a), x3 = fe_mul(fe_sqr(x2), a);
    Fe x6 = fe_mul(fe_sqrn(x3, 3), x3), x9 = fe_mul(fe_sqrn(x6, 3), x3), x11 = fe_mul(fe_sqrn(x9, 2), x2);
    Fe x22 = fe_mul(fe_sqrn(x11, 11), x11), x44 = fe_mul(fe_sqrn(x22, 22), x22), x88 = fe_mul(fe_sqrn(x44, 44), x44);
    Fe x176 = fe_mul(fe_sqrn(x88, 88), x88), x220 = fe_mul(fe_sqrn(x176, 44), x44), x223 = fe_mul(fe_sqrn(x220, 3), x3);
    Fe t = fe_mul(fe_sqrn(x223, 23), x22);
    t = fe_mul(fe_sqrn(t, 5), a);
    t = fe_mul(fe_sqrn(t, 3), x2);
    return fe_mul(fe_sqrn(t, 2), a);
}

struct KPt { Fe x, y, z; };
static const Fe FE_ZERO = {{0, 0, 0, 0, 0}}, FE_ONE = {{1, 0, 0, 0, 0}};
static inline KPt k_identity() { return KPt{FE_ZERO, FE_ONE, FE_ZERO}; }
// projective.rs:96-161 (non-zkvm branch), magnitudes as in the reference
static KPt k_add(const KPt& p, const KPt& q) {
    Fe xx = fe_mul(p.x, q.x), yy = fe_mul(p.y, q.y), zz = fe_mul(p.z, q.z);
    Fe nxy = fe_neg(fe_add(xx, yy), 2), nyz = fe_neg(fe_add(yy, zz), 2), nxz = fe_neg(fe_add(xx, zz), 2);
    Fe xy = fe_add(fe_mul(fe_add(p.x, p.y), fe_add(q.x, q.y)), nxy);
    Fe yz = fe_add(fe_mul(fe_add(p.y, p.z), fe_add(q.y, q.z)), nyz);
    Fe xz = fe_add(fe_mul(fe_add(p.x, p.z), fe_add(q.x, q.z)), nxz);
    Fe bzz = fe_muls(zz, 7);
    Fe bzz3 = fe_nw(fe_add(fe_dbl(bzz), bzz));
    Fe yym = fe_add(yy, fe_neg(bzz3, 1)), yyp = fe_add(yy, bzz3);
    Fe byz = fe_nw(fe_muls(yz, 7));
    Fe byz3 = fe_nw(fe_add(fe_dbl(byz), byz));
    Fe xx3 = fe_add(fe_dbl(xx), xx);
    Fe bxx9 = fe_nw(fe_muls(fe_nw(fe_add(fe_dbl(xx3), xx3)), 7));
    KPt r;
    r.x = fe_nw(fe_add(fe_mul(xy, yym), fe_neg(fe_mul(byz3, xz), 1)));
    r.y = fe_nw(fe_add(fe_mul(yyp, yym), fe_mul(bxx9, xz)));
    r.z = fe_nw(fe_add(fe_mul(yz, yyp), fe_mul(xx3, xy)));
    return r;
}
// projective.rs:225-274
static KPt k_dbl(const KPt& p) {
    Fe yy = fe_sqr(p.y), zz = fe_sqr(p.z), xy2 = fe_dbl(fe_mul(p.x, p.y));
    Fe bzz = fe_muls(zz, 7);
    Fe bzz3 = fe_nw(fe_add(fe_dbl(bzz), bzz));
    Fe bzz9 = fe_nw(fe_add(fe_dbl(bzz3), bzz3));
    Fe yym = fe_add(yy, fe_neg(bzz9, 1)), yyp = fe_add(yy, bzz3);
    Fe yyzz = fe_mul(yy, zz);
    Fe yyzz8 = fe_dbl(fe_dbl(fe_dbl(yyzz)));
    Fe t = fe_muls(fe_nw(fe_add(fe_dbl(yyzz8), yyzz8)), 7);
    KPt r;
    r.x = fe_mul(xy2, yym);
    r.y = fe_nw(fe_add(fe_mul(yym, yyp), t));
    r.z = fe_nw(fe_dbl(fe_dbl(fe_dbl(fe_mul(fe_mul(yy, p.y), p.z)))));
    return r;
}
static inline KPt k_neg(const KPt& p) { return KPt{p.x, fe_nw(fe_neg(p.y, 1)), p.z}; }

typedef MF<4, &K256N> KS;
static const u64 K_MINUS_LAMBDA[4] = {0xE0CFC810B51283CFull, 0xA880B9FC8EC739C2ull, 0x5AD9E3FD77ED9BA4ull, 0xAC9C52B33FA3CF1Full};
static const u64 K_MINUS_B1[4] = {0x6F547FA90ABFE4C3ull, 0xE4437ED6010E8828ull, 0, 0};
static const u64 K_MINUS_B2[4] = {0xD765CDA83DB1562Cull, 0x8A280AC50774346Dull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull};
static const u64 K_G1[4] = {0xE893209A45DBB031ull, 0x3DAA8A1471E8CA7Full, 0xE86C90E49284EB15ull, 0x3086D221A7D46BCDull};
static const u64 K_G2[4] = {0x1571B4AE8AC47F71ull, 0x221208AC9DF506C6ull, 0x6F547FA90ABFE4C4ull, 0xE4437ED6010E8828ull};
static const u64 K_BETA[4] = {0xC1396C28719501EEull, 0x9CF0497512F58995ull, 0x6E64479EAC3434E9ull, 0x7AE96A2B657C0710ull};
static const u64 K_P[4] = {0xFFFFFFFEFFFFFC2Full, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull};
static const u64 K_GX[4] = {0x59F2815B16F81798ull, 0x029BFCDB2DCE28D9ull, 0x55A06295CE870B07ull, 0x79BE667EF9DCBBACull};
static const u64 K_GY[4] = {0x9C47D08FFB10D4B8ull, 0xFD17B448A6855419ull, 0x5DA4FBFC0E1108A8ull, 0x483ADA7726A3C465ull};
static KS ks_minus_lambda, ks_minus_b1, ks_minus_b2;
static Fe k_beta;
static KPt k_gen;
static KPt k_gen_table[33][8];

// mul_shift_vartime(k, g, 384) (wide64.rs:64-119): top 128 bits of the 512-bit product, rounded
static void mul_shift_384(u64* out, const u64* k, const u64* g) {
    u64 t[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)k[j] * g[i] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
        t[i + 4] = (u64)c;
    }
    u64 rnd = t[5] >> 63;
    u128 s = (u128)t[6] + rnd;
    out[0] = (u64)s; s >>= 64; s += t[7]; out[1] = (u64)s; out[2] = (u64)(s >> 64); out[3] = 0;
}
struct KSplit { u64 r1[4], r2[4]; bool s1, s2; };
static void k_decompose(KSplit& o, const u64* k) {   // mul.rs:260-268 + 350-362
    u64 q1[4], q2[4], hn[4];
    mul_shift_384(q1, k, K_G1);
    mul_shift_384(q2, k, K_G2);
    KS c1 = KS::from_plain(q1) * ks_minus_b1, c2 = KS::from_plain(q2) * ks_minus_b2;
    KS r2 = c1 + c2;
    KS r1 = KS::from_plain(k) + r2 * ks_minus_lambda;
    r1.to_plain(o.r1); r2.to_plain(o.r2);
    for (int i = 0; i < 4; i++) hn[i] = (k256n_[i] >> 1) | (i < 3 ? k256n_[i + 1] << 63 : 0);
    o.s1 = !geq<4>(hn, o.r1);
    o.s2 = !geq<4>(hn, o.r2);
    if (o.s1) { u64 t[4]; mod_sub<4>(t, k256n_, o.r1, k256n_); memcpy(o.r1, t, 32); }
    if (o.s2) { u64 t[4]; mod_sub<4>(t, k256n_, o.r2, k256n_); memcpy(o.r2, t, 32); }
}
template <int D> static void radix16(int8_t* d, const u64* x) {   // mul.rs:281-304
    for (int i = 0; i < D; i++) d[i] = 0;
    for (int i = 0; i < D - 1; i++) d[i] = (int8_t)((x[i / 16] >> ((i % 16) * 4)) & 15);
    for (int i = 0; i < D - 1; i++) { int8_t carry = (int8_t)((d[i] + 8) >> 4); d[i] -= (int8_t)(carry << 4); d[i + 1] += carry; }
}
static inline KPt k_select(const KPt* tab, int8_t x) {   // LookupTable::select, mul.rs:92-127 (scan kept)
    int a = x < 0 ? -x : x;
    KPt t = k_identity();
    for (int j = 1; j < 9; j++) if (j == a) t = tab[j - 1];
    return x < 0 ? k_neg(t) : t;
}
static void k_table(KPt* tab, const KPt& p) { tab[0] = p; for (int j = 0; j < 7; j++) tab[j + 1] = k_add(p, tab[j]); }   // mul.rs:65-73
// lincomb, mul.rs:342-393
static KPt k_lincomb(const KPt* xs, const u64 (*ks)[4], int n) {
    std::vector<KPt> tabs((size_t)n * 16);
    std::vector<int8_t> digs((size_t)n * 66);
    for (int i = 0; i < n; i++) {
        KSplit s;
        k_decompose(s, ks[i]);
        KPt xb = xs[i];
        xb.x = fe_mul(xb.x, k_beta);
        k_table(&tabs[(size_t)i * 16], s.s1 ? k_neg(xs[i]) : xs[i]);
        k_table(&tabs[(size_t)i * 16 + 8], s.s2 ? k_neg(xb) : xb);
        radix16<33>(&digs[(size_t)i * 66], s.r1);
        radix16<33>(&digs[(size_t)i * 66 + 33], s.r2);
    }
    KPt acc = k_identity();
    for (int i = 32; i >= 0; i--) {
        if (i != 32) for (int j = 0; j < 4; j++) acc = k_dbl(acc);
        for (int c = 0; c < n; c++) {
            acc = k_add(acc, k_select(&tabs[(size_t)c * 16], digs[(size_t)c * 66 + i]));
            acc = k_add(acc, k_select(&tabs[(size_t)c * 16 + 8], digs[(size_t)c * 66 + 33 + i]));
        }
    }
    return acc;
}
// mul_by_generator with precomputed tables, mul.rs:397-439
static KPt k_mul_gen(const u64* k) {
    int8_t d[65];
    radix16<65>(d, k);
    KPt acc = k_select(k_gen_table[32], d[64]), acc2 = k_identity();
    for (int i = 31; i >= 0; i--) {
        acc2 = k_add(acc2, k_select(k_gen_table[i], d[2 * i + 1]));
        acc = k_add(acc, k_select(k_gen_table[i], d[2 * i]));
    }
    for (int j = 0; j < 4; j++) acc2 = k_dbl(acc2);
    return k_add(acc, acc2);
}
static void k_reduce_scalar(u64* k) {
    if (geq<4>(k, k256n_)) { u64 t[4]; u64 bw = 0; for (int j = 0; j < 4; j++) { u128 s = (u128)k[j] - k256n_[j] - bw; t[j] = (u64)s; bw = (u64)(s >> 64) & 1; } memcpy(k, t, 32); }
}
static void k_encode(u8* out, const KPt& p, bool compress) {
    int sz = compress ? 33 : 65;
    if (fe_is_zero_n(p.z)) { memset(out, 0, sz); return; }
    Fe zi = fe_inv(p.z);
    Fe x = fe_norm(fe_mul(p.x, zi)), y = fe_norm(fe_mul(p.y, zi));
    u64 t[4];
    fe_to_u64x4(t, x); store_be<4>(out + 1, t);
    fe_to_u64x4(t, y);
    if (compress) out[0] = 2 + (t[0] & 1);
    else { out[0] = 4; store_be<4>(out + 33, t); }
}
static bool k_load_affine(KPt& p, const u8* xy) {
    u64 x[4], y[4];
    load_be<4>(x, xy); load_be<4>(y, xy + 32);
    if (geq<4>(x, K_P) || geq<4>(y, K_P)) return false;
    p.x = fe_from_u64x4(x); p.y = fe_from_u64x4(y); p.z = FE_ONE;
    Fe lhs = fe_sqr(p.y), rhs = fe_add(fe_mul(fe_sqr(p.x), p.x), fe_muls(FE_ONE, 7));
    return fe_is_zero_n(fe_add(lhs, fe_neg(rhs, 2)));
}
static bool k_verify(const u8* q, const u8* zb, const u8* rs) {
    u64 r[4], s[4], z[4], hn[4];
    load_be<4>(r, rs); load_be<4>(s, rs + 32); load_be<4>(z, zb);
    if (is_zero<4>(r) || is_zero<4>(s) || geq<4>(r, k256n_) || geq<4>(s, k256n_)) return false;
    for (int i = 0; i < 4; i++) hn[i] = (k256n_[i] >> 1) | (i < 3 ? k256n_[i + 1] << 63 : 0);
    if (!geq<4>(hn, s)) return false;   // high-s rejected: k256/src/ecdsa.rs:203-205
    KPt Q;
    if (!k_load_affine(Q, q)) return false;
    k_reduce_scalar(z);
    KS w = KS::from_plain(s).inv_vartime();   // hazmat::verify_prehashed: s.invert_vartime() (k256 scalar.rs:467-516)
    u64 u[2][4];
    (KS::from_plain(z) * w).to_plain(u[0]);
    (KS::from_plain(r) * w).to_plain(u[1]);
    KPt xs[2] = {k_gen, Q};
    KPt R = k_lincomb(xs, u, 2);
    if (fe_is_zero_n(R.z)) return false;
    Fe x = fe_norm(fe_mul(R.x, fe_inv(R.z)));
    u64 xa[4];
    fe_to_u64x4(xa, x);
    k_reduce_scalar(xa);
    return memcmp(xa, r, 32) == 0;
}

// =============================================================================================
static bool g_init = false;
template <int N, class FT> static FT from_hex(const char* hex) {
    u8 b[8 * N];
    for (int i = 0; i < 8 * N; i++) { unsigned v; sscanf(hex + 2 * i, "%2x", &v); b[i] = (u8)v; }
    u64 w[N];
    load_be<N>(w, b);
    return FT::from_plain(w);
}
static void init_all() {
    if (g_init) return;
    K256N.init(k256n_); P256P.init(p256p_); P256N.init(p256n_); SM2P.init(sm2p_); SM2N.init(sm2n_); P384P.init(p384p_); P384N.init(p384n_);
    ks_minus_lambda = KS::from_plain(K_MINUS_LAMBDA); ks_minus_b1 = KS::from_plain(K_MINUS_B1); ks_minus_b2 = KS::from_plain(K_MINUS_B2);
    k_beta = fe_from_u64x4(K_BETA);
    k_gen = KPt{fe_from_u64x4(K_GX), fe_from_u64x4(K_GY), FE_ONE};
    KPt g = k_gen;
    for (int i = 0; i < 33; i++) { k_table(k_gen_table[i], g); for (int j = 0; j < 8; j++) g = k_dbl(g); }   // mul.rs:400-413
    typedef MF<4, &P256P> F2; typedef MF<4, &SM2P> F3; typedef MF<6, &P384P> F4;
    CP256.b = from_hex<4, F2>("5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B");
    CP256.gx = from_hex<4, F2>("6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296");
    CP256.gy = from_hex<4, F2>("4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5");
    memcpy(CP256.n, p256n_, 32);
    CSM2.b = from_hex<4, F3>("28E9FA9E9D9F5E344D5A9E4BCF6509A7F39789F515AB8F92DDBCBD414D940E93");
    CSM2.gx = from_hex<4, F3>("32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7");
    CSM2.gy = from_hex<4, F3>("BC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0");
    memcpy(CSM2.n, sm2n_, 32);
    CP384.b = from_hex<6, F4>("B3312FA7E23EE7E4988E056BE3F82D19181D9C6EFE8141120314088F5013875AC656398D8A2ED19D2A85C8EDD3EC2AEF");
    CP384.gx = from_hex<6, F4>("AA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7");
    CP384.gy = from_hex<6, F4>("3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F");
    memcpy(CP384.n, p384n_, 48);
    g_init = true;
}

template <class CV, int N> static void po_mul_batch(const CV& cv, long n, const u8* pts, const u8* k, u8* out, int compress, bool gen) {
    const int fb = 8 * N, slot = 1 + (compress ? fb : 2 * fb);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        u64 kk[N];
        load_be<N>(kk, k + (size_t)i * fb);
        cv.reduce_scalar(kk);
        typename CV::Pt P;
        bool ok = true;
        if (gen) P = typename CV::Pt{cv.gx, cv.gy, CV::F::one()};
        else {
            u64 x[N], y[N];
            load_be<N>(x, pts + (size_t)i * 2 * fb); load_be<N>(y, pts + (size_t)i * 2 * fb + fb);
            ok = !geq<N>(x, CV::F::C().p) && !geq<N>(y, CV::F::C().p);
            P = typename CV::Pt{CV::F::from_plain(x), CV::F::from_plain(y), CV::F::one()};
            ok = ok && cv.on_curve(P.x, P.y);
        }
        if (!ok) { memset(out + (size_t)i * slot, 0, slot); continue; }
        cv.encode(out + (size_t)i * slot, cv.mul(P, kk), compress != 0);
    }
}

extern "C" {
int port_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void port_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
// out slots as the C ABI (include/ecb200.h); compress: 0/1
int port_mul_gen(int curve, long n, const u8* k, u8* out, int compress) {
    init_all();
    if (curve == 0) {
        const int slot = compress ? 33 : 65;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) {
            u64 kk[4];
            load_be<4>(kk, k + (size_t)i * 32);
            k_reduce_scalar(kk);
            k_encode(out + (size_t)i * slot, k_mul_gen(kk), compress != 0);
        }
    } else if (curve == 1) po_mul_batch<decltype(CP256), 4>(CP256, n, nullptr, k, out, compress, true);
    else if (curve == 2) po_mul_batch<decltype(CP384), 6>(CP384, n, nullptr, k, out, compress, true);
    else if (curve == 3) po_mul_batch<decltype(CSM2), 4>(CSM2, n, nullptr, k, out, compress, true);
    else return -1;
    return 0;
}
int port_mul_var(int curve, long n, const u8* pts, const u8* k, u8* out, int compress) {
    init_all();
    if (curve == 0) {
        const int slot = compress ? 33 : 65;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) {
            u64 kk[1][4];
            load_be<4>(kk[0], k + (size_t)i * 32);
            k_reduce_scalar(kk[0]);
            KPt P;
            if (!k_load_affine(P, pts + (size_t)i * 64)) { memset(out + (size_t)i * slot, 0, slot); continue; }
            k_encode(out + (size_t)i * slot, k_lincomb(&P, kk, 1), compress != 0);
        }
    } else if (curve == 1) po_mul_batch<decltype(CP256), 4>(CP256, n, pts, k, out, compress, false);
    else if (curve == 2) po_mul_batch<decltype(CP384), 6>(CP384, n, pts, k, out, compress, false);
    else if (curve == 3) po_mul_batch<decltype(CSM2), 4>(CSM2, n, pts, k, out, compress, false);
    else return -1;
    return 0;
}
int port_verify(int curve, long n, const u8* q, const u8* z, const u8* rs, u8* ok) {
    init_all();
    const int fb = curve == 2 ? 48 : 32;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        bool v = false;
        const u8 *qi = q + (size_t)i * 2 * fb, *zi = z + (size_t)i * fb, *ri = rs + (size_t)i * 2 * fb;
        if (curve == 0) v = k_verify(qi, zi, ri);
        else if (curve == 1) v = CP256.verify(qi, zi, ri);
        else if (curve == 2) v = CP384.verify(qi, zi, ri);
        else if (curve == 3) v = CSM2.verify(qi, zi, ri);
        ok[i] = v ? 1 : 0;
    }
    return 0;
}
}
