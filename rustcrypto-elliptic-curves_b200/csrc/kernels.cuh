// kernels.cuh — per-element bodies of every batch kernel, generic over the curve descriptor.
//
// Each body_* function processes the element(s) owned by logical thread `tid` out of `nthreads`;
// the __global__ wrappers (kernels_impl.cuh) just compute tid.  Keeping the bodies free of CUDA
// built-ins lets tests/emu run exactly this logic on the CPU in the GPU-less container
// (logic check only — parity is asserted on the B200 through the C ABI).
//
// Device-side data layouts (HBM):
//   byte inputs / outputs: as the C ABI (include/ecb200.h) — big-endian, AoS, one element after another;
//   internal projective points: 3L little-endian u32 limbs per element (X|Y|Z, field-internal form,
//     i.e. Montgomery for the primeorder curves), AoS, 16-byte aligned;
//   internal affine tables: 2L limbs per entry (x|y);
//   per-row window tables of the primeorder public-input path: 16L limbs per row ({1..8}Q affine) + 7L limbs of Z scratch.
#pragma once
#include "ec.cuh"
#include "jac.cuh"

namespace ecb {

enum : u32 {
    F_CT = 1u,             // secret-scalar path: fixed windows, full table scans, no secret-dependent branch/address
    F_COMPRESSED = 2u,
    F_UNCOMPRESSED = 4u,
    F_PROJ = 8u,           // input points are X||Y||Z projective instead of x||y affine
};

enum : int { NORM_SEC1 = 0, NORM_XY_BYTES = 1, NORM_AFF_LIMBS = 2,
             NORM_AFF_STRIDED = 3 };   // affine limbs, output i at out_limbs + i * stride words (stride passed in `compress`)

// what the two-term public-input pipeline (prep -> main -> optional finish) computes
enum : int {
    VM_ECDSA = 0,     // u1 = z/s, u2 = r/s, Q from x||y;      accept <=> x(R) mod n == r
    VM_SM2DSA = 1,    // u1 = s,   u2 = r+s, Q from x||y;      accept <=> (e + x(R)) mod n == r      (sm2/src/dsa/verifying.rs:130-168)
    VM_SCHNORR = 2,   // u1 = s,   u2 = -e,  Q = lift_x(pk);   R goes to the normalise + finish kernels (k256/src/schnorr/verifying.rs:63-89)
    VM_RECOVER = 3,   // u1 = -z/r, u2 = s/r, Q = decompress(r [+ n], recid); result point = the public key (ecdsa 0.16.9 recovery.rs)
};
enum : int { DEC_SEC1 = 0, DEC_COMPACT = 1 };
enum : int { FIN_SCHNORR = 0, FIN_RECOVER = 1 };

// How the Montgomery-trick bodies (normalise, verify prep, window tables, sign finish) invert the running product of
// the rows a thread owns.  OwnInv: every thread runs its own inversion chain - what the host emulation runs.  The CUDA
// wrappers pass BlockInv (kernels_impl.cuh), which continues the trick across the CTA (warp shuffles + shared memory)
// so that ONE chain serves all 128 threads; a cooperative policy needs every thread of the CTA to reach run().
struct OwnInv {
    static constexpr bool COOPERATIVE = false;
    template <class FF> ECB_DEV static void run(typename FF::E& r, const typename FF::E& a) { FF::inv_trick(r, a); }
};

// KTW: window width of the per-key tables for THIS instantiation (0 = the curve's default width; the launchers instantiate the
// three table kernels a second time with the narrow width, kernels_impl.cuh kt_narrow)
template <class C, int KTW = 0> struct Bodies {
    typedef EC<C> G;
    typedef typename G::Proj Proj;
    typedef typename G::Aff Aff;
    typedef typename G::E E;
    typedef typename C::F F;
    typedef typename C::Fn Fn;
    static constexpr int L = C::L;
    static constexpr int FB = C::FB;

    // internal limb arrays: widest aligned access per coordinate (bigint.cuh ld_words / st_words)
    ECB_DEV static void store_proj(u32* dst, const Proj& p) { st_words<L>(dst, p.X.v); st_words<L>(dst + L, p.Y.v); st_words<L>(dst + 2 * L, p.Z.v); }
    ECB_DEV static void load_proj_limbs(Proj& p, const u32* src) { ld_words<L>(p.X.v, src); ld_words<L>(p.Y.v, src + L); ld_words<L>(p.Z.v, src + 2 * L); }
    ECB_DEV static void load_aff_limbs(Aff& a, const u32* src) { ld_words<L>(a.x.v, src); ld_words<L>(a.y.v, src + L); }

    // ------------------------------------------------------------------ field test hook
    // which: 0 = base field, 1 = scalar field.  op: 0 add, 1 sub, 2 mul, 3 sqr, 4 neg, 5 inv (Fermat), 6 sqrt, 7 inv (divsteps)
    // out = FB bytes canonical; out_ok[i] = 0 when an input was not canonical (>= modulus) or sqrt does not exist
    template <class FF> ECB_DEV static void field_op_one(int op, const u8* a, const u8* b, u8* out, u8* ok) {
        typename FF::E x, y, r;
        u32 t[L];
        load_be<L>(t, a);
        bool v = FF::from_limbs(x, t);
        if (b) { load_be<L>(t, b); v = FF::from_limbs(y, t) && v; } else { y = x; }
        switch (op) {
            case 0: FF::add(r, x, y); break;
            case 1: FF::sub(r, x, y); break;
            case 2: FF::mul(r, x, y); break;
            case 3: FF::sqr(r, x); break;
            case 4: FF::neg(r, x); break;
            case 5: FF::inv(r, x); break;
            case 7: FF::inv_gcd(r, x); break;
            default: {
                FF::sqrt_candidate(r, x);
                typename FF::E c;
                FF::sqr(c, r);
                v = v && FF::eq(c, x);
            }
        }
        FF::to_limbs(t, r);
        store_be<L>(out, t);
        *ok = v ? 1 : 0;
    }
    ECB_DEV static void body_field_op(int tid, int n, int which, int op, const u8* a, const u8* b, u8* out, u8* ok) {
        if (tid >= n) return;
        const u8* pa = a + (size_t)tid * FB;
        const u8* pb = b ? b + (size_t)tid * FB : nullptr;
        if (which == 0) field_op_one<F>(op, pa, pb, out + (size_t)tid * FB, ok + tid);
        else field_op_one<Fn>(op, pa, pb, out + (size_t)tid * FB, ok + tid);
    }

    // ------------------------------------------------------------------ variable-base k*P -> projective
    // pts: n x 2FB (x||y) or n x 3FB (X||Y||Z, F_PROJ); inf: optional n flags (affine identity);
    // k: n x FB; out: n x 3L limbs.  invalid[i] = 1 when the point fails validation (result = identity).
    template <bool CT> ECB_DEV static void body_mul_var(int tid, int n, u32 flags, const u8* pts, const u8* inf,
                                                        const u8* k, u32* out, u8* invalid) {
        if (tid >= n) return;
        Proj p;
        bool ok;
        if (flags & F_PROJ) {
            ok = G::load_proj(p, pts + (size_t)tid * 3 * FB);
        } else {
            Aff a;
            ok = G::load_affine(a, pts + (size_t)tid * 2 * FB);
            G::from_affine(p, a);
            if (inf && inf[tid]) { ok = true; G::set_identity(p); }
        }
        if (!ok) G::set_identity(p);
        u32 kk[L];
        G::load_scalar(kk, k + (size_t)tid * FB);
        Proj r;
        VarMul<C, CT>::run(r, p, kk);
        store_proj(out + (size_t)tid * 3 * L, r);
        if (invalid) invalid[tid] = ok ? 0 : 1;
    }

    // ------------------------------------------------------------------ projective bytes -> internal limbs
    ECB_DEV static void body_load_proj(int tid, int n, const u8* xyz, u32* out, u8* invalid) {
        if (tid >= n) return;
        Proj p;
        bool ok = G::load_proj(p, xyz + (size_t)tid * 3 * FB);
        if (!ok) G::set_identity(p);
        store_proj(out + (size_t)tid * 3 * L, p);
        if (invalid) invalid[tid] = ok ? 0 : 1;
    }

    // ------------------------------------------------------------------ batch normalisation
    // Montgomery's trick over the EPT elements {tid, tid + nthreads, ...} owned by one thread
    // (BatchInvert + batch_normalize_generic: k256/src/arithmetic/projective.rs:350-379,
    // primeorder/src/projective.rs:382-413): zero Z is replaced by ONE in the product chain and the
    // slot becomes IDENTITY.  One field inversion per thread, 3 mul per element for the trick,
    // 2 mul for (X*zinv, Y*zinv).
    static constexpr int EPT = 16;
    // sum_with (optional): a second array of n projective points; element i is first replaced by proj[i] + sum_with[i]
    // (complete addition, written back to proj) - the tail of the split fixed-base path, whose two halves of k*G arrive
    // as separate partial sums (body_gen_half)
    template <class INV = OwnInv>
    ECB_DEV static void body_normalize(int tid, int nthreads, int n, const u32* proj, int mode, int compress,
                                       u8* out_bytes, u8* out_inf, u32* out_limbs, const u32* sum_with = nullptr) {
        E pref[EPT];
        E acc;
        F::set_one(acc);
        int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < EPT; j++) {
            int i = tid + j * nthreads;
            if (i >= n) break;
            E z, one;
            if (sum_with) {
                Proj a, b;
                load_proj_limbs(a, proj + (size_t)i * 3 * L);
                load_proj_limbs(b, sum_with + (size_t)i * 3 * L);
                G::add(a, a, b);
                store_proj(const_cast<u32*>(proj) + (size_t)i * 3 * L, a);
            }
            ld_words<L>(z.v, proj + (size_t)i * 3 * L + 2 * L);
            F::set_one(one);
            bool zz = F::is_zero(z);
            F::select(z, zz, one, z);
            pref[j] = acc;
            F::mul(acc, acc, z);
            cnt++;
        }
        if (!INV::COOPERATIVE && cnt == 0) return;
        E inv;
        INV::template run<F>(inv, acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = cnt - 1; j >= 0; j--) {
            int i = tid + j * nthreads;
            Proj p;
            load_proj_limbs(p, proj + (size_t)i * 3 * L);
            bool isinf = F::is_zero(p.Z);
            E one, z, zinv, x, y;
            F::set_one(one);
            F::select(z, isinf, one, p.Z);
            F::mul(zinv, inv, pref[j]);
            F::mul(inv, inv, z);
            F::mul(x, p.X, zinv);
            F::mul(y, p.Y, zinv);
            if (mode == NORM_SEC1) {
                const int stride = compress ? 1 + FB : 1 + 2 * FB;
                G::encode(out_bytes + (size_t)i * stride, isinf, x, y, compress != 0);
            } else if (mode == NORM_XY_BYTES) {
                u32 t[L];
                u8* o = out_bytes + (size_t)i * 2 * FB;
                if (isinf) {
                    for (int b = 0; b < 2 * FB; b++) o[b] = 0;   // AffinePoint::IDENTITY = (0, 0, infinity=1)
                } else {
                    F::to_limbs(t, x); store_be<L>(o, t);
                    F::to_limbs(t, y); store_be<L>(o + FB, t);
                }
                if (out_inf) out_inf[i] = isinf ? 1 : 0;
            } else {
                u32* o = out_limbs + (size_t)i * (mode == NORM_AFF_STRIDED ? (size_t)compress : (size_t)2 * L);   // stride: a multiple of 2L words
                if (isinf) { F::set_zero(x); F::set_zero(y); }
                st_words<L>(o, x.v); st_words<L>(o + L, y.v);
            }
        }
    }

    // ------------------------------------------------------------------ sum of points (lincomb tail)
    // thread-strided partial sums; the wrapper reduces across the block through shared memory
    ECB_DEV static void body_partial_sum(Proj& acc, int tid, int nthreads, int n, const u32* proj) {
        G::set_identity(acc);
        for (int i = tid; i < n; i += nthreads) {
            Proj p;
            load_proj_limbs(p, proj + (size_t)i * 3 * L);
            G::add(acc, acc, p);
        }
    }

    // ------------------------------------------------------------------ ECDSA verify
    // ecdsa::hazmat::verify_prehashed semantics (call sites k256/src/ecdsa.rs:200-209,
    // p256/src/ecdsa.rs:71-75).  q: x||y, z: FB bytes after bits2field, rs: r||s.
    // gtab: affine multiples 1..NG of G (NG = 8 for k256, 15 for the primeorder curves), 2L limbs each.
    ECB_DEV static bool scalar_in_range(const u32* v) {   // 1 <= v < n
        u32 nn[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) nn[i] = C::n(i);
        // both tests always run and are combined with a bitwise AND: the signing kernel calls this on secret scalars, and
        // `a && b` compiled to a branch on `a` there (tools/ct_sass_audit.py)
        const u32 nz = is_zero_n<L>(v) ? 0u : 1u;
        const u32 lt = geq_n<L>(v, nn) ? 0u : 1u;
        return (nz & lt) != 0u;
    }
    ECB_DEV static bool finish_verify(const Proj& R, const u32* r, bool valid) {
        // accept <=> Z != 0 and (X == r*Z or (r + n < p and X == (r+n)*Z))   (x mod n == r without inversion)
        E re, t;
        bool okr = F::from_limbs(re, r);    // r < n < p
        F::mul(t, re, R.Z);
        bool hit = F::eq(t, R.X);
        u32 pmn[L], nn[L], rn[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) { pmn[i] = C::p_minus_n(i); nn[i] = C::n(i); }
        bool second = !geq_n<L>(r, pmn);    // r + n < p
        add_n<L>(rn, r, nn);
        E rne;
        F::from_limbs(rne, rn);
        F::mul(t, rne, R.Z);
        hit = hit || (second && F::eq(t, R.X));
        return valid && okr && !F::is_zero(R.Z) && hit;
    }
    ECB_DEV static void body_verify(int tid, int n, const u8* q, const u8* z, const u8* rs, const u32* gtab, u8* ok_out) {
        if (tid >= n) return;
        u32 r[L], s[L], zz[L];
        load_be<L>(r, rs + (size_t)tid * 2 * FB);
        load_be<L>(s, rs + (size_t)tid * 2 * FB + FB);
        bool valid = scalar_in_range(r) && scalar_in_range(s);
        if constexpr (C::LOW_S) {   // k256/src/ecdsa.rs:203-205
            u32 hn[L];
            ECB_UNROLL
            for (int i = 0; i < L; i++) hn[i] = C::half_n(i);
            valid = valid && geq_n<L>(hn, s);
        }
        Aff Q;
        valid = G::load_affine(Q, q + (size_t)tid * 2 * FB) && valid;
        if (!valid) { G::generator(Q); s[0] |= 1u; }   // keep the arithmetic well defined; result is masked
        G::load_scalar(zz, z + (size_t)tid * FB);
        typename Fn::E sm, wm;
        Fn::from_limbs(sm, s);
        Fn::inv(wm, sm);
        u32 u1[L], u2[L];
        Fn::mul_plain(u1, zz, wm);
        Fn::mul_plain(u2, r, wm);
        Proj R;
        lincomb_g_q(R, u1, u2, Q, gtab);
        ok_out[tid] = finish_verify(R, r, valid) ? 1 : 0;
    }

    // R = u1*G + u2*Q (public inputs: variable time allowed)
    ECB_DEV static void lincomb_g_q(Proj& R, const u32* u1, const u32* u2, const Aff& Q, const u32* gtab) {
        Proj pq;
        G::from_affine(pq, Q);
        if constexpr (C::A_IS_ZERO) {
            // 2-term GLV lincomb, the N = 2 case of k256/src/arithmetic/mul.rs:342-393
            K256Glv::Split sg, sq;
            K256Glv::decompose(sg, u1);
            K256Glv::decompose(sq, u2);
            typename G::template Table<9> tab;
            G::template build_table<9>(tab, pq);
            Proj acc, e;
            G::set_identity(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int i = 32; i >= 0; i--) {
                if (i != 32) { G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); }
                u32 mag, neg;
                K256Glv::digit(sq.a1, i, mag, neg);
                e = tab.e[mag];
                G::cneg(e, neg ^ sq.neg1);
                G::add(acc, acc, e);
                K256Glv::digit(sq.a2, i, mag, neg);
                e = tab.e[mag];
                K256Glv::endo(e);
                G::cneg(e, neg ^ sq.neg2);
                G::add(acc, acc, e);
                K256Glv::digit(sg.a1, i, mag, neg);
                if (mag) {
                    Aff g;
                    load_aff_limbs(g, gtab + (size_t)(mag - 1) * 2 * L);
                    E ny;
                    F::neg(ny, g.y);
                    F::cmov(g.y, ny, neg ^ sg.neg1);
                    G::add_mixed(acc, acc, g);
                }
                K256Glv::digit(sg.a2, i, mag, neg);
                if (mag) {
                    Aff g;
                    load_aff_limbs(g, gtab + (size_t)(mag - 1) * 2 * L);
                    E beta, ny;
                    ECB_UNROLL
                    for (int l = 0; l < L; l++) beta.v[l] = CurveK256::beta(l);
                    F::mul(g.x, g.x, beta);
                    F::neg(ny, g.y);
                    F::cmov(g.y, ny, neg ^ sg.neg2);
                    G::add_mixed(acc, acc, g);
                }
            }
            R = acc;
        } else {
            // x*k + y*l with shared doublings (primeorder/src/projective.rs:415-420 computes the two
            // products separately; the sum is the same point)
            typename G::template Table<16> tab;
            G::template build_table<16>(tab, pq);
            Proj acc;
            G::set_identity(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int w = 8 * L - 1; w >= 0; w--) {
                if (w != 8 * L - 1) { G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); G::dbl(acc, acc); }
                u32 nq = (u2[w >> 3] >> ((w & 7) * 4)) & 15u;
                { Proj e_ = tab.e[nq]; G::add(acc, acc, e_); }   // copy first: passing the dynamically indexed
                                                                // local-array element by reference to the non-inlined add
                                                                // produced wrong results on sm_100a (nvcc 12.9)
                u32 ng = (u1[w >> 3] >> ((w & 7) * 4)) & 15u;
                if (ng) {
                    Aff g;
                    load_aff_limbs(g, gtab + (size_t)(ng - 1) * 2 * L);
                    G::add_mixed(acc, acc, g);
                }
            }
            R = acc;
        }
    }


    // ------------------------------------------------------------------ ECDSA verify, fast path (v2)
    // Two kernels.  (1) prep: everything mod n — range / low-s checks, w = s^-1 by Montgomery's trick over the
    // PREP_EPT rows a thread owns (one Fermat inversion per thread instead of one per row), u1 = z w,
    // u2 = r w, and for secp256k1 the GLV split of u2.  Results go to a per-row scratch record.
    // (2) main: everything mod p — key validation, u2*Q (Jacobian, jac.cuh), u1*G from the big fixed-base
    // table, inversion-free comparison of x(R) mod n with r.
    static constexpr int PREP_EPT = 16;
    static constexpr int PREP_WORDS = C::A_IS_ZERO ? 20 : 2 * L + 4;   // u32 words per row of scratch
    // value inverted by Montgomery's trick in the prep kernel: s for ECDSA, r for recovery (1 when out of range)
    ECB_DEV static void prep_invertible(u32* v, int mode, const u8* rs_row) {
        load_be<L>(v, rs_row + (mode == VM_RECOVER ? 0 : FB));
        if (!scalar_in_range(v)) { zero_n<L>(v); v[0] = 1; }
    }
    template <class INV = OwnInv>
    ECB_DEV static void body_verify_prep(int tid, int nthreads, int n, int mode, const u8* z, const u8* rs, u32* scratch) {
        typename Fn::E pref[PREP_EPT];
        typename Fn::E acc, inv;
        Fn::set_one(acc);
        const bool need_inv = (mode == VM_ECDSA || mode == VM_RECOVER);
        int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < PREP_EPT; j++) {
            int i = tid + j * nthreads;
            if (i >= n) break;
            if (need_inv) {
                u32 s[L];
                prep_invertible(s, mode, rs + (size_t)i * 2 * FB);
                typename Fn::E sm;
                Fn::from_limbs(sm, s);
                pref[j] = acc;
                Fn::mul(acc, acc, sm);
            }
            cnt++;
        }
        if (!INV::COOPERATIVE && cnt == 0) return;
        if (need_inv) INV::template run<Fn>(inv, acc);   // need_inv depends on the mode only: uniform across the CTA
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = cnt - 1; j >= 0; j--) {
            int i = tid + j * nthreads;
            u32 r[L], s[L], zz[L];
            load_be<L>(r, rs + (size_t)i * 2 * FB);
            load_be<L>(s, rs + (size_t)i * 2 * FB + FB);
            G::load_scalar(zz, z + (size_t)i * FB);
            bool valid = scalar_in_range(s);
            if (mode == VM_SCHNORR) {   // r is a field element: 0 < r < p (k256/src/schnorr.rs:143-160)
                u32 pp[L];
                ECB_UNROLL
                for (int l = 0; l < L; l++) pp[l] = C::p(l);
                valid = valid && !is_zero_n<L>(r) && !geq_n<L>(r, pp);
            } else {
                valid = valid && scalar_in_range(r);
            }
            if (C::LOW_S && (mode == VM_ECDSA || mode == VM_RECOVER)) {   // k256/src/ecdsa.rs:203-205 (recovery ends with a verify)
                u32 hn[L];
                ECB_UNROLL
                for (int l = 0; l < L; l++) hn[l] = C::half_n(l);
                valid = valid && geq_n<L>(hn, s);
            }
            u32 u1[L], u2[L];
            if (need_inv) {
                u32 d[L];
                prep_invertible(d, mode, rs + (size_t)i * 2 * FB);
                typename Fn::E dm, wm;
                Fn::from_limbs(dm, d);
                Fn::mul(wm, inv, pref[j]);       // d_i^-1 (Montgomery form)
                Fn::mul(inv, inv, dm);
                if (mode == VM_ECDSA) {
                    Fn::mul_plain(u1, zz, wm);
                    Fn::mul_plain(u2, r, wm);
                } else {                          // recovery: u1 = -(z / r), u2 = s / r
                    typename Fn::E t, nt;
                    Fn::mul_plain(t.v, zz, wm);
                    Fn::neg(nt, t);
                    copy_n<L>(u1, nt.v);
                    Fn::mul_plain(u2, s, wm);
                }
            } else if (mode == VM_SM2DSA) {       // t = r + s (mod n), t = 0 is rejected
                typename Fn::E a, b, t;
                copy_n<L>(a.v, r); copy_n<L>(b.v, s);
                if (!valid) { Fn::set_zero(a); Fn::set_zero(b); }
                Fn::add(t, a, b);
                valid = valid && !Fn::is_zero(t);
                copy_n<L>(u1, b.v);
                copy_n<L>(u2, t.v);
            } else {                              // Schnorr: R = s*G - e*P
                typename Fn::E e, ne;
                copy_n<L>(e.v, zz);
                Fn::neg(ne, e);
                copy_n<L>(u1, s);
                if (!valid) zero_n<L>(u1);
                copy_n<L>(u2, ne.v);
            }
            u32 o[PREP_WORDS];                    // the record is assembled in registers and leaves as 128-bit stores
            ECB_UNROLL
            for (int l = 0; l < L; l++) o[l] = u1[l];
            if constexpr (C::A_IS_ZERO) {
                K256Glv::Split sp;
                K256Glv::decompose(sp, u2);
                ECB_UNROLL
                for (int l = 0; l < 5; l++) { o[8 + l] = sp.a1[l]; o[13 + l] = sp.a2[l]; }
                o[18] = (valid ? 1u : 0u) | (sp.neg1 & 2u) | (sp.neg2 & 4u);
                o[19] = 0;
            } else {
                ECB_UNROLL
                for (int l = 0; l < L; l++) o[L + l] = u2[l];
                o[2 * L] = valid ? 1u : 0u;
                ECB_UNROLL
                for (int l = 2 * L + 1; l < PREP_WORDS; l++) o[l] = 0;
            }
            st_words<PREP_WORDS>(scratch + (size_t)i * PREP_WORDS, o);
        }
    }

    typedef Jac<C> JJ;
    // accept <=> x(R) mod n == t for a target t < n, without leaving Jacobian coordinates:
    // Z != 0 and (X == t Z^2 or (t + n < p and X == (t + n) Z^2)).  id_x0: the identity counts as x = 0
    // (SM2DSA reads AffinePoint::IDENTITY.x; ECDSA can never accept it because r >= 1).
    ECB_DEV static bool finish_verify_jac(const typename JJ::J& R, const u32* r, bool valid, bool id_x0 = false) {
        if (F::is_zero(R.Z)) return valid && id_x0 && is_zero_n<L>(r);
        E re, t, zz;
        bool okr = F::from_limbs(re, r);
        F::sqr(zz, R.Z);
        F::mul(t, re, zz);
        bool hit = F::eq(t, R.X);
        u32 pmn[L], nn[L], rn[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) { pmn[i] = C::p_minus_n(i); nn[i] = C::n(i); }
        bool second = !geq_n<L>(r, pmn);
        if (second && !hit) {       // rare: only r < p - n (about 2^-128 of the range for these curves) can have a second candidate
            add_n<L>(rn, r, nn);
            E rne;
            F::from_limbs(rne, rn);
            F::mul(t, rne, zz);
            hit = F::eq(t, R.X);
        }
        return valid && okr && hit;
    }
    // DecompressPoint::decompress (k256 affine.rs:184-202, primeorder affine.rs:129-150): x limbs (plain) -> point
    ECB_DEV static bool decompress(typename JJ::A& Q, const u32* x, u32 y_is_odd) {
        E alpha, beta, t;
        bool ok = F::from_limbs(Q.x, x);          // rejects x >= p
        F::sqr(t, Q.x);
        F::mul(alpha, t, Q.x);
        if constexpr (!C::A_IS_ZERO) {
            F::dbl(t, Q.x); F::add(t, t, Q.x);
            F::sub(alpha, alpha, t);
        }
        G::const_b(t);
        F::add(alpha, alpha, t);
        F::sqrt_candidate(beta, alpha);
        F::sqr(t, beta);
        ok = ok && F::eq(t, alpha);
        E nb;
        F::neg(nb, beta);
        u32 odd = F::is_odd(beta) ? 1u : 0u;
        F::select(Q.y, odd == (y_is_odd & 1u), beta, nb);
        return ok;
    }
    // DecompactPoint::decompact: secp256k1 takes the even root (k256 affine.rs:204-211, the BIP340 convention); the
    // primeorder curves take the root with the smaller integer y (primeorder/src/affine.rs:148-156 with to_compact :66-77)
    ECB_DEV static bool decompact(typename JJ::A& Q, const u32* x) {
        bool ok = decompress(Q, x, 0u);
        if constexpr (!C::A_IS_ZERO) {
            E ny;
            F::neg(ny, Q.y);
            u32 a[L], b[L];
            F::to_limbs(a, Q.y);
            F::to_limbs(b, ny);
            F::select(Q.y, geq_n<L>(b, a), Q.y, ny);      // keep y when y <= p - y
        }
        return ok;
    }
    // q: x||y (ECDSA / SM2DSA), x only (Schnorr); aux: recovery ids (VM_RECOVER), zin: e bytes (VM_SM2DSA target).
    // VM_SCHNORR / VM_RECOVER write the result point to proj_out (identity for rejected rows) and the validity so far to ok_out.
    // NS / stab / sstride: window-table entries kept in shared memory (secp256k1 GLV path, jac.cuh WinTab); NS = 0: none
    template <int NS = 0>
    ECB_DEV static void body_verify_main(int tid, int n, int mode, const u8* q, const u8* rs, const u8* zin, const u8* aux,
                                         const u32* scratch, const u32* gbig, int gw, u8* ok_out, u32* proj_out,
                                         u32* stab = nullptr, int sstride = 0, const u32* wtab = nullptr) {
        if (tid >= n) return;
        const u32* rec = scratch + (size_t)tid * PREP_WORDS;
        typename JJ::A Q;
        bool valid;
        if (mode == VM_SCHNORR) {
            u32 x[L];
            load_be<L>(x, q + (size_t)tid * FB);
            valid = decompress(Q, x, 0u);             // lift_x: the even root (k256/src/schnorr/verifying.rs:35-45)
        } else if (mode == VM_RECOVER) {
            const u32 id = aux[tid];
            u32 x[L], nn[L], r[L];
            load_be<L>(r, rs + (size_t)tid * 2 * FB);
            valid = id < 4u;
            copy_n<L>(x, r);
            if (id & 2u) {                            // x(R) was reduced: restore r + n, reject on overflow (>= p fails in decompress)
                ECB_UNROLL
                for (int i = 0; i < L; i++) nn[i] = C::n(i);
                valid = (add_n<L>(x, r, nn) == 0) && valid;
            }
            valid = decompress(Q, x, id & 1u) && valid;
        } else {
            Aff Qa;
            valid = G::load_affine(Qa, q + (size_t)tid * 2 * FB);
            Q.x = Qa.x; Q.y = Qa.y;
        }
        if (!valid) { Aff g; G::generator(g); Q.x = g.x; Q.y = g.y; }
        typename JJ::J acc;
        if constexpr (C::A_IS_ZERO) {
            K256Glv::Split sp;
            ECB_UNROLL
            for (int l = 0; l < 5; l++) { sp.a1[l] = rec[8 + l]; sp.a2[l] = rec[13 + l]; }
            u32 fl = rec[18];
            valid = valid && (fl & 1u);
            sp.neg1 = (u32)0 - ((fl >> 1) & 1u);
            sp.neg2 = (u32)0 - ((fl >> 2) & 1u);
            K256Fast::template mul_glv<NS>(acc, Q, sp, stab, sstride);
        } else {
            u32 u2[L];
            ECB_UNROLL
            for (int l = 0; l < L; l++) u2[l] = rec[L + l];
            valid = valid && (rec[2 * L] & 1u);
            // ECDSA / SM2DSA: the key's window table was made affine for the whole batch by k_wintab; recovery derives its
            // point inside this kernel and keeps the per-thread Jacobian table
            // (a run-time switch on purpose: with the Jacobian loop compiled out, ptxas allocated the P-384 squarer so badly
            // - predicate spills, 367 instead of 272 instructions - that the kernel lost 10 %; P-256 was 1 % slower too)
            if (wtab && (mode == VM_ECDSA || mode == VM_SM2DSA)) JJ::mul_window_affine(acc, wtab + (size_t)tid * 16 * L, u2);
            else JJ::mul_window_signed(acc, Q, u2);
        }
        // u1 and r are fetched only now: nothing of them stays live across the window loop
        u32 u1[L], r[L];
        ECB_UNROLL
        for (int l = 0; l < L; l++) u1[l] = rec[l];
        JJ::add_fixed_base(acc, u1, gbig, gw);
        if (mode == VM_ECDSA || mode == VM_SM2DSA) load_be<L>(r, rs + (size_t)tid * 2 * FB);
        if (mode == VM_ECDSA) {
            ok_out[tid] = finish_verify_jac(acc, r, valid) ? 1 : 0;
        } else if (mode == VM_SM2DSA) {
            // target = (r - e) mod n
            u32 e[L];
            G::load_scalar(e, zin + (size_t)tid * FB);
            typename Fn::E a, b, t;
            copy_n<L>(a.v, r); copy_n<L>(b.v, e);
            if (!valid) Fn::set_zero(a);              // r may be out of range on rejected rows
            Fn::sub(t, a, b);
            ok_out[tid] = finish_verify_jac(acc, t.v, valid, true) ? 1 : 0;
        } else {
            Proj o;
            if (valid) JJ::to_proj(o, acc); else G::set_identity(o);
            store_proj(proj_out + (size_t)tid * 3 * L, o);
            ok_out[tid] = valid ? 1 : 0;
        }
    }

    // ------------------------------------------------------------------ SEC1 decoding
    // FromEncodedPoint (k256 affine.rs:241-270, primeorder affine.rs:164-195).  enc: n slots of `stride` bytes.
    // DEC_SEC1: tag 02/03 + x, tag 05 + x (compact) (stride >= 1+FB), tag 04 + x + y (stride >= 1+2FB), all-zero slot =
    // identity; DEC_COMPACT: x only, decompact (even root on secp256k1 = BIP340 keys; smaller y on the primeorder curves).  status: 1 point, 2 identity, 0 invalid; xy zeroed unless 1.
    ECB_DEV static void body_decode(int tid, int n, int mode, const u8* enc, int stride, u8* xy, u8* status) {
        if (tid >= n) return;
        const u8* e = enc + (size_t)tid * stride;
        u8* o = xy + (size_t)tid * 2 * FB;
        typename JJ::A Q;
        u32 st = 0;
        u32 x[L];
        if (mode == DEC_COMPACT) {
            load_be<L>(x, e);
            st = decompact(Q, x) ? 1u : 0u;
        } else {
            const u32 tag = e[0];
            if (tag == 5u && stride >= 1 + FB) {          // sec1::Tag::Compact
                load_be<L>(x, e + 1);
                st = decompact(Q, x) ? 1u : 0u;
            } else if ((tag == 2u || tag == 3u) && stride >= 1 + FB) {
                load_be<L>(x, e + 1);
                st = decompress(Q, x, tag & 1u) ? 1u : 0u;
            } else if (tag == 4u && stride >= 1 + 2 * FB) {
                Aff a;
                st = G::load_affine(a, e + 1) ? 1u : 0u;
                Q.x = a.x; Q.y = a.y;
            } else if (tag == 0u) {
                u32 any = 0;
                for (int b = 1; b < stride; b++) any |= e[b];
                st = any ? 0u : 2u;
            }
        }
        if (st == 1u) {
            u32 t[L];
            F::to_limbs(t, Q.x); store_be<L>(o, t);
            F::to_limbs(t, Q.y); store_be<L>(o + FB, t);
        } else {
            for (int b = 0; b < 2 * FB; b++) o[b] = 0;
        }
        status[tid] = (u8)st;
    }

    // ------------------------------------------------------------------ byte-level epilogues after normalisation
    // FIN_SCHNORR: ok &= !identity && y even && x == r           (a: x||y bytes, inf: identity flags, rs: r||s)
    // FIN_RECOVER: ok &= recovered key != identity                (a: SEC1 slots of `stride` bytes)
    ECB_DEV static void body_finish(int tid, int n, int kind, const u8* a, int stride, const u8* inf, const u8* rs, u8* ok) {
        if (tid >= n) return;
        u32 v = ok[tid];
        if (kind == FIN_SCHNORR) {
            const u8* xy = a + (size_t)tid * 2 * FB;
            const u8* r = rs + (size_t)tid * 2 * FB;
            u32 diff = 0;
            for (int b = 0; b < FB; b++) diff |= (u32)(xy[b] ^ r[b]);
            v = v && !inf[tid] && !(xy[2 * FB - 1] & 1u) && diff == 0;
        } else {
            v = v && a[(size_t)tid * stride] != 0;
        }
        ok[tid] = v ? 1 : 0;
    }

    // ------------------------------------------------------------------ ECDSA signing epilogue
    // hazmat::sign_prehashed (ecdsa 0.16.9) + k256 try_sign_prehashed (k256/src/ecdsa.rs:181-198).  R = k*G arrives as
    // affine limbs from the constant-time fixed-base kernel + normalisation; here r = x(R) mod n,
    // s = k^-1 (z + r d), recid = y_odd | x_reduced << 1, low-s normalisation for k256.  k^-1 comes from Montgomery's trick
    // over the rows of a thread.  Secret-dependent values (d, k, k^-1, the point) never steer a branch or an address:
    // selections are masks, the inversion exponent is public.  ok = 0 (and zeroed outputs) when d or k is 0 or >= n, or r or s is 0.
    ECB_DEV static bool load_secret_scalar(u32* v, const u8* bytes) {   // plain limbs; false (and v = 1) unless 1 <= v < n
        load_be<L>(v, bytes);
        bool ok = scalar_in_range(v);
        u32 one[L];
        zero_n<L>(one); one[0] = 1;
        select_n<L>(v, ok, v, one);
        return ok;
    }
    template <class INV = OwnInv>
    ECB_DEV static void body_sign_finish(int tid, int nthreads, int n, const u8* d, const u8* k, const u8* z, const u32* aff,
                                         u8* rs_out, u8* recid_out, u8* ok_out) {
        typename Fn::E pref[PREP_EPT];
        typename Fn::E acc, inv;
        Fn::set_one(acc);
        int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < PREP_EPT; j++) {
            int i = tid + j * nthreads;
            if (i >= n) break;
            u32 kk[L];
            load_secret_scalar(kk, k + (size_t)i * FB);
            typename Fn::E km;
            Fn::from_limbs(km, kk);
            pref[j] = acc;
            Fn::mul(acc, acc, km);
            cnt++;
        }
        if (!INV::COOPERATIVE && cnt == 0) return;
        INV::template run<Fn>(inv, acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = cnt - 1; j >= 0; j--) {
            int i = tid + j * nthreads;
            u32 kk[L], dd[L], zz[L];
            u32 okm = load_secret_scalar(kk, k + (size_t)i * FB) ? 1u : 0u;       // validity flags are combined as masks, never
            okm &= load_secret_scalar(dd, d + (size_t)i * FB) ? 1u : 0u;          // with && / || (no branch on secret-derived bits)
            G::load_scalar(zz, z + (size_t)i * FB);
            typename Fn::E km, kinv;
            Fn::from_limbs(km, kk);
            Fn::mul(kinv, inv, pref[j]);            // k_i^-1 (Montgomery form)
            Fn::mul(inv, inv, km);
            // r = x(R) mod n, recovery bits
            Aff R;
            load_aff_limbs(R, aff + (size_t)i * 2 * L);
            u32 x[L], y[L], xr[L], nn[L];
            F::to_limbs(x, R.x);
            F::to_limbs(y, R.y);
            ECB_UNROLL
            for (int l = 0; l < L; l++) nn[l] = C::n(l);
            u32 bw = sub_n<L>(xr, x, nn);
            u32 rr[L];
            select_n<L>(rr, bw == 0, xr, x);
            u32 recid = (y[0] & 1u) | ((bw == 0) ? 2u : 0u);
            // s = k^-1 (z + r d)
            typename Fn::E rm, dm, t, zm, sm;
            Fn::from_limbs(rm, rr);                  // Montgomery form of r
            Fn::mul_plain(t.v, dd, rm);              // r*d, plain
            copy_n<L>(zm.v, zz);
            Fn::add(t, t, zm);
            Fn::mul_plain(sm.v, t.v, kinv);          // plain s
            okm &= is_zero_n<L>(rr) ? 0u : 1u;
            okm &= Fn::is_zero(sm) ? 0u : 1u;
            if constexpr (C::LOW_S) {                // normalize_s + parity flip (k256/src/ecdsa.rs:192-196)
                u32 hn[L];
                ECB_UNROLL
                for (int l = 0; l < L; l++) hn[l] = C::half_n(l);
                bool high = !geq_n<L>(hn, sm.v);
                typename Fn::E ns;
                Fn::neg(ns, sm);
                Fn::select(sm, high, ns, sm);
                recid ^= high ? 1u : 0u;
            }
            const u32 m = (u32)0 - okm;
            ECB_UNROLL
            for (int l = 0; l < L; l++) { rr[l] &= m; sm.v[l] &= m; }
            store_be<L>(rs_out + (size_t)i * 2 * FB, rr);
            store_be<L>(rs_out + (size_t)i * 2 * FB + FB, sm.v);
            recid_out[i] = (u8)(recid & m);
            ok_out[i] = (u8)(m & 1u);
        }
    }

    // ------------------------------------------------------------------ per-row affine window tables (primeorder public-input path)
    // wtab[row][j] = (j+1) * Q_row as AFFINE field-internal limbs (x|y, 2L words), j = 0..7, for the window loop of
    // Jac::mul_window_affine.  Pass 1 builds the Jacobian multiples 2Q..8Q of every row a thread owns (rows tid, tid +
    // nthreads, ...; X, Y parked in the table slots, Z_2..Z_8 in zbuf) and multiplies the rows' Z-products together;
    // one Fermat inversion per thread; pass 2 unwinds Montgomery's trick at both levels (across the rows, then inside
    // a row) and rewrites the slots as (X/Z^2, Y/Z^3).  An invalid or identity point is replaced by the generator: its
    // row is rejected (or answered with the identity) by the consumer, the table only has to stay invertible.  On these
    // curves (cofactor 1, prime order) no multiple 2Q..8Q of a valid Q is the identity, so no Z is zero.
    //   pts: n x 2FB bytes (x||y) or aff_limbs: n x 2L field-internal limbs;  zbuf: n x 7L words of scratch.
#ifndef ECB_WT_EPT
#define ECB_WT_EPT 16
#endif
    static constexpr int WT_EPT = ECB_WT_EPT;
    template <class INV = OwnInv>
    ECB_DEV static void body_wintab(int tid, int nthreads, int n, const u8* pts, const u32* aff_limbs, u32* wtab, u32* zbuf) {
        E pref[WT_EPT];
        E run;
        F::set_one(run);
        int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < WT_EPT; j++) {
            const int i = tid + j * nthreads;
            if (i >= n) break;
            Aff a;
            bool ok;
            if (aff_limbs) { load_aff_limbs(a, aff_limbs + (size_t)i * 2 * L); ok = G::on_curve(a); }
            else ok = G::load_affine(a, pts + (size_t)i * 2 * FB);
            if (!ok) G::generator(a);
            typename JJ::A Q;
            Q.x = a.x; Q.y = a.y;
            u32* row = wtab + (size_t)i * 16 * L;
            u32* zr = zbuf + (size_t)i * 7 * L;
            store_entry(row, Q.x, Q.y);
            typename JJ::J d2, t, d4;
            E prod;
            JJ::dbl_affine(d2, Q);                 store_entry(row + 2 * L, d2.X, d2.Y);  store_fe(zr, d2.Z);          prod = d2.Z;
            JJ::madd(t, d2, Q, nullptr);           store_entry(row + 4 * L, t.X, t.Y);    store_fe(zr + L, t.Z);       F::mul(prod, prod, t.Z);
            JJ::dbl(d4, d2);                       store_entry(row + 6 * L, d4.X, d4.Y);  store_fe(zr + 2 * L, d4.Z);  F::mul(prod, prod, d4.Z);
            JJ::dbl(d2, t);                        /* 6Q */
            JJ::madd(t, d4, Q, nullptr);           store_entry(row + 8 * L, t.X, t.Y);    store_fe(zr + 3 * L, t.Z);   F::mul(prod, prod, t.Z);
            store_entry(row + 10 * L, d2.X, d2.Y); store_fe(zr + 4 * L, d2.Z);            F::mul(prod, prod, d2.Z);
            JJ::madd(t, d2, Q, nullptr);           store_entry(row + 12 * L, t.X, t.Y);   store_fe(zr + 5 * L, t.Z);   F::mul(prod, prod, t.Z);
            JJ::dbl(d4, d4);                       store_entry(row + 14 * L, d4.X, d4.Y); store_fe(zr + 6 * L, d4.Z);  F::mul(prod, prod, d4.Z);
            pref[j] = run;
            F::mul(run, run, prod);
            cnt++;
        }
        if (!INV::COOPERATIVE && cnt == 0) return;
        E inv;
        INV::template run<F>(inv, run);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = cnt - 1; j >= 0; j--) {
            const int i = tid + j * nthreads;
            u32* row = wtab + (size_t)i * 16 * L;
            const u32* zr = zbuf + (size_t)i * 7 * L;
            E z[7], c[7];
            ECB_UNROLL
            for (int k = 0; k < 7; k++) load_fe(z[k], zr + k * L);
            c[0] = z[0];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 1; k < 7; k++) F::mul(c[k], c[k - 1], z[k]);      // c[6] = this row's Z-product
            E u;
            F::mul(u, inv, pref[j]);                                        // 1 / c[6]
            F::mul(inv, inv, c[6]);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 6; k >= 0; k--) {
                E zi;
                if (k > 0) { F::mul(zi, u, c[k - 1]); F::mul(u, u, z[k]); } else zi = u;   // zi = 1 / z[k]
                E zi2, zi3, x, y;
                u32* e = row + (size_t)(k + 1) * 2 * L;
                load_fe(x, e); load_fe(y, e + L);
                F::sqr(zi2, zi);
                F::mul(zi3, zi2, zi);
                F::mul(x, x, zi2);
                F::mul(y, y, zi3);
                store_entry(e, x, y);
            }
        }
    }
    ECB_DEV static void store_fe(u32* dst, const E& a) { st_words<L>(dst, a.v); }
    ECB_DEV static void load_fe(E& a, const u32* src) { ld_words<L>(a.v, src); }
    ECB_DEV static void store_entry(u32* dst, const E& x, const E& y) { store_fe(dst, x); store_fe(dst + L, y); }


    // ------------------------------------------------------------------ per-key window tables (keys that repeat inside one call)
    // Public-input verification spends 84 % of its multiplications on u2*Q: 128 (256, 384) doublings and ~62 window additions
    // per row.  When the same public key signs many rows of a call - validators, servers, the 2^16 keys of BASELINE's configs
    // 2 / 3 - the doublings depend on the KEY only.  The verify pipeline therefore groups the rows of a call by key (a hash table
    // in HBM, byte-exact comparison), builds for every distinct key ONCE the full-position table
    //     T[g][w][v-1] = v * 2^(W w) * Q_g,   v = 1..2^(W-1),  w = 0..KT_WINDOWS-1      (affine, field-internal limbs, 2L words per entry)
    // and verifies each row with KT_WINDOWS (x2 with the GLV split) mixed additions gathered from HBM and NO doublings:
    // secp256k1 ~920 instead of ~1890 field multiplications per row, P-256 ~890 instead of ~2900.  Tables cost ~5 k
    // multiplications per key (2.6 verifications), so the path is taken only when the call reuses keys (abi.cu: at least 4 rows
    // per key); otherwise the per-row path above runs.  Nothing is kept between calls: every call groups and builds afresh.
    // The reference verifies row by row (k256/src/ecdsa.rs:200-209 -> lincomb, mul.rs:342-393) and recomputes u2*Q every time.
    // window width W (ECB_KT_W, bits): signed digits d_w in [-2^(W-1), 2^(W-1)) below the top window, which is unsigned and
    // absorbs the carry of the recoding; KT_E = 2^(W-1) entries per window.  Wider windows trade table construction (KT_E - 1
    // affine operations per window, once per key) for fewer additions per row: measured on the B200 at 2^22 rows / 2^16 keys, see DESIGN.md.
    // Measured at 2^22 rows / 2^16 keys (64 rows per key).  With the first k_kt_fill (32-bit accesses, 7.2 ms per 2^16 secp256k1
    // keys; profiles/r02_ab_keytab_window_width.txt): secp256k1 88.6 (W = 4) / 96.8 (5) / 94.1 (6) M verifies/s, P-256 83.2 / 83.8 /
    // 75.0; at 2^20 rows (16 rows per key) P-384 15.6 / 13.9 / 10.4, SM2 45.1 / 38.6 / 29.2.  With 128-bit accesses the fill costs a
    // third (2.5 ms) and one more bit pays: secp256k1 111.9 (5) / 118.7 (6), P-256 97.0 (4) / 107.3 (5)
    // (profiles/r02_ab_vector_access.txt).  secp256k1 recodes two 128-bit halves, so a wider window removes twice the additions per
    // table entry added: W = 6 (22 windows of 32 entries per half, 45 KB per key; a wide table costs ~8 rows' worth of the per-row
    // path); P-256 takes W = 5 (52 windows of 16 entries, 53 KB per key); P-384 / SM2 and the small curves keep W = 4 (their
    // BASELINE-shaped case has 16 rows per key).  -DECB_KT_W=n forces one width for every curve.
    // The widths above are for calls with many rows per key; a table is paid per KEY, so calls with 4 .. 20 rows per key take
    // the narrow width W = 4 instead (the launchers instantiate the three table kernels for both): per 2^16 keys the secp256k1 fill
    // costs ~1.3 / 2.5 / 5.0 ms at W = 4 / 5 / 6, and measured the widths cross at ~16 rows per key (abi.cu KT_WIDE_REUSE).
#ifdef ECB_KT_W
    static constexpr int KT_W_DEFAULT = ECB_KT_W;
#else
    static constexpr int KT_W_DEFAULT = C::A_IS_ZERO ? 6 : (C::ID == 1 ? 5 : 4);
#endif
    static constexpr int KT_W = KTW ? KTW : KT_W_DEFAULT;
    static constexpr int KT_E = 1 << (KT_W - 1);                          // entries per window: multiples 1 .. 2^(W-1)
    static constexpr int KT_BITS = C::A_IS_ZERO ? 128 : 32 * L;           // bits of the recoded value: a GLV half / a full scalar
    static constexpr int KT_WINDOWS = KT_BITS / KT_W + 1;                 // signed windows below bit W*floor(bits/W), then the top window
    static constexpr int KT_KEY_WORDS = KT_WINDOWS * KT_E * 2 * L;        // u32 words of one key's table
    // v <= 2^BITS - 1 and bias <= 2^T - 1 (T = W * (KT_WINDOWS - 1)) give a top window of at most 2^(BITS - T)
    static_assert((1 << (KT_BITS - KT_W * (KT_WINDOWS - 1))) + (C::A_IS_ZERO ? 1 : 0) <= KT_E, "top window must fit the table");
    static constexpr int KBW = 2 * FB / 4;                                // u32 words of one key as it arrives (x||y bytes)
    static constexpr int KT_EMPTY = -1, KT_OVERFLOW = -2, KT_TAG = 0x40000000;

    ECB_DEV static u32 key_hash(const u32* k) {
        u32 h = k[0] * 0x9E3779B1u ^ k[1] * 0x85EBCA77u ^ k[2] * 0xC2B2AE3Du ^ k[3] * 0x27D4EB2Fu ^ k[KBW - 1] * 0x165667B1u ^ k[KBW / 2] * 0x9E3779B9u;
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
        return h;
    }
    ECB_DEV static bool key_eq(const u32* a, const u32* b) {
        u32 d = 0;
        ECB_UNROLL
        for (int i = 0; i < KBW; i++) d |= a[i] ^ b[i];
        return d == 0;
    }
    // (1) look every row's key up among the groups numbered by earlier chunks of this call (read only): gid = group or -1
    ECB_DEV static void body_kt_lookup(int tid, int n, const u32* q32, const int* htab, u32 hmask, const u32* gkeys, int* gid) {
        if (tid >= n) return;
        const u32* key = q32 + (size_t)tid * KBW;
        u32 slot = key_hash(key) & hmask;
        int g = -1;
        for (;;) {
            const int v = htab[slot];
            if (v == KT_EMPTY) break;
            if (v >= 0 && key_eq(gkeys + (size_t)v * KBW, key)) { g = v; break; }
            slot = (slot + 1) & hmask;
        }
        gid[tid] = g;
    }
    // (2) rows with an unseen key: the first to claim a slot represents the key, the others find it by comparing key bytes
    ECB_DEV static void body_kt_insert(int tid, int n, const u32* q32, int* htab, u32 hmask, const int* gid, int* rep, int* rep_slot) {
        if (tid >= n || gid[tid] >= 0) return;
        const u32* key = q32 + (size_t)tid * KBW;
        u32 slot = key_hash(key) & hmask;
        for (;;) {
            const int old = atomic_cas_i32(htab + slot, KT_EMPTY, KT_TAG | tid);
            if (old == KT_EMPTY) { rep[tid] = tid; rep_slot[tid] = (int)slot; return; }
            if (old >= KT_TAG) {                       // a row of THIS chunk (numbered groups are < KT_TAG and were ruled out by the lookup)
                const int r2 = old & ~KT_TAG;
                if (key_eq(q32 + (size_t)r2 * KBW, key)) { rep[tid] = r2; return; }
            }
            slot = (slot + 1) & hmask;
        }
    }
    // (3) representatives take the next group number, publish it in their slot and store the key bytes for later chunks
    ECB_DEV static void body_kt_number(int tid, int n, const u32* q32, int* htab, const int* gid, const int* rep, const int* rep_slot,
                                       int* counter, int cap, u32* gkeys, int* newgid) {
        if (tid >= n || gid[tid] >= 0 || rep[tid] != tid) return;
        const int g = atomic_add_i32(counter, 1);
        newgid[tid] = g;
        if (g < cap) {
            const u32* key = q32 + (size_t)tid * KBW;
            ECB_UNROLL
            for (int i = 0; i < KBW; i++) gkeys[(size_t)g * KBW + i] = key[i];
            htab[rep_slot[tid]] = g;
        } else {
            htab[rep_slot[tid]] = KT_OVERFLOW;         // more distinct keys than tables: the host falls back to the per-row path
        }
    }
    // (4) every other row of a new group copies its representative's number
    ECB_DEV static void body_kt_assign(int tid, int n, int* gid, const int* rep, const int* newgid) {
        if (tid >= n || gid[tid] >= 0) return;
        gid[tid] = newgid[rep[tid]];
    }
    // table construction, step 1: B_w = 16^w * Q for the keys of groups [g0, g0 + cnt) as homogeneous projective limbs (normalised
    // next by the shared kernel); kvalid[g] = the key decodes to a point on the curve (else the generator stands in and every row
    // of the group is rejected - ecdsa::VerifyingKey cannot hold such a key)
    ECB_DEV static void body_kt_base(int tid, int g0, int cnt, const u32* gkeys, u32* proj, u8* kvalid) {
        if (tid >= cnt) return;
        const int g = g0 + tid;
        Aff a;
        const bool ok = G::load_affine(a, reinterpret_cast<const u8*>(gkeys + (size_t)g * KBW));
        if (!ok) G::generator(a);
        kvalid[g] = ok ? 1 : 0;
        typename JJ::J p;
        typename JJ::A qa;
        qa.x = a.x; qa.y = a.y;
        JJ::from_affine(p, qa);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int w = 0; w < KT_WINDOWS; w++) {
            Proj o;
            JJ::to_proj(o, p);
            store_proj(proj + ((size_t)tid * KT_WINDOWS + w) * 3 * L, o);
            if (w + 1 < KT_WINDOWS) JJ::dbl_n(p, KT_W);
        }
    }
    // table construction, step 2: entries 2..8 of every (key, window) item by AFFINE additions / doublings with shared
    // inversions - Montgomery's trick over the items a thread owns, continued across the CTA by INV (6-7 multiplications per
    // entry instead of 11-16 for Jacobian additions plus a normalisation).  The seven results form a tree of depth three,
    //     round 0: 2B = 2(B)          round 1: 3B = 2B + B, 4B = 2(2B)          round 2: 5B = 4B + B, 6B = 2(3B), 7B = 4B + 3B, 8B = 2(4B)  ...
    // so W - 1 inversion rounds serve all 2^(W-1) - 1 of them (the first version ran seven sequential rounds: its CTA-wide inversion
    // chains, 7 x ~270 dependent multiplications on one thread, were 60 % of the kernel's time - ncu, profiles/).  The prefix
    // products of the trick are parked in the destination entry's x slot, which is free until the result is written: no
    // per-thread scratch array.  No exceptional case can occur: B has prime order n > 16, so no operand pair is equal or
    // opposite and no y is zero.
#ifndef ECB_KT_EPT
#define ECB_KT_EPT 32
#endif
    static constexpr int KT_EPT = ECB_KT_EPT;
    // operation k of round r (k < 2^r) makes entry v = 2^r + 1 + k: an even v is the double of v / 2 (made in round r - 1), an
    // odd v is 2^r + (v - 2^r) - both made earlier, and never equal or opposite
    ECB_DEV static void kt_op(int r, int k, int& dst, int& a, int& b) {      // b == a: doubling
        dst = (1 << r) + 1 + k;
        if ((dst & 1) == 0) { a = dst >> 1; b = a; }
        else { a = 1 << r; b = dst - a; }
    }
    // L2 prefetch of the part of an item round r touches (entries 1 .. 2^(r+1)), requested while the previous item is being
    // computed: a thread's items lie nthreads KB apart in a table far larger than L2, so every first touch is a DRAM round trip
    // (ncu: long_scoreboard 3.1 per issue at four resident CTAs).  MEASURED SLOWER on the B200 and off by default: k_kt_fill 4.35 ->
    // 4.94 ms per 2^16 secp256k1 keys (profiles/r02_ab_prefetch.txt) - the prefetched lines evict each other before they are used
    // (1 MB in flight per SM).  ECB_KT_PREFETCH_MAIN (next window's entries in k_verify_keytab): secp256k1 -0.5 %, P-256 -0.7 %,
    // P-384 +0.8 %; off as well.
#ifndef ECB_KT_PREFETCH
#define ECB_KT_PREFETCH 0
#endif
#ifndef ECB_KT_PREFETCH_MAIN
#define ECB_KT_PREFETCH_MAIN 0
#endif
    ECB_DEV static void kt_prefetch_item(const u32* ent, int r) {
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU) && ECB_KT_PREFETCH
        const int bytes = (2 << r) * 2 * L * 4;
        for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ent) + o));
#else
        (void)ent; (void)r;
#endif
    }
    template <class INV = OwnInv>
    ECB_DEV static void body_kt_fill(int tid, int nthreads, int items, u32* tab) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int r = 0; r < KT_W - 1; r++) {
            const int nops = 1 << r;
            E acc;
            F::set_one(acc);
            int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int j = 0; j < KT_EPT; j++) {
                const int i = tid + j * nthreads;
                if (i >= items) break;
                u32* ent = tab + (size_t)i * KT_E * 2 * L;
                if (j + 1 < KT_EPT && i + nthreads < items) kt_prefetch_item(ent + (size_t)nthreads * KT_E * 2 * L, r);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int k = 0; k < nops; k++) {
                    int dst, a, b;
                    kt_op(r, k, dst, a, b);
                    E d;
                    if (a == b) { E ya; load_fe(ya, ent + (size_t)(a - 1) * 2 * L + L); F::dbl(d, ya); }
                    else { E xa, xb; load_fe(xa, ent + (size_t)(a - 1) * 2 * L); load_fe(xb, ent + (size_t)(b - 1) * 2 * L); F::sub(d, xa, xb); }
                    store_fe(ent + (size_t)(dst - 1) * 2 * L, acc);           // prefix product, parked in the result's own slot
                    F::mul(acc, acc, d);
                }
                cnt++;
            }
            E inv;
            if (INV::COOPERATIVE || cnt) INV::template run<F>(inv, acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int j = cnt - 1; j >= 0; j--) {
                const int i = tid + j * nthreads;
                u32* ent = tab + (size_t)i * KT_E * 2 * L;
                if (j > 0) kt_prefetch_item(ent - (size_t)nthreads * KT_E * 2 * L, r);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int k = nops - 1; k >= 0; k--) {
                    int dst, a, b;
                    kt_op(r, k, dst, a, b);
                    E xa, ya, xb, yb, d, dinv, pre, lam, x3, y3, t;
                    load_fe(xa, ent + (size_t)(a - 1) * 2 * L); load_fe(ya, ent + (size_t)(a - 1) * 2 * L + L);
                    load_fe(pre, ent + (size_t)(dst - 1) * 2 * L);
                    if (a == b) {                              // tangent: lambda = (3 x^2 + a) / (2 y)
                        xb = xa; yb = ya;
                        F::dbl(d, ya);
                        F::sqr(t, xa);
                        if constexpr (!C::A_IS_ZERO) { E one; F::set_one(one); F::sub(t, t, one); }
                        F::dbl(lam, t); F::add(lam, lam, t);
                    } else {                                   // chord: lambda = (y_a - y_b) / (x_a - x_b)
                        load_fe(xb, ent + (size_t)(b - 1) * 2 * L); load_fe(yb, ent + (size_t)(b - 1) * 2 * L + L);
                        F::sub(d, xa, xb);
                        F::sub(lam, ya, yb);
                    }
                    F::mul(dinv, inv, pre);
                    F::mul(inv, inv, d);
                    F::mul(lam, lam, dinv);
                    F::sqr(x3, lam); F::sub(x3, x3, xa); F::sub(x3, x3, xb);
                    F::sub(t, xa, x3); F::mul(y3, lam, t); F::sub(y3, y3, ya);
                    store_entry(ent + (size_t)(dst - 1) * 2 * L, x3, y3);
                }
            }
        }
    }
    // signed radix-2^W recoding of a KT_BITS-bit value v (NW words, in place, two spare words above): v + bias with one bit
    // 2^(W i + W - 1) per signed window i; window i then holds d_i + 2^(W-1), the top window d_top >= 0 including the carry
    ECB_DEV static void kt_recode(u32* v) {
        constexpr int NW = KT_BITS / 32;
        u32 c = 0;
        ECB_UNROLL
        for (int j = 0; j < NW; j++) {
            u32 bias = 0;
            ECB_UNROLL
            for (int i = 0; i < KT_WINDOWS - 1; i++) {
                const int bit = KT_W * i + KT_W - 1;
                if (bit >= 32 * j && bit < 32 * j + 32) bias |= 1u << (bit - 32 * j);
            }
            const u32 t = v[j] + bias;
            const u32 c1 = t < bias ? 1u : 0u;
            v[j] = t + c;
            c = c1 | (v[j] < t ? 1u : 0u);
        }
        v[NW] += c;                                        // v[NW] holds bit KT_BITS of the input if it has one (0 otherwise)
        v[NW + 1] = 0;
    }
    ECB_DEV static void kt_digit(const u32* v, int w, u32& mag, u32& neg) {
        const int bit = KT_W * w;
        const u64 two = ((u64)v[(bit >> 5) + 1] << 32) | v[bit >> 5];
        const u32 raw = (u32)(two >> (bit & 31));
        if (w == KT_WINDOWS - 1) { mag = raw; neg = 0; return; }           // unsigned top window (everything above is zero)
        const int d = (int)(raw & ((1u << KT_W) - 1u)) - (1 << (KT_W - 1));
        neg = (u32)(d >> 31);
        mag = (u32)((d ^ (int)neg) - (int)neg);
    }
    // |r| + 0x8888...8 (five words, K256Glv::bias) -> |r| (five words: bit 128 stays where it is, should a half ever reach 2^128)
    ECB_DEV static void kt_unbias16(u32* o, const u32* a) {
        o[0] = sub_cc(a[0], 0x88888888u);
        o[1] = subc_cc(a[1], 0x88888888u);
        o[2] = subc_cc(a[2], 0x88888888u);
        o[3] = subc_cc(a[3], 0x88888888u);
        o[4] = subc(a[4], 0u);
    }
    // the verify main kernel on per-key tables: no doublings, KT_WINDOWS (x2) gathered mixed additions, then u1*G and the
    // inversion-free comparison exactly as body_verify_main
    ECB_DEV static void body_verify_keytab(int tid, int n, int mode, const u8* rs, const u8* zin, const u32* scratch, const int* gid, const u8* kvalid,
                                           const u32* tab, const u32* gbig, int gw, u8* ok_out) {
        if (tid >= n) return;
        const u32* rec = scratch + (size_t)tid * PREP_WORDS;
        const int g = gid[tid];
        bool valid = kvalid[g] != 0;
        const u32* tk = tab + (size_t)g * KT_KEY_WORDS;
        typename JJ::J acc;
        JJ::set_inf(acc);
        constexpr int NW = KT_BITS / 32;                    // words of the recoded value
        if constexpr (C::A_IS_ZERO) {
            const u32 fl = rec[18];
            valid = valid && (fl & 1u);
            const u32 neg1 = (u32)0 - ((fl >> 1) & 1u), neg2 = (u32)0 - ((fl >> 2) & 1u);
            // the prep kernel stores |r1|, |r2| biased for radix 16 (K256Glv::bias): take the bias off, recode for this width
            u32 k1[NW + 2], k2[NW + 2];
            kt_unbias16(k1, rec + 8);
            kt_unbias16(k2, rec + 13);
            kt_recode(k1);
            kt_recode(k2);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int w = 0; w < KT_WINDOWS; w++) {
                u32 mag, neg;
#if defined(__CUDA_ARCH__) && ECB_KT_PREFETCH_MAIN
                if (w + 1 < KT_WINDOWS) {                  // request the two entries of the next window while this one is added
                    kt_digit(k1, w + 1, mag, neg);
                    if (mag) asm volatile("prefetch.global.L2 [%0];" ::"l"(tk + ((size_t)(w + 1) * KT_E + mag - 1) * 2 * L));
                    kt_digit(k2, w + 1, mag, neg);
                    if (mag) asm volatile("prefetch.global.L2 [%0];" ::"l"(tk + ((size_t)(w + 1) * KT_E + mag - 1) * 2 * L));
                }
#endif
                kt_digit(k1, w, mag, neg);
                if (mag) {
                    typename JJ::A e;
                    JJ::load_entry(e, tk + ((size_t)w * KT_E + mag - 1) * 2 * L);
                    JJ::cneg_y(e, neg ^ neg1);
                    JJ::madd(acc, acc, e, nullptr);
                }
                kt_digit(k2, w, mag, neg);
                if (mag) {
                    typename JJ::A e;
                    JJ::load_entry(e, tk + ((size_t)w * KT_E + mag - 1) * 2 * L);
                    // beta is rebuilt from immediates here (the empty volatile asm pins that): hoisted out of the loop it held
                    // eight registers across every call of the window loop, which ptxas paid for in spills
                    E beta;
                    ECB_UNROLL
                    for (int l = 0; l < 8; l++) {
                        u32 c = CurveK256::beta(l);
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
                        asm volatile("" : "+r"(c));
#endif
                        beta.v[l] = c;
                    }
                    F::mul(e.x, e.x, beta);                 // lambda * (x, y) = (beta x, y)
                    JJ::cneg_y(e, neg ^ neg2);
                    JJ::madd(acc, acc, e, nullptr);
                }
            }
        } else {
            u32 kb[NW + 2];
            ECB_UNROLL
            for (int l = 0; l < L; l++) kb[l] = rec[L + l];
            kb[L] = 0;
            valid = valid && (rec[2 * L] & 1u);
            kt_recode(kb);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int w = 0; w < KT_WINDOWS; w++) {
                u32 mag, neg;
#if defined(__CUDA_ARCH__) && ECB_KT_PREFETCH_MAIN
                if (w + 1 < KT_WINDOWS) {
                    kt_digit(kb, w + 1, mag, neg);
                    if (mag) asm volatile("prefetch.global.L2 [%0];" ::"l"(tk + ((size_t)(w + 1) * KT_E + mag - 1) * 2 * L));
                }
#endif
                kt_digit(kb, w, mag, neg);
                if (mag) {
                    typename JJ::A e;
                    JJ::load_entry(e, tk + ((size_t)w * KT_E + mag - 1) * 2 * L);
                    JJ::cneg_y(e, neg);
                    JJ::madd(acc, acc, e, nullptr);
                }
            }
        }
        // u1 and r are fetched only now: nothing of them stays live across the window loop
        u32 r[L], u1[L];
        ECB_UNROLL
        for (int l = 0; l < L; l++) u1[l] = rec[l];
        JJ::add_fixed_base(acc, u1, gbig, gw);
        load_be<L>(r, rs + (size_t)tid * 2 * FB);
        if (mode == VM_ECDSA) {
            ok_out[tid] = finish_verify_jac(acc, r, valid) ? 1 : 0;
        } else {                                           // VM_SM2DSA: target = (r - e) mod n, the identity counts as x = 0
            u32 e[L];
            G::load_scalar(e, zin + (size_t)tid * FB);
            typename Fn::E a, b, t;
            copy_n<L>(a.v, r); copy_n<L>(b.v, e);
            if (!valid) Fn::set_zero(a);
            Fn::sub(t, a, b);
            ok_out[tid] = finish_verify_jac(acc, t.v, valid, true) ? 1 : 0;
        }
    }

    // ------------------------------------------------------------------ variable-base k*P, vartime fast path (v2)
    // pts: n x 2FB big-endian affine bytes (aff_limbs == nullptr), or internal affine limbs produced by the
    // normalisation kernel from projective inputs (all-zero entry = identity); output projective limbs
    template <int NS = 0>
    ECB_DEV static void body_mul_var_fast(int tid, int n, const u8* pts, const u32* aff_limbs, const u8* inf, const u8* k, u32* out, u8* invalid,
                                          u32* stab = nullptr, int sstride = 0, const u32* wtab = nullptr) {
        if (tid >= n) return;
        Aff a;
        bool ok, isinf;
        if (aff_limbs) {
            load_aff_limbs(a, aff_limbs + (size_t)tid * 2 * L);
            isinf = F::is_zero(a.x) && F::is_zero(a.y);
            ok = isinf || G::on_curve(a);
        } else {
            ok = G::load_affine(a, pts + (size_t)tid * 2 * FB);
            isinf = inf && inf[tid];
        }
        u32 kk[L];
        G::load_scalar(kk, k + (size_t)tid * FB);
        Proj r;
        if (!ok || isinf || is_zero_n<L>(kk)) {
            G::set_identity(r);
        } else {
            typename JJ::A Q;
            Q.x = a.x; Q.y = a.y;
            typename JJ::J acc;
            if constexpr (C::A_IS_ZERO) {
                K256Glv::Split sp;
                K256Glv::decompose(sp, kk);
                K256Fast::template mul_glv<NS>(acc, Q, sp, stab, sstride);
            } else {
                if (wtab) JJ::mul_window_affine(acc, wtab + (size_t)tid * 16 * L, kk);
                else JJ::mul_window_signed(acc, Q, kk);
            }
            JJ::to_proj(r, acc);
        }
        store_proj(out + (size_t)tid * 3 * L, r);
        if (invalid) invalid[tid] = (ok || isinf) ? 0 : 1;
    }

    // ------------------------------------------------------------------ fixed-base k*G -> projective
    // 8*L + 1 signed radix-16 digits, table tab[i][j] = (j+1) * 16^i * G affine (GEN_WINDOWS x 8 entries), one complete
    // mixed addition per digit and no doublings (the reference spaces 33 tables by 2^8 and pays 4 doublings for k256,
    // k256/src/arithmetic/mul.rs:397-439, and runs the generic window multiplication for the primeorder curves,
    // primeorder/src/projective.rs:422-431; the result point is the same).
    static constexpr int GEN_WINDOWS = 8 * L + 1;

    // ------------------------------------------------------------------ fixed-base k*G, split form (round 2)
    // The same sum  k*G = sum_w d_w * 2^(W w) * G  over signed radix-2^W digits (W = ECB_GEN2_W, 5 by default: 52 windows of
    // 16 entries for a 256-bit curve instead of 65 windows of 8), but
    //   * the windows are cut into two contiguous halves that are summed by two different threads (blockIdx.y of
    //     k_gen_half): twice the threads for the same work - a 2^16-row batch (BASELINE configs[0]) fills 7 resident CTAs
    //     per SM in one wave instead of leaving the chip at 4 warps per scheduler - and each CTA stages only its half of the
    //     table in shared memory (26 KB for secp256k1: seven CTAs per SM fit);
    //   * inside a half the additions are Jacobian mixed additions (8M + 3S, Jac::madd_ct) instead of complete projective
    //     ones (11M + 2 m_3b): every entry of window w is v * 2^(W w) * G with v >= 1, while the running sum of the lower
    //     windows of the same half is smaller than 2^(W w) in absolute value (signed digits, |d| <= 2^(W-1)) and, in the upper
    //     half, a multiple of 2^(W NH) just like the entry, so sum = +-entry (mod n) is impossible and the exceptional cases
    //     of the incomplete formulas never arise; "the sum is still the identity" and "the digit is zero" are handled by masks;
    //   * the two partial sums are added with the COMPLETE formula (they may be equal, opposite or the identity) inside
    //     the normalisation kernel (body_normalize, sum_with).
    // Secret-scalar discipline as before (k256/src/arithmetic/mul.rs:424-439 scans its tables the same way): every entry of
    // a window is read and combined by mask, digits never steer a branch or an address.  CT = false (public scalars) indexes
    // the table directly and skips zero digits.
    // Table layout: tab[(w * G2_E + v - 1) * 2L], v = 1..G2_E; the top window is unsigned (bits above W * (G2_WINDOWS - 1)
    // plus the carry of the recoding, at most 2^(32L - W (G2_WINDOWS - 1)) <= G2_E).
#ifndef ECB_GEN2_W
#define ECB_GEN2_W 5
#endif
    static constexpr int G2_W = ECB_GEN2_W;
    static constexpr int G2_E = 1 << (G2_W - 1);
    static constexpr int G2_WINDOWS = (32 * L) / G2_W + 1;
    static constexpr int G2_NH = (G2_WINDOWS + 1) / 2;                 // windows of the lower half; the upper half has the rest
    static_assert((1 << (32 * L - G2_W * (G2_WINDOWS - 1))) <= G2_E, "top window must fit the table");
    // k (L words, < n) -> k + bias in place (L + 2 words), one bias bit 2^(W i + W - 1) per signed window
    ECB_DEV static void g2_recode(u32* v) {
        u32 c = 0;
        ECB_UNROLL
        for (int j = 0; j < L; j++) {
            u32 bias = 0;
            ECB_UNROLL
            for (int i = 0; i < G2_WINDOWS - 1; i++) {
                const int bit = G2_W * i + G2_W - 1;
                if (bit >= 32 * j && bit < 32 * j + 32) bias |= 1u << (bit - 32 * j);
            }
            const u32 t = v[j] + bias;
            const u32 c1 = t < bias ? 1u : 0u;
            v[j] = t + c;
            c = c1 | (v[j] < t ? 1u : 0u);
        }
        v[L] = c;
        v[L + 1] = 0;
    }
    ECB_DEV static void g2_digit(const u32* v, int w, u32& mag, u32& neg) {
        const int bit = G2_W * w;
        const u64 two = ((u64)v[(bit >> 5) + 1] << 32) | v[bit >> 5];
        const u32 raw = (u32)(two >> (bit & 31));
        const int d = (int)(raw & ((1u << G2_W) - 1u)) - (1 << (G2_W - 1));
        const u32 sneg = (u32)(d >> 31);
        const u32 smag = (u32)((d ^ (int)sneg) - (int)sneg);
        const u32 top = (u32)0 - (u32)(w == G2_WINDOWS - 1);            // public: the window index
        mag = (raw & top) | (smag & ~top);
        neg = sneg & ~top;
    }
    ECB_DEV static void g2_load_entry(typename JJ::A& e, const u32* p) {
#if defined(__CUDA_ARCH__)
        if constexpr (L % 4 == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
            ECB_UNROLL
            for (int k = 0; k < L / 4; k++) {
                const uint4 a = q[k], b = q[L / 4 + k];
                e.x.v[4 * k] = a.x; e.x.v[4 * k + 1] = a.y; e.x.v[4 * k + 2] = a.z; e.x.v[4 * k + 3] = a.w;
                e.y.v[4 * k] = b.x; e.y.v[4 * k + 1] = b.y; e.y.v[4 * k + 2] = b.z; e.y.v[4 * k + 3] = b.w;
            }
            return;
        }
#endif
        ECB_UNROLL
        for (int l = 0; l < L; l++) { e.x.v[l] = p[l]; e.y.v[l] = p[L + l]; }
    }
    // half: 0 = windows [0, G2_NH), 1 = [G2_NH, G2_WINDOWS); tabh: the entries of THIS half's windows (shared memory on the
    // device); part: 2 x n x 3L limbs, half h of row i at (h * n + i) * 3L as a homogeneous projective point
    // ECB_GEN2_SHFL (device only): the 16 entries of a window are spread over the lanes of the warp (lane j holds entry
    // (j & 15) + 1, a PUBLIC address) and every lane fetches the entry of its own digit with one warp shuffle per word - the
    // secret digit selects a source LANE of a register exchange, never a memory address or a branch; 4 LDS.128 + 16 SHFL per
    // window instead of 64 LDS.128 + 256 LOP3 for the masked scan.  Every lane of the warp must take part, so rows past the
    // end of the batch are clamped to the last row instead of returning early (their result is not stored).
    // Measured (profiles/r02_ab_vector_access.txt): k*G at 2^16 scalars 0.579 -> 0.531 ms, 2^20 134.9 -> 149.8 M/s, signing 138.4 ->
    // 153.7 M/s; the dynamic audit (scripts/ct_audit.sh on this build: instruction, branch, local / global sector and shared
    // wavefront counters of five secret patterns) finds no launch that depends on the secrets.  -DECB_GEN2_SHFL=0 builds the scan.
#ifndef ECB_GEN2_SHFL
#define ECB_GEN2_SHFL 1
#endif
    template <bool CT> ECB_DEV static void body_gen_half(int tid, int n, int half, const u8* k, const u32* tabh, u32* part) {
#if defined(__CUDA_ARCH__) && ECB_GEN2_SHFL
        const bool live = tid < n;
        if (!live) tid = n - 1;
        const int lane16 = (int)(threadIdx.x & 15u);
#else
        if (tid >= n) return;
#endif
        u32 kb[L + 2];
        G::load_scalar(kb, k + (size_t)tid * FB);
        g2_recode(kb);
        const int w0 = half * G2_NH;
        const int w1 = half ? G2_WINDOWS : G2_NH;
        typename JJ::J acc;
        JJ::set_inf(acc);
        u32 inf = 0xFFFFFFFFu;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int w = w0; w < w1; w++) {
            u32 mag, neg;
            g2_digit(kb, w, mag, neg);
            const u32* win = tabh + (size_t)(w - w0) * G2_E * 2 * L;
            typename JJ::A g;
            if constexpr (CT) {
#if defined(__CUDA_ARCH__) && ECB_GEN2_SHFL
                static_assert(G2_E <= 16, "one entry per lane of a half warp");
                typename JJ::A mine;
                g2_load_entry(mine, win + (size_t)(lane16 & (G2_E - 1)) * 2 * L);
                const int src = (int)((mag - 1u) & (u32)(G2_E - 1));       // digit 0 reads some entry; the addition is masked out below
                ECB_UNROLL
                for (int l = 0; l < L; l++) {
                    g.x.v[l] = __shfl_sync(0xFFFFFFFFu, mine.x.v[l], src);
                    g.y.v[l] = __shfl_sync(0xFFFFFFFFu, mine.y.v[l], src);
                }
#else
                F::set_zero(g.x); F::set_zero(g.y);
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
                for (u32 j = 1; j <= (u32)G2_E; j++) {
                    typename JJ::A c;
                    g2_load_entry(c, win + (size_t)(j - 1) * 2 * L);
                    const u32 m = (u32)0 - (u32)(j == mag);
                    F::cmov(g.x, c.x, m); F::cmov(g.y, c.y, m);
                }
#endif
                JJ::cneg_y(g, neg);
                inf = JJ::madd_ct(acc, inf, g, (u32)0 - (u32)(mag != 0));
            } else {
                if (mag) {
                    g2_load_entry(g, win + (size_t)(mag - 1) * 2 * L);
                    JJ::cneg_y(g, neg);
                    JJ::madd(acc, acc, g, nullptr);
                    inf = 0;
                }
            }
        }
        // Jacobian -> homogeneous (X Z : Y : Z^3); the identity becomes (0 : 1 : 0) by mask
        Proj o, id;
        E zz;
        F::sqr(zz, acc.Z);
        F::mul(o.X, acc.X, acc.Z);
        o.Y = acc.Y;
        F::mul(o.Z, zz, acc.Z);
        G::set_identity(id);
        if constexpr (!CT) inf = JJ::is_inf(acc) ? 0xFFFFFFFFu : 0u;
        G::cmov(o, id, inf);
#if defined(__CUDA_ARCH__) && ECB_GEN2_SHFL
        if (!live) return;
#endif
        store_proj(part + ((size_t)half * n + tid) * 3 * L, o);
    }
    template <bool CT> ECB_DEV static void body_mul_gen(int tid, int n, const u8* k, const u32* tab, u32* out) {
        if (tid >= n) return;
        u32 kk[L];
        G::load_scalar(kk, k + (size_t)tid * FB);
        Proj r;
        u32 kb[L + 1];
        kb[0] = add_cc(kk[0], 0x88888888u);
        ECB_UNROLL
        for (int i = 1; i < L; i++) kb[i] = addc_cc(kk[i], 0x88888888u);
        kb[L] = addc(0u, 0u);
        G::set_identity(r);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < GEN_WINDOWS; i++) {
            u32 mag, neg;
            if (i == 8 * L) { mag = kb[L]; neg = 0; }
            else {
                int d = (int)((kb[i >> 3] >> ((i & 7) * 4)) & 15u) - 8;
                neg = (u32)(d >> 31);
                mag = (u32)((d ^ (int)neg) - (int)neg);
            }
            const u32* win = tab + (size_t)i * 8 * 2 * L;
            Aff g;
            if constexpr (CT) {
                F::set_zero(g.x); F::set_zero(g.y);
                for (u32 j = 1; j <= 8; j++) {
                    Aff c;
                    load_aff_limbs(c, win + (size_t)(j - 1) * 2 * L);
                    u32 m = (u32)0 - (u32)(j == mag);
                    F::cmov(g.x, c.x, m); F::cmov(g.y, c.y, m);
                }
                E ny;
                F::neg(ny, g.y);
                F::cmov(g.y, ny, neg);
                Proj sum;
                G::add_mixed(sum, r, g);
                G::cmov(r, sum, (u32)0 - (u32)(mag != 0));
            } else {
                if (mag) {
                    load_aff_limbs(g, win + (size_t)(mag - 1) * 2 * L);
                    E ny;
                    F::neg(ny, g.y);
                    F::cmov(g.y, ny, neg);
                    G::add_mixed(r, r, g);
                }
            }
        }
        store_proj(out + (size_t)tid * 3 * L, r);
    }
};

}  // namespace ecb
