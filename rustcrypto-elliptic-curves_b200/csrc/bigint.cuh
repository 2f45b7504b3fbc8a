// bigint.cuh — 32-bit-limb multiprecision primitives for sm_100a.
//
// Everything the field layers need: PTX carry-chain wrappers (add.cc/addc, sub.cc/subc,
// mad.lo.cc/madc.hi.cc — ptxas fuses a {mad.lo.cc, madc.hi.cc} pair on the same operands into one
// IMAD.WIDE.U32(.X), see profiles/), an even/odd-column schoolbook multiplier that keeps two
// independent carry chains in flight, and small helpers (compare, select, byte I/O).
//
// Limb order is little-endian (v[0] least significant), as in the reference's 32-bit backends
// (k256/src/arithmetic/field/field_8x32_risc0.rs, scalar/wide32.rs).
//
// ECB_EMU: when this header is compiled by a plain host compiler for the CPU-only logic tests
// (tests/emu), the PTX wrappers are replaced by a carry-flag emulation.  The product library is
// always built by nvcc for sm_100a and contains no host arithmetic path.
#pragma once
#include <cstdint>

namespace ecb {

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

#if defined(__CUDACC__) && !defined(ECB_EMU)
#define ECB_DEV __device__ __forceinline__
#define ECB_HD __host__ __device__ __forceinline__
#define ECB_UNROLL _Pragma("unroll")
#else
#define ECB_DEV inline
#define ECB_HD inline
#define ECB_UNROLL
struct uint4 { u32 x, y, z, w; };
#endif

// ------------------------------------------------------------------------------------------
// carry-chain primitives
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
ECB_DEV u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 madlo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madloc_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madhi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madhic_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madhic(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 mulhi(u32 a, u32 b) { return __umulhi(a, b); }
ECB_DEV u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }
#else
// host emulation of the PTX condition-code register (tests/emu only)
static thread_local u32 ecb_cf_ = 0;
inline u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + ecb_cf_; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 addc(u32 a, u32 b) { return a + b + ecb_cf_; }
inline u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; ecb_cf_ = (u32)((t >> 32) & 1); return (u32)t; }
inline u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - ecb_cf_; ecb_cf_ = (u32)((t >> 32) & 1); return (u32)t; }
inline u32 subc(u32 a, u32 b) { return a - b - ecb_cf_; }
inline u32 madlo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)((u64)a * b) + c; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 madloc_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)((u64)a * b) + c + ecb_cf_; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 madhi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 madhic_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c + ecb_cf_; ecb_cf_ = (u32)(t >> 32); return (u32)t; }
inline u32 madhic(u32 a, u32 b, u32 c) { return (u32)((((u64)a * b) >> 32) + c + ecb_cf_); }
inline u32 mulhi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 bswap32(u32 x) { return __builtin_bswap32(x); }
#endif

// 32-bit atomics for the key-grouping kernels (host emulation: the logical threads run one after another)
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
ECB_DEV int atomic_cas_i32(int* p, int cmp, int val) { return atomicCAS(p, cmp, val); }
ECB_DEV int atomic_add_i32(int* p, int v) { return atomicAdd(p, v); }
#else
inline int atomic_cas_i32(int* p, int cmp, int val) { int o = *p; if (o == cmp) *p = val; return o; }
inline int atomic_add_i32(int* p, int v) { int o = *p; *p += v; return o; }
#endif

// ------------------------------------------------------------------------------------------
// n-limb helpers (all loops fully unrolled; L is a compile-time constant)

template <int L> ECB_DEV u32 add_n(u32* r, const u32* a, const u32* b) {   // returns carry
    r[0] = add_cc(a[0], b[0]);
    ECB_UNROLL
    for (int i = 1; i < L; i++) r[i] = addc_cc(a[i], b[i]);
    return addc(0, 0);
}
template <int L> ECB_DEV u32 sub_n(u32* r, const u32* a, const u32* b) {   // returns borrow (0/1)
    r[0] = sub_cc(a[0], b[0]);
    ECB_UNROLL
    for (int i = 1; i < L; i++) r[i] = subc_cc(a[i], b[i]);
    return subc(0, 0) & 1;
}
template <int L> ECB_DEV bool is_zero_n(const u32* a) {
    u32 t = a[0];
    ECB_UNROLL
    for (int i = 1; i < L; i++) t |= a[i];
    return t == 0;
}
template <int L> ECB_DEV bool eq_n(const u32* a, const u32* b) {
    u32 t = a[0] ^ b[0];
    ECB_UNROLL
    for (int i = 1; i < L; i++) t |= a[i] ^ b[i];
    return t == 0;
}
// a >= b  (borrow-free subtraction test)
template <int L> ECB_DEV bool geq_n(const u32* a, const u32* b) {
    u32 t[L];
    return sub_n<L>(t, a, b) == 0;
}
template <int L> ECB_DEV void copy_n(u32* r, const u32* a) {
    ECB_UNROLL
    for (int i = 0; i < L; i++) r[i] = a[i];
}
template <int L> ECB_DEV void zero_n(u32* r) {
    ECB_UNROLL
    for (int i = 0; i < L; i++) r[i] = 0;
}
// r = c ? a : b   (branch-free; SEL in SASS)
template <int L> ECB_DEV void select_n(u32* r, bool c, const u32* a, const u32* b) {
    ECB_UNROLL
    for (int i = 0; i < L; i++) r[i] = c ? a[i] : b[i];
}
// Optimisation barrier for masks derived from secrets: after it the compiler no longer knows that the value is 0 or ~0, so
// it cannot turn (a & m) | (r & ~m) back into a select and then PREDICATE THE LOAD of `a` on the secret comparison.  ptxas did
// exactly that to the table scans of round 1 (scripts/ct_audit.sh: the local-load sector and shared-wavefront counts of
// k_mul_var<*, true> / k_mul_gen_smem<*, true> depended on the scalars although no branch did): a scan that only touches the
// selected entry leaks the digit through the memory system, which is what the full scan exists to prevent
// (k256/src/arithmetic/mul.rs:92-127, primeorder/src/projective.rs:127-147 use subtle's ConditionallySelectable for the same reason).
ECB_DEV u32 value_barrier(u32 m) {
#if defined(__CUDA_ARCH__) || defined(__GNUC__)
    asm volatile("" : "+r"(m));
#endif
    return m;
}
// r = mask ? a : r  with an all-ones/zero mask (constant-time table scans)
template <int L> ECB_DEV void cmov_n(u32* r, const u32* a, u32 mask) {
    mask = value_barrier(mask);
    ECB_UNROLL
    for (int i = 0; i < L; i++) r[i] = (a[i] & mask) | (r[i] & ~mask);
}

// ------------------------------------------------------------------------------------------
// r[0..2L) = a * b.  Even/odd-column schoolbook: products whose weight i+j is even accumulate in
// e[], odd-weight products in o[] (o[k] has weight k+1), so every IMAD.WIDE lands on an aligned
// register pair and two carry chains run independently; r = e + (o << 32) at the end.
// A row chain can only carry out of its top pair when an earlier row already accumulated there: the e chain after
// even rows i >= 2, the o chain after odd rows.  The other carry-outs are provably zero (the top pair is fresh: one
// product plus at most a captured carry cannot reach 2^64) and are not captured at all - each capture costs a SEL, an
// IMAD.X and a zeroing IMAD.MOV, two of them on the multiplier pipe.  (Checked by simulation with all-ones operands,
// which maximise every partial sum; the host emulation asserts it on every call.)
#if defined(ECB_EMU)
#define ECB_ASSERT_NO_CARRY() do { if (ecb_cf_ != 0) __builtin_trap(); } while (0)
#else
#define ECB_ASSERT_NO_CARRY() do { } while (0)
#endif
// Odd limb counts (P-224, L = 7): plain column-wise schoolbook with a three-word accumulator.  The aligned-pair scheme
// below needs an even L; this one is about 30 % slower per product and only serves the curve that needs it.
template <int L> ECB_DEV void mul_wide_columns(u32* r, const u32* a, const u32* b) {
    u64 lo = 0;        // low 64 bits of the column sum
    u32 hi = 0;        // overflow beyond 2^64
    ECB_UNROLL
    for (int k = 0; k < 2 * L - 1; k++) {
        ECB_UNROLL
        for (int i = 0; i < L; i++) {
            const int j = k - i;
            if (j < 0 || j >= L) continue;
            const u64 pr = (u64)a[i] * b[j];
            lo += pr;
            hi += lo < pr ? 1u : 0u;
        }
        r[k] = (u32)lo;
        lo = (lo >> 32) | ((u64)hi << 32);
        hi = 0;
    }
    r[2 * L - 1] = (u32)lo;
}
template <int L> ECB_DEV void mul_wide(u32* r, const u32* a, const u32* b);
template <int L> ECB_DEV void sqr_wide(u32* r, const u32* a);
// odd L on the device: zero-extend to L + 1 limbs and run the aligned-pair multiplier (the products with the zero limb
// cost an instruction each but no carry logic); the column-wise form above stays as the host-emulation cross-check
template <int L> ECB_DEV void mul_wide_padded(u32* r, const u32* a, const u32* b) {
    u32 aa[L + 1], bb[L + 1], rr[2 * L + 2];
    ECB_UNROLL
    for (int i = 0; i < L; i++) { aa[i] = a[i]; bb[i] = b[i]; }
    aa[L] = 0; bb[L] = 0;
    mul_wide<L + 1>(rr, aa, bb);
    ECB_UNROLL
    for (int i = 0; i < 2 * L; i++) r[i] = rr[i];
}
template <int L> ECB_DEV void sqr_wide_padded(u32* r, const u32* a) {
    u32 aa[L + 1], rr[2 * L + 2];
    ECB_UNROLL
    for (int i = 0; i < L; i++) aa[i] = a[i];
    aa[L] = 0;
    sqr_wide<L + 1>(rr, aa);
    ECB_UNROLL
    for (int i = 0; i < 2 * L; i++) r[i] = rr[i];
}
template <int L> ECB_DEV void mul_wide(u32* r, const u32* a, const u32* b) {
    if constexpr (L % 2 != 0) {
#if defined(__CUDA_ARCH__)
        mul_wide_padded<L>(r, a, b);
#else
        mul_wide_columns<L>(r, a, b);
#endif
        return;
    }
    u32 e[2 * L], o[2 * L];
    ECB_UNROLL
    for (int i = 0; i < 2 * L; i++) { e[i] = 0; o[i] = 0; }
    ECB_UNROLL
    for (int i = 0; i < L; i++) {
        const u32 bi = b[i];
        if ((i & 1) == 0) {
            ECB_UNROLL
            for (int j = 0; j < L; j += 2) {
                e[i + j] = (j == 0) ? madlo_cc(a[j], bi, e[i + j]) : madloc_cc(a[j], bi, e[i + j]);
                e[i + j + 1] = madhic_cc(a[j], bi, e[i + j + 1]);
            }
            if (i >= 2) e[i + L] = addc(e[i + L], 0); else ECB_ASSERT_NO_CARRY();
            ECB_UNROLL
            for (int j = 1; j < L; j += 2) {
                o[i + j - 1] = (j == 1) ? madlo_cc(a[j], bi, o[i + j - 1]) : madloc_cc(a[j], bi, o[i + j - 1]);
                o[i + j] = madhic_cc(a[j], bi, o[i + j]);
            }
            ECB_ASSERT_NO_CARRY();
        } else {
            ECB_UNROLL
            for (int j = 0; j < L; j += 2) {
                o[i + j - 1] = (j == 0) ? madlo_cc(a[j], bi, o[i + j - 1]) : madloc_cc(a[j], bi, o[i + j - 1]);
                o[i + j] = madhic_cc(a[j], bi, o[i + j]);
            }
            if (i + L - 1 < 2 * L) o[i + L - 1] = addc(o[i + L - 1], 0);
            ECB_UNROLL
            for (int j = 1; j < L; j += 2) {
                e[i + j] = (j == 1) ? madlo_cc(a[j], bi, e[i + j]) : madloc_cc(a[j], bi, e[i + j]);
                e[i + j + 1] = madhic_cc(a[j], bi, e[i + j + 1]);
            }
            ECB_ASSERT_NO_CARRY();
        }
    }
    r[0] = e[0];
    r[1] = add_cc(e[1], o[0]);
    ECB_UNROLL
    for (int i = 2; i < 2 * L - 1; i++) r[i] = addc_cc(e[i], o[i - 1]);
    r[2 * L - 1] = addc(e[2 * L - 1], o[2 * L - 2]);
}

// r[0..2L) = a^2: off-diagonal products once (same even/odd scheme), doubled, plus the diagonal.
template <int L> ECB_DEV void sqr_wide(u32* r, const u32* a) {
    if constexpr (L % 2 != 0) {
#if defined(__CUDA_ARCH__)
        sqr_wide_padded<L>(r, a);
#else
        mul_wide_columns<L>(r, a, a);
#endif
        return;
    }
    u32 e[2 * L], o[2 * L];
    ECB_UNROLL
    for (int i = 0; i < 2 * L; i++) { e[i] = 0; o[i] = 0; }
    // off-diagonal: sum_{i<j} a[i]*a[j] * 2^(32(i+j))
    ECB_UNROLL
    for (int i = 0; i < L - 1; i++) {
        const u32 ai = a[i];
        // j = i+1, i+3, ... : weight i+j odd -> o[i+j-1], o[i+j]
        {
            bool first = true;
            ECB_UNROLL
            for (int j = i + 1; j < L; j += 2) {
                o[i + j - 1] = first ? madlo_cc(a[j], ai, o[i + j - 1]) : madloc_cc(a[j], ai, o[i + j - 1]);
                o[i + j] = madhic_cc(a[j], ai, o[i + j]);
                first = false;
            }
            // carry out goes to the next o limb above the last pair touched
            const int last = i + 1 + 2 * ((L - 1 - (i + 1)) / 2);   // last j used
            if ((i & 1) && i <= L - 3) o[i + last + 1] = addc(o[i + last + 1], 0);   // other rows end on a fresh pair
            else ECB_ASSERT_NO_CARRY();
        }
        // j = i+2, i+4, ... : weight even -> e[i+j], e[i+j+1]
        if (i + 2 < L) {
            bool first = true;
            ECB_UNROLL
            for (int j = i + 2; j < L; j += 2) {
                e[i + j] = first ? madlo_cc(a[j], ai, e[i + j]) : madloc_cc(a[j], ai, e[i + j]);
                e[i + j + 1] = madhic_cc(a[j], ai, e[i + j + 1]);
                first = false;
            }
            const int last = i + 2 + 2 * ((L - 1 - (i + 2)) / 2);
            if (!(i & 1) && i >= 2 && i <= L - 4) e[i + last + 2] = addc(e[i + last + 2], 0);
            else ECB_ASSERT_NO_CARRY();
        }
    }
    // t = e + (o << 32)
    u32 t[2 * L];
    t[0] = e[0];
    t[1] = add_cc(e[1], o[0]);
    ECB_UNROLL
    for (int i = 2; i < 2 * L - 1; i++) t[i] = addc_cc(e[i], o[i - 1]);
    t[2 * L - 1] = addc(e[2 * L - 1], o[2 * L - 2]);
    // t = 2t
    ECB_UNROLL
    for (int i = 2 * L - 1; i > 0; i--) t[i] = (t[i] << 1) | (t[i - 1] >> 31);
    t[0] <<= 1;
    // r = t + diagonal
    ECB_UNROLL
    for (int i = 0; i < L; i++) {
        r[2 * i] = (i == 0) ? madlo_cc(a[i], a[i], t[0]) : madloc_cc(a[i], a[i], t[2 * i]);
        r[2 * i + 1] = (i == L - 1) ? madhic(a[i], a[i], t[2 * i + 1]) : madhic_cc(a[i], a[i], t[2 * i + 1]);
    }
}

// ------------------------------------------------------------------------------------------
// big-endian byte string <-> little-endian limbs (FB = 4L bytes)

template <int L> ECB_DEV void load_be_bytes(u32* v, const u8* p) {
    ECB_UNROLL
    for (int i = 0; i < L; i++) {
        const u8* q = p + 4 * (L - 1 - i);
        v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
    }
}
template <int L> ECB_DEV void store_be_bytes(u8* p, const u32* v) {
    ECB_UNROLL
    for (int i = 0; i < L; i++) {
        u8* q = p + 4 * (L - 1 - i);
        q[0] = (u8)(v[i] >> 24); q[1] = (u8)(v[i] >> 16); q[2] = (u8)(v[i] >> 8); q[3] = (u8)v[i];
    }
}
// N consecutive words between registers and global / local memory with the widest access the word count allows: 128-bit
// for N % 4 == 0, 64-bit for even N, 32-bit otherwise.  The address must be aligned accordingly (16 / 8 / 4 bytes): every
// internal limb array of the engine is (element sizes are multiples of 4L bytes on cudaMalloc'd bases).  Per-thread rows of
// the table kernels lie hundreds of bytes apart, so every access of a lane is its own L1 wavefront whatever its width: the
// 32-bit form cost k_kt_fill / k_wintab four times the LSU cycles (ncu: LDG.E x 32 wavefronts dominated the kernels).
template <int N> ECB_DEV void ld_words(u32* dst, const u32* src) {
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
    if constexpr (N % 4 == 0) {
        const ::uint4* q = reinterpret_cast<const ::uint4*>(src);
        ECB_UNROLL
        for (int k = 0; k < N / 4; k++) { const ::uint4 w = q[k]; dst[4 * k] = w.x; dst[4 * k + 1] = w.y; dst[4 * k + 2] = w.z; dst[4 * k + 3] = w.w; }
        return;
    } else if constexpr (N % 2 == 0) {
        const ::uint2* q = reinterpret_cast<const ::uint2*>(src);
        ECB_UNROLL
        for (int k = 0; k < N / 2; k++) { const ::uint2 w = q[k]; dst[2 * k] = w.x; dst[2 * k + 1] = w.y; }
        return;
    }
#endif
    ECB_UNROLL
    for (int k = 0; k < N; k++) dst[k] = src[k];
}
template <int N> ECB_DEV void st_words(u32* dst, const u32* src) {
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
    if constexpr (N % 4 == 0) {
        ::uint4* q = reinterpret_cast<::uint4*>(dst);
        ECB_UNROLL
        for (int k = 0; k < N / 4; k++) q[k] = make_uint4(src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]);
        return;
    } else if constexpr (N % 2 == 0) {
        ::uint2* q = reinterpret_cast<::uint2*>(dst);
        ECB_UNROLL
        for (int k = 0; k < N / 2; k++) q[k] = make_uint2(src[2 * k], src[2 * k + 1]);
        return;
    }
#endif
    ECB_UNROLL
    for (int k = 0; k < N; k++) dst[k] = src[k];
}

// The ABI's element arrays are 4L-byte strings one after another.  When the element address is aligned for the widest access
// its size allows (16 bytes for L % 4 == 0, 8 for even L, 4 otherwise - always true for the context's own staging buffers and
// for cudaMalloc'd caller buffers) the string moves as 128 / 64 / 32-bit words + byte swaps instead of 4L byte accesses; the
// test is uniform across a launch (same base, strides that are multiples of the element size).
template <int L> ECB_DEV void load_be(u32* v, const u8* p) {
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
    constexpr unsigned long long A = (L % 4 == 0) ? 15ull : (L % 2 == 0) ? 7ull : 3ull;
    if ((reinterpret_cast<unsigned long long>(p) & A) == 0ull) {
        u32 w[L];
        ld_words<L>(w, reinterpret_cast<const u32*>(p));
        ECB_UNROLL
        for (int i = 0; i < L; i++) v[i] = bswap32(w[L - 1 - i]);
        return;
    }
#endif
    load_be_bytes<L>(v, p);
}
template <int L> ECB_DEV void store_be(u8* p, const u32* v) {
#if defined(__CUDA_ARCH__) && !defined(ECB_EMU)
    constexpr unsigned long long A = (L % 4 == 0) ? 15ull : (L % 2 == 0) ? 7ull : 3ull;
    if ((reinterpret_cast<unsigned long long>(p) & A) == 0ull) {
        u32 w[L];
        ECB_UNROLL
        for (int i = 0; i < L; i++) w[i] = bswap32(v[L - 1 - i]);
        st_words<L>(reinterpret_cast<u32*>(p), w);
        return;
    }
#endif
    store_be_bytes<L>(p, v);
}

}  // namespace ecb
