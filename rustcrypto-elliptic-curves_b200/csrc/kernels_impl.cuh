// kernels_impl.cuh — __global__ wrappers around the bodies in kernels.cuh and the launcher table
// for one curve.  Included by exactly one translation unit per curve (curve_*.cu).
#pragma once
#include "kernels.cuh"
#include "launch.h"
#include <atomic>
#include <cstdlib>

namespace ecb {

constexpr int BLK = 128;   // threads per CTA of the arithmetic kernels (register-heavy: 2-4 CTAs per SM)
// Resident CTAs per SM the public-input kernels are compiled for (register cap 65536 / (128 N)).  Measured on the B200
// (k256 verify at 2^22 rows / P-256 verify at 2^20, M rows/s): N = 3: 50.7 / 27.6, 4: 50.7 / 28.5, 5: 51.9 / 28.7,
// 6: 52.7 / 28.9, 7: 53.0 / 29.1, 8: 52.7 / 29.1.  6 (80 registers, ~120 bytes of spills) is the knee for the 8-limb curves;
// the 12-limb curve keeps 4 (3 was slower, see git history).  Re-measured on the final round-1 build at 2^22 rows (batch-affine window
// tables, sign-driven reduction): 6 / 7 / 8 = 52.77 / 53.07 / 52.74 (k256), 32.72 / 32.66 / 32.91 (P-256); P-384 at 4 / 5 / 6 = 10.26 / 10.32 / 10.41.
#ifndef ECB_FAST_MIN_CTAS
#define ECB_FAST_MIN_CTAS 6
#endif
#ifndef ECB_FAST_MIN_CTAS_WIDE
#define ECB_FAST_MIN_CTAS_WIDE 4
#endif
// secp256k1 takes 7 (72 registers, 140 / 204 bytes of spill stores / loads in the verify kernel): three back-to-back A/B rounds on one
// box at 2^22 rows, 6 vs 7: verify 53.19 / 53.19 / 53.20 vs 53.50 / 53.49 / 53.50 M/s (+0.6 %), P*k 57.85 vs 58.28 M/s (+0.7 %);
// the earlier sweeps had shown the same +0.5 % twice.  P-256 / SM2 measured flat to slightly slower at 7 and stay at 6.
#ifndef ECB_FAST_MIN_CTAS_K256
#define ECB_FAST_MIN_CTAS_K256 7
#endif
template <class C> constexpr int fast_min_ctas() { return C::L > 8 ? ECB_FAST_MIN_CTAS_WIDE : (C::A_IS_ZERO ? ECB_FAST_MIN_CTAS_K256 : ECB_FAST_MIN_CTAS); }
// same knob for the secret-scalar kernels (complete formulas, full table scans).  Measured (k256 G*k CT at 2^18 / k256 P*k CT /
// P-256 G*k CT / P-256 P*k CT, M/s): unconstrained (146-154 registers, 3 CTAs) 93.5 / 37.7 / 70.5 / 15.6,
// 4: 95.5 / 39.4 / 73.6 / 16.3, 5: 93.4 / 39.3 / 70.2 / 15.6, 6: 91.5 / 39.7 / 60.8 / 16.2.
#ifndef ECB_CT_MIN_CTAS
#define ECB_CT_MIN_CTAS 4
#endif
template <class C> constexpr int ct_min_ctas() { return C::L > 8 ? 3 : ECB_CT_MIN_CTAS; }

// Curve descriptor with the two-call squarer (mont.cuh SQSPLIT), for the kernels in which ptxas compiles the one-function
// squarer of P-384 with spilled carry predicates (tools/sass_funcs.py): the complete-formula (secret-scalar) scalar
// multiplication, normalisation (one 384-squaring inversion per thread), SEC1 decoding (square root), the window-table kernel.
template <class C> struct SplitSqr : C { typedef Mont<typename C::F::Params, true> F; };
template <class C> struct CtCurve { typedef C type; };
template <> struct CtCurve<CurveP384> { typedef SplitSqr<CurveP384> type; };

// One inversion chain per CTA for the Montgomery-trick kernels (normalise, verify prep, window tables, sign finish).
// Lanes run in lockstep, so an inversion costs a warp the same ~270-330 multiplications whether one lane or all 32 run
// it; sharing it only pays across warps.  Each thread brings the product a of its rows.  Inside a warp: inclusive prefix
// and suffix products by shuffles (5 + 5 multiplications); the four warp products go through shared memory, thread 0
// inverts their product and gives every warp the inverse of its own product (Montgomery's trick over four values); a
// thread's result is (product of the lanes below) x (product of the lanes above) x (inverse of its warp's product).
// While thread 0 runs the chain the other three warps wait at the barrier and the SM issues other CTAs: per thread the
// kernels spend 12 + ~280/4 multiplications on inversion instead of ~280.  No secret-dependent branch or address (the
// sign-finish kernel inverts secret nonces).  The product must be non-zero: every body substitutes 1 for a zero factor.
// -DECB_BLOCK_INV=0 restores one chain per thread (A/B measurements).
#ifndef ECB_BLOCK_INV
#define ECB_BLOCK_INV 1
#endif
struct BlockInv {
    static constexpr bool COOPERATIVE = true;
    template <class E> __device__ __forceinline__ static void shfl_up(E& r, const E& a, int d) {
#pragma unroll
        for (int w = 0; w < (int)(sizeof(E) / sizeof(u32)); w++) r.v[w] = __shfl_up_sync(0xffffffffu, a.v[w], d);
    }
    template <class E> __device__ __forceinline__ static void shfl_down(E& r, const E& a, int d) {
#pragma unroll
        for (int w = 0; w < (int)(sizeof(E) / sizeof(u32)); w++) r.v[w] = __shfl_down_sync(0xffffffffu, a.v[w], d);
    }
    template <class FF> __device__ static void run(typename FF::E& r, const typename FF::E& a) {
        typedef typename FF::E E;
        constexpr int W = (int)(sizeof(E) / sizeof(u32));
        constexpr int NW = BLK / 32;
        __shared__ u32 sh[2 * NW * W];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        E pre = a, suf = a, t, m;
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            shfl_up(t, pre, d);
            FF::mul(m, pre, t);
            FF::select(pre, lane >= d, m, pre);
            shfl_down(t, suf, d);
            FF::mul(m, suf, t);
            FF::select(suf, lane + d < 32, m, suf);
        }
        E below, above, one;
        FF::set_one(one);
        shfl_up(t, pre, 1);
        FF::select(below, lane == 0, one, t);
        shfl_down(t, suf, 1);
        FF::select(above, lane == 31, one, t);
        if (lane == 31) {
#pragma unroll
            for (int w = 0; w < W; w++) sh[warp * W + w] = pre.v[w];
        }
        __syncthreads();
        // The chain runs on thread 0.  Picking the warp by a hash of the CTA index, so that the chains of the CTAs resident
        // on one SM land on different sub-partitions (warp w sits on sub-partition w % 4), was measured and is SLOWER on the
        // B200: k_verify_prep<K256> 1.74 -> 2.05 ms per 2^22 rows, <P256> 0.51 -> 0.57 ms per 2^20 (ncu launch lists, same box);
        // per-sub-partition instruction counts differ by only 14 % with the fixed choice.  -DECB_BLOCK_INV_SPREAD=1 builds it.
#ifndef ECB_BLOCK_INV_SPREAD
#define ECB_BLOCK_INV_SPREAD 0
#endif
        const int chain_thread = ECB_BLOCK_INV_SPREAD ? (int)((blockIdx.x * 0x9E3779B1u) >> 30) % NW * 32 : 0;
        if ((int)threadIdx.x == chain_thread) {   // warp products in sh[0 .. NW), running products before each warp in sh[NW .. 2 NW), overwritten by the result
            E tot, inv;
            FF::set_one(tot);
#pragma unroll 1
            for (int k = 0; k < NW; k++) {
#pragma unroll
                for (int w = 0; w < W; w++) { t.v[w] = sh[k * W + w]; sh[(NW + k) * W + w] = tot.v[w]; }
                FF::mul(tot, tot, t);
            }
            FF::inv_trick(inv, tot);      // divsteps (safegcd.cuh): ~4x shorter dependent path than the Fermat chain
#pragma unroll 1
            for (int k = NW - 1; k >= 0; k--) {
#pragma unroll
                for (int w = 0; w < W; w++) { t.v[w] = sh[(NW + k) * W + w]; m.v[w] = sh[k * W + w]; }
                FF::mul(t, inv, t);
                FF::mul(inv, inv, m);
#pragma unroll
                for (int w = 0; w < W; w++) sh[(NW + k) * W + w] = t.v[w];
            }
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < W; w++) t.v[w] = sh[(NW + warp) * W + w];
        FF::mul(m, below, above);
        FF::mul(r, m, t);
    }
};
#if ECB_BLOCK_INV
typedef BlockInv TrickInv;
#else
typedef OwnInv TrickInv;
#endif

template <class C> __global__ void __launch_bounds__(BLK) k_field_op(int n, int which, int op, const u8* a, const u8* b, u8* out, u8* ok) {
    Bodies<C>::body_field_op(blockIdx.x * BLK + threadIdx.x, n, which, op, a, b, out, ok);
}
template <class C, bool CT> __global__ void __launch_bounds__(BLK, ct_min_ctas<C>()) k_mul_var(int n, u32 flags, const u8* pts, const u8* inf, const u8* k, u32* proj, u8* invalid) {
    Bodies<typename CtCurve<C>::type>::template body_mul_var<CT>(blockIdx.x * BLK + threadIdx.x, n, flags, pts, inf, k, proj, invalid);
}
template <class C, bool CT> __global__ void __launch_bounds__(BLK) k_mul_gen(int n, const u8* k, const u32* tab, u32* proj) {
    Bodies<C>::template body_mul_gen<CT>(blockIdx.x * BLK + threadIdx.x, n, k, tab, proj);
}
template <class C> __global__ void __launch_bounds__(BLK) k_load_proj(int n, const u8* xyz, u32* proj, u8* invalid) {
    Bodies<C>::body_load_proj(blockIdx.x * BLK + threadIdx.x, n, xyz, proj, invalid);
}
template <class C> __global__ void __launch_bounds__(BLK) k_normalize(int n, const u32* proj, int mode, int compress, u8* out_bytes, u8* out_inf, u32* out_limbs) {
    Bodies<typename CtCurve<C>::type>::template body_normalize<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, proj, mode, compress, out_bytes, out_inf, out_limbs);
}
// proj[i] += sum_with[i] (complete addition), then the same normalisation: tail of the split fixed-base path
template <class C> __global__ void __launch_bounds__(BLK) k_sum_normalize(int n, u32* proj, const u32* sum_with, int mode, int compress, u8* out_bytes, u8* out_inf, u32* out_limbs) {
    Bodies<typename CtCurve<C>::type>::template body_normalize<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, proj, mode, compress, out_bytes, out_inf, out_limbs, sum_with);
}
template <class C> __global__ void __launch_bounds__(BLK) k_verify(int n, const u8* q, const u8* z, const u8* rs, const u32* gtab, u8* ok) {
    Bodies<C>::body_verify(blockIdx.x * BLK + threadIdx.x, n, q, z, rs, gtab, ok);
}
// secp256k1: ECB_WIN_SMEM_ENTRIES of the 8 per-thread window-table entries live in dynamic shared memory (jac.cuh
// WinTab).  Measured on the B200 at 2^22 rows with 7 entries (56 KB per CTA, still 4 CTAs per SM): DRAM traffic of the
// verify kernel 7.3 -> 3.3 GB per launch, but 4 % slower (85.7 vs 82.1 ms) - the 224 KB carve-out leaves almost no L1
// for the remaining local frames and the fixed-base gathers; with 4 entries (32 KB per CTA) 4.5 GB and 1 % slower.
// Throughput wins: the default keeps the table in local memory.
#ifndef ECB_WIN_SMEM_ENTRIES
#define ECB_WIN_SMEM_ENTRIES 0
#endif
template <class C> constexpr int win_smem_entries() { return C::A_IS_ZERO ? ECB_WIN_SMEM_ENTRIES : 0; }
template <class C> constexpr u32 win_smem_bytes() { return (u32)win_smem_entries<C>() * 16u * 4u * BLK; }
extern __shared__ __align__(16) u32 ecb_dyn_smem[];

// wtab != NULL: window tables come from k_wintab (affine, global memory) instead of a per-thread Jacobian table (primeorder curves)
template <class C> __global__ void __launch_bounds__(BLK, fast_min_ctas<C>()) k_mul_var_fast(int n, const u8* pts, const u32* aff_limbs, const u8* inf, const u8* k, u32* proj, u8* invalid,
                                                                                                       const u32* wtab) {
    Bodies<C>::template body_mul_var_fast<win_smem_entries<C>()>(blockIdx.x * BLK + threadIdx.x, n, pts, aff_limbs, inf, k, proj, invalid,
                                                                      ecb_dyn_smem + threadIdx.x, BLK, wtab);
}
// per-row affine window tables {1..8}Q for the primeorder public-input kernels (Montgomery's trick over the rows a thread owns)
#ifndef ECB_WT_MIN_CTAS
#define ECB_WT_MIN_CTAS 4
#endif
template <class C> constexpr int wt_min_ctas() { return C::L > 8 ? 3 : ECB_WT_MIN_CTAS; }
template <class C> __global__ void __launch_bounds__(BLK, wt_min_ctas<C>()) k_wintab(int n, const u8* pts, const u32* aff_limbs, u32* wtab, u32* zbuf) {
    Bodies<typename CtCurve<C>::type>::template body_wintab<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, pts, aff_limbs, wtab, zbuf);
}
template <class C> __global__ void __launch_bounds__(BLK) k_verify_prep(int n, int mode, const u8* z, const u8* rs, u32* scratch) {
    Bodies<C>::template body_verify_prep<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, mode, z, rs, scratch);
}
// MODE is a template parameter: the ECDSA instance carries no decompression / projective-output code
template <class C, int MODE> __global__ void __launch_bounds__(BLK, fast_min_ctas<C>()) k_verify_main(int n, const u8* q, const u8* rs, const u8* z, const u8* aux, const u32* scratch, const u32* gbig, int gw, u8* ok, u32* proj_out,
                                                                                                        const u32* wtab) {
    Bodies<C>::template body_verify_main<win_smem_entries<C>()>(blockIdx.x * BLK + threadIdx.x, n, MODE, q, rs, z, aux, scratch, gbig, gw, ok, proj_out,
                                                                 ecb_dyn_smem + threadIdx.x, BLK, wtab);
}
// ---- per-key window tables: grouping (four small kernels), construction (base points, batched-affine fill), verification
template <class C> __global__ void __launch_bounds__(256) k_kt_lookup(int n, const u32* q32, const int* htab, u32 hmask, const u32* gkeys, int* gid) {
    Bodies<C>::body_kt_lookup(blockIdx.x * 256 + threadIdx.x, n, q32, htab, hmask, gkeys, gid);
}
template <class C> __global__ void __launch_bounds__(256) k_kt_insert(int n, const u32* q32, int* htab, u32 hmask, const int* gid, int* rep, int* rep_slot) {
    Bodies<C>::body_kt_insert(blockIdx.x * 256 + threadIdx.x, n, q32, htab, hmask, gid, rep, rep_slot);
}
template <class C> __global__ void __launch_bounds__(256) k_kt_number(int n, const u32* q32, int* htab, const int* gid, const int* rep, const int* rep_slot, int* counter, int cap,
                                                                       u32* gkeys, int* newgid) {
    Bodies<C>::body_kt_number(blockIdx.x * 256 + threadIdx.x, n, q32, htab, gid, rep, rep_slot, counter, cap, gkeys, newgid);
}
template <class C> __global__ void __launch_bounds__(256) k_kt_assign(int n, int* gid, const int* rep, const int* newgid) {
    Bodies<C>::body_kt_assign(blockIdx.x * 256 + threadIdx.x, n, gid, rep, newgid);
}
// W: window width of the tables (0 = the curve's default, else the narrow width - see kt_narrow below)
template <class C, int W = 0> __global__ void __launch_bounds__(BLK, fast_min_ctas<C>()) k_kt_base(int g0, int cnt, const u32* gkeys, u32* proj, u8* kvalid) {
    Bodies<C, W>::body_kt_base(blockIdx.x * BLK + threadIdx.x, g0, cnt, gkeys, proj, kvalid);
}
#ifndef ECB_KTF_MIN_CTAS
#define ECB_KTF_MIN_CTAS ECB_WT_MIN_CTAS
#endif
template <class C> constexpr int ktf_min_ctas() { return C::L > 8 ? 3 : ECB_KTF_MIN_CTAS; }
template <class C, int W = 0> __global__ void __launch_bounds__(BLK, ktf_min_ctas<C>()) k_kt_fill(int items, u32* tab) {
    Bodies<typename CtCurve<C>::type, W>::template body_kt_fill<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, items, tab);
}
// resident CTAs the table-path main kernel is compiled for (its own knob: no per-thread table, smaller frame than k_verify_main)
#ifndef ECB_KT_MIN_CTAS
#define ECB_KT_MIN_CTAS 0
#endif
// Measured at 2^22 rows / 2^16 keys (P-384: 2^20 rows), M verifies/s at the per-row kernels' occupancy / 7 / 8 resident CTAs
// (profiles/r02_ab_occupancy_table_kernels.txt): secp256k1 (7) 123.5 / 123.5 / 121.7, P-256 (6) 111.0 / 111.8 / 112.3, P-384 (4) 21.17 /
// 21.82 / 21.23.  The +1 .. 3 % of the higher settings come with spills (P-256 at 64 registers: 76 / 60 B instead of 12 / 12;
// P-384 at 72: 160 / 136 B instead of none) and were not taken.
template <class C> constexpr int kt_min_ctas() { return ECB_KT_MIN_CTAS ? ECB_KT_MIN_CTAS : fast_min_ctas<C>(); }
template <class C, int MODE, int W = 0> __global__ void __launch_bounds__(BLK, kt_min_ctas<C>()) k_verify_keytab(int n, const u8* rs, const u8* z, const u32* scratch, const int* gid, const u8* kvalid,
                                                                                                                     const u32* tab, const u32* gbig, int gw, u8* ok) {
    Bodies<C, W>::body_verify_keytab(blockIdx.x * BLK + threadIdx.x, n, MODE, rs, z, scratch, gid, kvalid, tab, gbig, gw, ok);
}
// narrow table width of a curve (calls with few rows per key); equal to the default width where only one is built
template <class C> constexpr int kt_narrow() { return Bodies<C>::KT_W_DEFAULT > 4 ? 4 : Bodies<C>::KT_W_DEFAULT; }
template <class C> constexpr bool kt_two_widths() { return kt_narrow<C>() != Bodies<C>::KT_W_DEFAULT; }
template <class C> __global__ void __launch_bounds__(BLK) k_decode(int n, int mode, const u8* enc, int stride, u8* xy, u8* status) {
    Bodies<typename CtCurve<C>::type>::body_decode(blockIdx.x * BLK + threadIdx.x, n, mode, enc, stride, xy, status);
}
template <class C> __global__ void __launch_bounds__(BLK) k_finish(int n, int kind, const u8* a, int stride, const u8* inf, const u8* rs, u8* ok) {
    Bodies<C>::body_finish(blockIdx.x * BLK + threadIdx.x, n, kind, a, stride, inf, rs, ok);
}
template <class C> __global__ void __launch_bounds__(BLK) k_sign_finish(int n, const u8* d, const u8* k, const u8* z, const u32* aff, u8* rs_out, u8* recid_out, u8* ok_out) {
    Bodies<C>::template body_sign_finish<TrickInv>(blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, d, k, z, aff, rs_out, recid_out, ok_out);
}
template <class C> __global__ void __launch_bounds__(BLK) k_proj_to_bytes(int n, const u32* proj, u8* xyz) {
    int tid = blockIdx.x * BLK + threadIdx.x;
    if (tid >= n) return;
    typedef Bodies<C> B;
    typename B::Proj p;
    B::load_proj_limbs(p, proj + (size_t)tid * 3 * C::L);
    u32 t[C::L];
    C::F::to_limbs(t, p.X); store_be<C::L>(xyz + (size_t)tid * 3 * C::FB, t);
    C::F::to_limbs(t, p.Y); store_be<C::L>(xyz + (size_t)tid * 3 * C::FB + C::FB, t);
    C::F::to_limbs(t, p.Z); store_be<C::L>(xyz + (size_t)tid * 3 * C::FB + 2 * C::FB, t);
}
// out[i] = a[i] + b[i] (complete addition on internal projective limbs): the tail of the per-row two-term lincomb
template <class C> __global__ void __launch_bounds__(BLK) k_add_proj(int n, const u32* a, const u32* b, u32* out, u8* invalid, const u8* invalid_b) {
    int tid = blockIdx.x * BLK + threadIdx.x;
    if (tid >= n) return;
    typedef Bodies<C> B;
    typename B::Proj p, q, r;
    B::load_proj_limbs(p, a + (size_t)tid * 3 * C::L);
    B::load_proj_limbs(q, b + (size_t)tid * 3 * C::L);
    EC<C>::add(r, p, q);
    u32 bad = (invalid ? invalid[tid] : 0u) | (invalid_b ? invalid_b[tid] : 0u);
    if (bad) EC<C>::set_identity(r);          // an invalid term voids the row (the reference cannot represent such a point)
    B::store_proj(out + (size_t)tid * 3 * C::L, r);
    if (invalid) invalid[tid] = bad ? 1 : 0;
}
// block-level sum of projective points: thread-strided partials, then a shared-memory tree
template <class C> __global__ void __launch_bounds__(BLK) k_sum(int n, const u32* proj, u32* out) {
    typedef Bodies<C> B;
    typedef typename B::Proj Proj;
    __shared__ u32 sh[BLK * 3 * C::L];
    Proj acc;
    B::body_partial_sum(acc, blockIdx.x * BLK + threadIdx.x, gridDim.x * BLK, n, proj);
    B::store_proj(sh + threadIdx.x * 3 * C::L, acc);
    __syncthreads();
    for (int stride = BLK / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            Proj a, b;
            B::load_proj_limbs(a, sh + threadIdx.x * 3 * C::L);
            B::load_proj_limbs(b, sh + (threadIdx.x + stride) * 3 * C::L);
            EC<C>::add(a, a, b);
            B::store_proj(sh + threadIdx.x * 3 * C::L, a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3 * C::L; i++) out[(size_t)blockIdx.x * 3 * C::L + i] = sh[i];
    }
}

// ---------------------------------------------------------------------------------------------
// fixed-base kernel with the (8L+1) x 8 affine table staged in shared memory by one TMA bulk
// copy per CTA (cp.async.bulk global -> shared, completion on an mbarrier; UBLKCP in SASS).
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }

template <class C, bool CT> __global__ void __launch_bounds__(BLK, ct_min_ctas<C>()) k_mul_gen_smem(int n, const u8* k, const u32* tab, u32 tab_bytes, u32* proj) {
    extern __shared__ __align__(128) u32 stab[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(tab_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(stab)), "l"(tab), "r"(tab_bytes), "r"(smem_addr(&bar)) : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_addr(&bar)) : "memory");
    }
    Bodies<C>::template body_mul_gen<CT>(blockIdx.x * BLK + threadIdx.x, n, k, stab, proj);
}

// Split fixed-base kernel (kernels.cuh body_gen_half): blockIdx.y picks the half of the windows, the CTA stages that half of
// the table with one TMA bulk copy.  Compiled for the public-input kernels' occupancy (Jacobian arithmetic, acc in local
// memory): 7 resident CTAs per SM on the 8-limb curves, so that the 2 x 512 CTAs of a 2^16-row batch are one wave.
#ifndef ECB_GEN2_MIN_CTAS
#define ECB_GEN2_MIN_CTAS 7
#endif
template <class C> constexpr int gen2_min_ctas() { return C::L > 8 ? 3 : ECB_GEN2_MIN_CTAS; }
template <class C, bool CT> __global__ void __launch_bounds__(BLK, gen2_min_ctas<C>()) k_gen_half(int n, const u8* k, const u32* tab, u32* part) {
    typedef Bodies<C> B;
    extern __shared__ __align__(128) u32 stab[];
    __shared__ __align__(8) unsigned long long bar;
    const int half = blockIdx.y;
    const int w0 = half * B::G2_NH;
    const int nw = half ? B::G2_WINDOWS - B::G2_NH : B::G2_NH;
    const u32 bytes = (u32)nw * B::G2_E * 2u * C::L * 4u;
    const u32* src = tab + (size_t)w0 * B::G2_E * 2 * C::L;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(stab)), "l"(src), "r"(bytes), "r"(smem_addr(&bar)) : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_addr(&bar)) : "memory");
    }
    B::template body_gen_half<CT>(blockIdx.x * BLK + threadIdx.x, n, half, k, stab, part);
}

// ---------------------------------------------------------------------------------------------
template <class C> struct Launch {
    static int grid(int n) { return (n + BLK - 1) / BLK; }
    // The opt-in for > 48 KB of dynamic shared memory is per function AND per device: remember it per device, so a process
    // that drives several GPUs (one context each) sets it on every one of them.
    // (atomic: under ecb200_init_multi one host thread per device runs these launchers concurrently)
    static bool first_use_on_device(std::atomic<bool> (&seen)[64]) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        return !seen[dev].exchange(true);
    }

    static void field_op(cudaStream_t s, int n, int which, int op, const u8* a, const u8* b, u8* out, u8* ok) {
        if (n <= 0) return;
        k_field_op<C><<<grid(n), BLK, 0, s>>>(n, which, op, a, b, out, ok);
        count_launch();
    }
    static void mul_var(cudaStream_t s, bool ct, int n, u32 flags, const u8* pts, const u8* inf, const u8* k, u32* proj, u8* invalid) {
        if (n <= 0) return;
        if (ct) k_mul_var<C, true><<<grid(n), BLK, 0, s>>>(n, flags, pts, inf, k, proj, invalid);
        else k_mul_var<C, false><<<grid(n), BLK, 0, s>>>(n, flags, pts, inf, k, proj, invalid);
        count_launch();
    }
    static void mul_gen(cudaStream_t s, bool ct, int n, const u8* k, const u32* tab, u32* proj) {
        if (n <= 0) return;
        const u32 bytes = (u32)Bodies<C>::GEN_WINDOWS * 8u * 2u * C::L * 4u;   // 33 KB (L = 8), 74.5 KB (L = 12)
        static std::atomic<bool> seen[64] = {};
        if (first_use_on_device(seen)) {
            cudaFuncSetAttribute(k_mul_gen_smem<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            cudaFuncSetAttribute(k_mul_gen_smem<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        }
        if (ct) k_mul_gen_smem<C, true><<<grid(n), BLK, bytes, s>>>(n, k, tab, bytes, proj);
        else k_mul_gen_smem<C, false><<<grid(n), BLK, bytes, s>>>(n, k, tab, bytes, proj);
        count_launch();
    }
    // split form: part = 2 x n x 3L limbs (lower-half sums, then upper-half sums); finish with sum_normalize(part, part + n * 3L)
    static void mul_gen2(cudaStream_t s, bool ct, int n, const u8* k, const u32* tab2, u32* part) {
        if (n <= 0) return;
        typedef Bodies<C> B;
        const u32 bytes = (u32)B::G2_NH * B::G2_E * 2u * C::L * 4u;   // the larger half
        static std::atomic<bool> seen[64] = {};
        if (first_use_on_device(seen)) {
            cudaFuncSetAttribute(k_gen_half<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            cudaFuncSetAttribute(k_gen_half<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        }
        const dim3 g((unsigned)grid(n), 2u);
        if (ct) k_gen_half<C, true><<<g, BLK, bytes, s>>>(n, k, tab2, part);
        else k_gen_half<C, false><<<g, BLK, bytes, s>>>(n, k, tab2, part);
        count_launch();
    }
    static void sum_normalize(cudaStream_t s, int n, u32* proj, const u32* sum_with, int mode, int compress, u8* out_bytes, u8* out_inf, u32* out_limbs) {
        if (n <= 0) return;
        const int per = 148 * trick_ctas_per_sm() * BLK;
        int ept = (n + per - 1) / per;
        if (ept < 1) ept = 1;
        if (ept > Bodies<C>::EPT) ept = Bodies<C>::EPT;
        int threads = (n + ept - 1) / ept;
        k_sum_normalize<C><<<grid(threads), BLK, 0, s>>>(n, proj, sum_with, mode, compress, out_bytes, out_inf, out_limbs);
        count_launch();
    }
    static void load_proj(cudaStream_t s, int n, const u8* xyz, u32* proj, u8* invalid) {
        if (n <= 0) return;
        k_load_proj<C><<<grid(n), BLK, 0, s>>>(n, xyz, proj, invalid);
        count_launch();
    }
    // resident CTAs per SM the Montgomery-trick kernels aim at when they pick the rows per thread (ECB200_TRICK_CTAS_PER_SM
    // overrides, for measurements)
    static int trick_ctas_per_sm() {
        static const int v = [] { const char* e = getenv("ECB200_TRICK_CTAS_PER_SM"); const int x = e ? atoi(e) : 0; return x >= 1 && x <= 16 ? x : 1; }();
        return v;
    }
    static void normalize(cudaStream_t s, int n, const u32* proj, int mode, int compress, u8* out_bytes, u8* out_inf, u32* out_limbs) {
        if (n <= 0) return;
        // elements per thread: the kernel is one ~270-multiplication inversion chain per thread plus 5 multiplications per
        // element, so below ~2 warps per SM sub-partition it is latency-bound and more elements per thread are free:
        // aim at 1 CTA per SM, at most EPT (config 1 at 2^16, whole call: 0.944 ms with one element per thread, 0.890 ms at
        // 2 CTAs per SM, 0.865 ms at 1)
        const int per = 148 * trick_ctas_per_sm() * BLK;
        int ept = (n + per - 1) / per;
        if (ept < 1) ept = 1;
        if (ept > Bodies<C>::EPT) ept = Bodies<C>::EPT;
        int threads = (n + ept - 1) / ept;
        k_normalize<C><<<grid(threads), BLK, 0, s>>>(n, proj, mode, compress, out_bytes, out_inf, out_limbs);
        count_launch();
    }
    static constexpr int SUM_BLOCKS = 148;
    static void sum(cudaStream_t s, int n, const u32* proj, u32* partial, u32* out) {
        int blocks = grid(n);
        if (blocks > SUM_BLOCKS) blocks = SUM_BLOCKS;
        if (blocks < 1) blocks = 1;
        if (blocks == 1) {
            k_sum<C><<<1, BLK, 0, s>>>(n, proj, out);
            count_launch();
        } else {
            k_sum<C><<<blocks, BLK, 0, s>>>(n, proj, partial);
            k_sum<C><<<1, BLK, 0, s>>>(blocks, partial, out);
            count_launch(2);
        }
    }
    static void add_proj(cudaStream_t s, int n, const u32* a, const u32* b, u32* out, u8* invalid, const u8* invalid_b) {
        if (n <= 0) return;
        k_add_proj<C><<<grid(n), BLK, 0, s>>>(n, a, b, out, invalid, invalid_b);
        count_launch();
    }
    static void proj_to_bytes(cudaStream_t s, int n, const u32* proj, u8* xyz) {
        if (n <= 0) return;
        k_proj_to_bytes<C><<<grid(n), BLK, 0, s>>>(n, proj, xyz);
        count_launch();
    }
    static void verify(cudaStream_t s, int n, const u8* q, const u8* z, const u8* rs, const u32* gtab, u8* ok) {
        if (n <= 0) return;
        k_verify<C><<<grid(n), BLK, 0, s>>>(n, q, z, rs, gtab, ok);
        count_launch();
    }
    // Affine window tables for n rows (primeorder curves only; a no-op on secp256k1, whose shared-Z table needs no inversion).
    // wtab: n x 16L words of tables followed by n x 7L words of scratch for the Z coordinates (wintab_words(n) in all).
    static void wintab(cudaStream_t s, int n, const u8* pts, const u32* aff_limbs, u32* wtab) {
        if (n <= 0 || C::A_IS_ZERO) return;
        // One launch per round of at most one wave of resident CTAs x WT_EPT rows per thread: a launch never ends in a
        // nearly empty second wave, and the rounds are equal (2^22 rows = 3 rounds of 13 rows per thread, not 2 + a rest).
        // Small batches still take 3 rows per thread so that the inversion (~270 multiplications) is shared.
        const long wave = (long)148 * wt_min_ctas<C>() * BLK;
        const int rounds = (int)((n + wave * Bodies<C>::WT_EPT - 1) / (wave * Bodies<C>::WT_EPT));
        const int per_round = (n + rounds - 1) / rounds;
        u32* zbuf = wtab + (size_t)n * 16 * C::L;
        for (int off = 0; off < n; off += per_round) {
            const int cnt = n - off < per_round ? n - off : per_round;
            int ept = (int)((cnt + wave - 1) / wave);
            if (ept < 3) ept = 3;
            const int threads = (cnt + ept - 1) / ept;
            k_wintab<C><<<grid(threads), BLK, 0, s>>>(cnt, pts ? pts + (size_t)off * 2 * C::FB : nullptr, aff_limbs ? aff_limbs + (size_t)off * 2 * C::L : nullptr,
                                                      wtab + (size_t)off * 16 * C::L, zbuf + (size_t)off * 7 * C::L);
            count_launch();
        }
    }
    static void mul_var_fast(cudaStream_t s, int n, const u8* pts, const u32* aff_limbs, const u8* inf, const u8* k, u32* proj, u8* invalid,
                             const u32* wtab) {
        if (n <= 0) return;
        static std::atomic<bool> seen[64] = {};
        if (win_smem_bytes<C>() > 0 && first_use_on_device(seen))
            cudaFuncSetAttribute(k_mul_var_fast<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem_bytes<C>());
        k_mul_var_fast<C><<<grid(n), BLK, win_smem_bytes<C>(), s>>>(n, pts, aff_limbs, inf, k, proj, invalid, C::A_IS_ZERO ? nullptr : wtab);
        count_launch();
    }
    static int ept_threads(int n) {   // threads for the Montgomery-trick kernels: rows per thread that keep ~1 CTA per SM busy (latency-bound below that, see normalize)
        const int per = 148 * trick_ctas_per_sm() * BLK;
        int ept = (n + per - 1) / per;
        if (ept < 1) ept = 1;
        if (ept > Bodies<C>::PREP_EPT) ept = Bodies<C>::PREP_EPT;
        return (n + ept - 1) / ept;
    }
    static void verify_prep(cudaStream_t s, int n, int mode, const u8* z, const u8* rs, u32* scratch) {
        if (n <= 0) return;
        k_verify_prep<C><<<grid(ept_threads(n)), BLK, 0, s>>>(n, mode, z, rs, scratch);
        count_launch();
    }
    static void verify_main(cudaStream_t s, int n, int mode, const u8* q, const u8* rs, const u8* z, const u8* aux, const u32* scratch,
                            const u32* gbig, int gw, u8* ok, u32* proj_out, const u32* wtab) {
        if (n <= 0) return;
        // Schnorr exists for secp256k1 only, SM2DSA for SM2 only (abi.cu rejects other combinations before launching)
        const u32 sm = win_smem_bytes<C>();
        static std::atomic<bool> seen[64] = {};
        if (sm > 0 && first_use_on_device(seen)) {
            cudaFuncSetAttribute(k_verify_main<C, VM_ECDSA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(k_verify_main<C, VM_RECOVER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if constexpr (C::A_IS_ZERO) cudaFuncSetAttribute(k_verify_main<C, VM_SCHNORR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        }
        if (C::A_IS_ZERO) wtab = nullptr;
        if (mode == VM_ECDSA) k_verify_main<C, VM_ECDSA><<<grid(n), BLK, sm, s>>>(n, q, rs, z, aux, scratch, gbig, gw, ok, proj_out, wtab);
        else if (mode == VM_RECOVER) k_verify_main<C, VM_RECOVER><<<grid(n), BLK, sm, s>>>(n, q, rs, z, aux, scratch, gbig, gw, ok, proj_out, nullptr);
        else if (mode == VM_SCHNORR) {
            if constexpr (C::A_IS_ZERO) k_verify_main<C, VM_SCHNORR><<<grid(n), BLK, sm, s>>>(n, q, rs, z, aux, scratch, gbig, gw, ok, proj_out, nullptr);
        } else if (mode == VM_SM2DSA) {
            if constexpr (C::ID == 3) k_verify_main<C, VM_SM2DSA><<<grid(n), BLK, sm, s>>>(n, q, rs, z, aux, scratch, gbig, gw, ok, proj_out, wtab);
        }
        count_launch();
    }
    static void kt_group(cudaStream_t s, int n, const u32* q32, int* htab, u32 hmask, u32* gkeys, int* gid, int* rep, int* rep_slot, int* newgid, int* counter, int cap) {
        if (n <= 0) return;
        const int g = (n + 255) / 256;
        k_kt_lookup<C><<<g, 256, 0, s>>>(n, q32, htab, hmask, gkeys, gid);
        k_kt_insert<C><<<g, 256, 0, s>>>(n, q32, htab, hmask, gid, rep, rep_slot);
        k_kt_number<C><<<g, 256, 0, s>>>(n, q32, htab, gid, rep, rep_slot, counter, cap, gkeys, newgid);
        k_kt_assign<C><<<g, 256, 0, s>>>(n, gid, rep, newgid);
        count_launch(4);
    }
    // tables of groups [g0, g0 + cnt): 16^w * Q (Jacobian chain), one normalisation of all base points straight into entry 0 of
    // every window, then seven rounds of batched affine additions
    // wide = 1: the curve's default width, 0: the narrow one (kt_narrow; the same kernels where a curve has one width)
    static void kt_build(cudaStream_t s, int wide, int g0, int cnt, const u32* gkeys, u32* proj_scratch, u8* kvalid, u32* tab) {
        if (cnt <= 0) return;
        if constexpr (kt_two_widths<C>()) {
            if (!wide) { kt_build_w<kt_narrow<C>()>(s, g0, cnt, gkeys, proj_scratch, kvalid, tab); return; }
        }
        kt_build_w<0>(s, g0, cnt, gkeys, proj_scratch, kvalid, tab);
    }
    template <int W> static void kt_build_w(cudaStream_t s, int g0, int cnt, const u32* gkeys, u32* proj_scratch, u8* kvalid, u32* tab) {
        typedef Bodies<C, W> B;
        k_kt_base<C, W><<<grid(cnt), BLK, 0, s>>>(g0, cnt, gkeys, proj_scratch, kvalid);
        count_launch();
        const long items = (long)cnt * B::KT_WINDOWS;
        u32* t0 = tab + (size_t)g0 * B::KT_KEY_WORDS;
        {
            int ept = (int)((items + 148L * BLK - 1) / (148L * BLK));
            if (ept < 1) ept = 1;
            if (ept > B::EPT) ept = B::EPT;
            const int threads = (int)((items + ept - 1) / ept);
            k_normalize<C><<<grid(threads), BLK, 0, s>>>((int)items, proj_scratch, NORM_AFF_STRIDED, B::KT_E * 2 * C::L, nullptr, nullptr, t0);
            count_launch();
        }
        {
            // items per thread: whole waves of resident CTAs with equal shares (2^16 secp256k1 keys = 1.7 M items: one wave
            // of 23 items per thread instead of 1.4 waves of 16), at most KT_EPT
            const long wave = 148L * ktf_min_ctas<C>() * BLK;
            const long waves = (items + wave * B::KT_EPT - 1) / (wave * B::KT_EPT);
            long ept = (items + waves * wave - 1) / (waves * wave);
            if (ept < 1) ept = 1;
            if (ept > B::KT_EPT) ept = B::KT_EPT;
            const int threads = (int)((items + ept - 1) / ept);
            k_kt_fill<C, W><<<grid(threads), BLK, 0, s>>>((int)items, t0);
            count_launch();
        }
    }
    static void verify_keytab(cudaStream_t s, int wide, int n, int mode, const u8* rs, const u8* z, const u32* scratch, const int* gid, const u8* kvalid,
                              const u32* tab, const u32* gbig, int gw, u8* ok) {
        if (n <= 0) return;
        if constexpr (kt_two_widths<C>()) {
            if (!wide) { verify_keytab_w<kt_narrow<C>()>(s, n, mode, rs, z, scratch, gid, kvalid, tab, gbig, gw, ok); return; }
        }
        verify_keytab_w<0>(s, n, mode, rs, z, scratch, gid, kvalid, tab, gbig, gw, ok);
    }
    template <int W> static void verify_keytab_w(cudaStream_t s, int n, int mode, const u8* rs, const u8* z, const u32* scratch, const int* gid, const u8* kvalid,
                                                 const u32* tab, const u32* gbig, int gw, u8* ok) {
        if (mode == VM_SM2DSA) {
            if constexpr (C::ID == 3) k_verify_keytab<C, VM_SM2DSA, W><<<grid(n), BLK, 0, s>>>(n, rs, z, scratch, gid, kvalid, tab, gbig, gw, ok);
        } else {
            k_verify_keytab<C, VM_ECDSA, W><<<grid(n), BLK, 0, s>>>(n, rs, z, scratch, gid, kvalid, tab, gbig, gw, ok);
        }
        count_launch();
    }
    static void decode(cudaStream_t s, int n, int mode, const u8* enc, int stride, u8* xy, u8* status) {
        if (n <= 0) return;
        k_decode<C><<<grid(n), BLK, 0, s>>>(n, mode, enc, stride, xy, status);
        count_launch();
    }
    static void finish(cudaStream_t s, int n, int kind, const u8* a, int stride, const u8* inf, const u8* rs, u8* ok) {
        if (n <= 0) return;
        k_finish<C><<<grid(n), BLK, 0, s>>>(n, kind, a, stride, inf, rs, ok);
        count_launch();
    }
    static void sign_finish(cudaStream_t s, int n, const u8* d, const u8* k, const u8* z, const u32* aff, u8* rs_out, u8* recid_out, u8* ok_out) {
        if (n <= 0) return;
        k_sign_finish<C><<<grid(ept_threads(n)), BLK, 0, s>>>(n, d, k, z, aff, rs_out, recid_out, ok_out);
        count_launch();
    }
    static CurveLaunch make_table() {
        CurveLaunch t = {
            C::ID, C::L, C::FB, C::A_IS_ZERO ? 8 : 15, Bodies<C>::GEN_WINDOWS, 8, C::COMPRESS_DEFAULT, {0},
            &field_op, &mul_var, &mul_gen, &load_proj, &normalize, &sum, &proj_to_bytes, &verify,
            &mul_var_fast, &verify_prep, &verify_main, &decode, &finish, &sign_finish, &wintab, &add_proj,
            {Bodies<C, kt_narrow<C>()>::KT_WINDOWS, Bodies<C>::KT_WINDOWS}, {Bodies<C, kt_narrow<C>()>::KT_KEY_WORDS, Bodies<C>::KT_KEY_WORDS},
            {kt_narrow<C>(), Bodies<C>::KT_W_DEFAULT}, Bodies<C>::KBW, &kt_group, &kt_build, &verify_keytab, Bodies<C>::PREP_WORDS, SUM_BLOCKS,
            Bodies<C>::G2_W, Bodies<C>::G2_WINDOWS, Bodies<C>::G2_E, &mul_gen2, &sum_normalize};
        for (int i = 0; i < C::L; i++) {   // R mod n (the Montgomery "one" of the scalar field) as big-endian bytes
            const u32 w = C::Fn::Params::one(i);
            t.r_mod_n[C::FB - 4 * i - 1] = (u8)w; t.r_mod_n[C::FB - 4 * i - 2] = (u8)(w >> 8);
            t.r_mod_n[C::FB - 4 * i - 3] = (u8)(w >> 16); t.r_mod_n[C::FB - 4 * i - 4] = (u8)(w >> 24);
        }
        return t;
    }
    static const CurveLaunch* table() {
        static const CurveLaunch t = make_table();   // initialised once, thread-safe (C++11 magic static)
        return &t;
    }
};

}  // namespace ecb
