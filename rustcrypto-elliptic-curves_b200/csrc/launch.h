// launch.h — per-curve launcher table: the only interface between the host runtime (abi.cu) and
// the curve translation units (curve_*.cu), so each curve compiles in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ecb {

struct CurveLaunch {
    int id;
    int L;        // 32-bit limbs per field element
    int FB;       // field bytes
    int ngtab;    // entries of the small affine generator table used by verify (8 or 15)
    int gen_windows, gen_entries;   // fixed-base table shape: (8L + 1) windows x 8 entries
    bool compress_default;
    uint8_t r_mod_n[48];            // 2^(32L) mod n, big-endian FB bytes: the scalar of the top fixed-base window

    void (*field_op)(cudaStream_t s, int n, int which, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint8_t* ok);
    void (*mul_var)(cudaStream_t s, bool ct, int n, uint32_t flags, const uint8_t* pts, const uint8_t* inf,
                    const uint8_t* k, uint32_t* proj, uint8_t* invalid);
    void (*mul_gen)(cudaStream_t s, bool ct, int n, const uint8_t* k, const uint32_t* gentab, uint32_t* proj);
    void (*load_proj)(cudaStream_t s, int n, const uint8_t* xyz, uint32_t* proj, uint8_t* invalid);
    void (*normalize)(cudaStream_t s, int n, const uint32_t* proj, int mode, int compress, uint8_t* out_bytes,
                      uint8_t* out_inf, uint32_t* out_limbs);
    // sum n projective points into out (3L limbs); partial = scratch of sum_blocks()*3L limbs
    void (*sum)(cudaStream_t s, int n, const uint32_t* proj, uint32_t* partial, uint32_t* out);
    void (*proj_to_bytes)(cudaStream_t s, int n, const uint32_t* proj, uint8_t* xyz);
    void (*verify)(cudaStream_t s, int n, const uint8_t* q, const uint8_t* z, const uint8_t* rs, const uint32_t* gtab,
                   uint8_t* ok);
    // fast public-input path (jac.cuh)
    // wtab: per-row affine window tables made by `wintab` (primeorder curves), or NULL = per-thread Jacobian tables
    void (*mul_var_fast)(cudaStream_t s, int n, const uint8_t* pts, const uint32_t* aff_limbs, const uint8_t* inf,
                         const uint8_t* k, uint32_t* proj, uint8_t* invalid, const uint32_t* wtab);
    // mode = VM_* (kernels.cuh): ECDSA verify, SM2DSA verify, BIP340 Schnorr verify, ECDSA public-key recovery
    void (*verify_prep)(cudaStream_t s, int n, int mode, const uint8_t* z, const uint8_t* rs, uint32_t* scratch);
    void (*verify_main)(cudaStream_t s, int n, int mode, const uint8_t* q, const uint8_t* rs, const uint8_t* z, const uint8_t* aux,
                        const uint32_t* scratch, const uint32_t* gbig, int gw, uint8_t* ok, uint32_t* proj_out, const uint32_t* wtab);
    void (*decode)(cudaStream_t s, int n, int mode, const uint8_t* enc, int stride, uint8_t* xy, uint8_t* status);
    void (*finish)(cudaStream_t s, int n, int kind, const uint8_t* a, int stride, const uint8_t* inf, const uint8_t* rs, uint8_t* ok);
    void (*sign_finish)(cudaStream_t s, int n, const uint8_t* d, const uint8_t* k, const uint8_t* z, const uint32_t* aff,
                        uint8_t* rs_out, uint8_t* recid_out, uint8_t* ok_out);
    // affine window tables {1..8}Q of n points (x||y bytes or field-internal limbs) into wtab (23 L words per row: 16 L of
    // table, 7 L of scratch); no-op on secp256k1
    void (*wintab)(cudaStream_t s, int n, const uint8_t* pts, const uint32_t* aff_limbs, uint32_t* wtab);
    // out[i] = a[i] + b[i] on projective limbs (complete addition), invalid[i] |= invalid_b[i]: tail of the per-row 2-term lincomb
    void (*add_proj)(cudaStream_t s, int n, const uint32_t* a, const uint32_t* b, uint32_t* out, uint8_t* invalid, const uint8_t* invalid_b);
    // per-key window tables of the verify path (kernels.cuh "per-key window tables"): group the rows of a chunk by public key,
    // build the tables of groups [g0, g0 + cnt), verify on the tables
    // two table widths per curve, index 0 = narrow (calls with few rows per key), 1 = wide (the default); equal where a curve
    // has only one
    int kt_windows[2], kt_key_words[2], kt_w[2], kt_kbw;
    void (*kt_group)(cudaStream_t s, int n, const uint32_t* q32, int* htab, uint32_t hmask, uint32_t* gkeys, int* gid, int* rep, int* rep_slot,
                     int* newgid, int* counter, int cap);
    void (*kt_build)(cudaStream_t s, int wide, int g0, int cnt, const uint32_t* gkeys, uint32_t* proj_scratch, uint8_t* kvalid, uint32_t* tab);
    void (*verify_keytab)(cudaStream_t s, int wide, int n, int mode, const uint8_t* rs, const uint8_t* z, const uint32_t* scratch, const int* gid,
                          const uint8_t* kvalid, const uint32_t* tab, const uint32_t* gbig, int gw, uint8_t* ok);
    int prep_words;   // u32 words of scratch per row between verify_prep and verify_main
    int sum_blocks;
    // split fixed-base path (kernels.cuh body_gen_half): table of gen2_windows x gen2_entries affine entries
    // v * 2^(gen2_w * w) * G; mul_gen2 leaves the two half sums in part (2 x n x 3L limbs), sum_normalize adds and normalises
    int gen2_w, gen2_windows, gen2_entries;
    void (*mul_gen2)(cudaStream_t s, bool ct, int n, const uint8_t* k, const uint32_t* tab2, uint32_t* part);
    void (*sum_normalize)(cudaStream_t s, int n, uint32_t* proj, const uint32_t* sum_with, int mode, int compress, uint8_t* out_bytes,
                          uint8_t* out_inf, uint32_t* out_limbs);
};

const CurveLaunch* launch_k256();
const CurveLaunch* launch_p256();
const CurveLaunch* launch_p384();
const CurveLaunch* launch_sm2();
const CurveLaunch* launch_p192();
const CurveLaunch* launch_p224();

// Launch accounting (bench.py's gpu_launches evidence).  Every entry point of abi.cu points this thread-local at the
// calling context's own counter before it enqueues anything, so two contexts driven from two host threads (one per
// device under ecb200_init_multi) never share a counter.
extern thread_local uint64_t* tl_launch_counter;
inline void count_launch(int k = 1) { if (tl_launch_counter) *tl_launch_counter += (uint64_t)k; }

}  // namespace ecb
