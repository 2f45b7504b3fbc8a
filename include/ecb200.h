/* ecb200.h — C ABI of the B200-native batched elliptic-curve engine.
 *
 * This is the drop-in boundary for the data-parallel hot path of the RustCrypto `elliptic-curves`
 * workspace (risc0 fork): independent scalar multiplications, batch normalisation, small linear
 * combinations and ECDSA verify_prehash over secp256k1 (k256) and the primeorder curves P-256,
 * P-384 and SM2 (and P-192 / P-224 through the same template).  The reference has no FFI of its own; its only accelerator precedent is the risc0
 * zkVM syscall `modmul_u256_denormalized(&U256,&U256,&U256)` (k256/src/arithmetic/field/
 * field_8x32_risc0.rs:177-193, k256/src/arithmetic/scalar.rs:114-134), which swaps ONE modular
 * multiplication.  Across PCIe that granularity is useless, so the boundary moves up to whole batches
 * of the trait-level operations; each entry point cites the reference item it replaces.
 *
 * Conventions
 *   - All integers cross the boundary as big-endian byte strings of FB bytes (FB = 32 for
 *     K256/P256/SM2, 48 for P384, 24 for P192, 28 for P224), arrays of structures, one element after another — the layout of
 *     `FieldBytes` / `Scalar::to_bytes()` / `EncodedPoint` coordinates in the reference.
 *   - Caller owns every buffer; the library copies and never retains.  Host entry points take host
 *     pointers (pageable or pinned) and return when the results are in the output buffers.
 *     `_dev` entry points take device pointers on the context's device and enqueue on `stream`
 *     (a cudaStream_t cast to void*; NULL = the context's own stream) without synchronising.
 *   - Return value: 0 on success, negative ecb200_status otherwise (ecb200_last_error gives text).
 *     Per-element failures are data, not status (ok[i] / invalid[i]), mirroring
 *     `Result<(), signature::Error>` and `CtOption`.
 *   - A context made by ecb200_init is bound to one CUDA device and is thread-compatible (one call at a time; two
 *     contexts may be driven from two host threads concurrently).  A context made by ecb200_init_multi spans several
 *     devices of the box: its host entry points cut [0, n) into contiguous index shards, one per device, and run them
 *     concurrently on one host thread + streams + pinned staging per device - no collective on the data path.
 *     The alternative deployment is one process per GPU with one single-device context each (bench.py under torchrun).
 *   - There is no CPU fallback: every entry point fails with ECB200_ERR_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef ECB200_H
#define ECB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecb200_ctx ecb200_ctx;

typedef enum {
    ECB200_K256 = 0, /* secp256k1  — k256/src/lib.rs:76-111 */
    ECB200_P256 = 1, /* NIST P-256 — p256/src/lib.rs:74-120 */
    ECB200_P384 = 2, /* NIST P-384 — p384/src/lib.rs:50-76 */
    ECB200_SM2 = 3,  /* SM2        — sm2/src/lib.rs:60-85 */
    ECB200_P192 = 4, /* NIST P-192 — p192/src/lib.rs:42-66 (the primeorder template on 24-byte fields) */
    ECB200_P224 = 5  /* NIST P-224 — p224/src/lib.rs (28-byte fields, 7 limbs, p = 1 mod 4) */
} ecb200_curve;

typedef enum {
    ECB200_OK = 0,
    ECB200_ERR_ARG = -1,   /* null pointer, bad curve id, negative count */
    ECB200_ERR_CUDA = -2,  /* CUDA runtime failure (no device, launch or copy error) */
    ECB200_ERR_ALLOC = -3, /* out of device or pinned memory */
    ECB200_ERR_POINT = -4  /* ecb200_lincomb: a term is not a valid point (coordinate >= p or off the curve); the reference
                              cannot even construct such a ProjectivePoint, so the sum is refused rather than computed without it */
} ecb200_status;

/* flags */
#define ECB200_FLAG_CT 1u           /* secret-scalar path: fixed windows, full table scans, no secret-dependent
                                       branch or address (mirrors k256 mul.rs:92-127, primeorder projective.rs:127-147) */
#define ECB200_FLAG_COMPRESSED 2u   /* force 02/03||x output */
#define ECB200_FLAG_UNCOMPRESSED 4u /* force 04||x||y output; neither flag = the reference default for the curve
                                       (k256 compressed, k256/src/lib.rs:108-111; others uncompressed) */
#define ECB200_FLAG_PROJ 8u         /* input points are X||Y||Z homogeneous projective (x = X/Z), 3*FB bytes each */

/* Field bytes of a curve (24, 28, 32 or 48); 0 for an unknown curve. */
size_t ecb200_field_bytes(int curve);
/* Output slot size for one encoded point under `flags`: 1+FB (compressed) or 1+2*FB.
 * A slot holds the SEC1 encoding; the identity is an all-zero slot (tag 00), which is also
 * GroupEncoding::to_bytes of the identity (k256/src/arithmetic/affine.rs:233-238). */
size_t ecb200_point_slot_bytes(int curve, uint32_t flags);

/* Create / destroy a context on CUDA device `device`.  Builds the fixed-base tables on the device. */
int ecb200_init(int device, ecb200_ctx** out);
/* One context over n_dev devices ("one call shards [0, n) over the GPUs of the box").  devices = NULL
 * means devices 0 .. n_dev-1; n_dev <= 0 (with devices = NULL) means every visible device; an ordinal may be listed
 * twice (two shards on one device - how the single-GPU tests exercise the sharding).  Every host-pointer entry
 * point below accepts such a context: rows [floor(i n / g), floor((i+1) n / g)) go to device i, each on its own host
 * thread, streams and pinned staging; ecb200_lincomb reduces each shard to one projective partial on its device and
 * adds the <= g partials on the first device.  `_dev` entry points (device pointers) need a single-device context. */
int ecb200_init_multi(int n_dev, const int* devices, ecb200_ctx** out);
/* Number of devices behind a context (1 for ecb200_init). */
int ecb200_device_count(const ecb200_ctx* ctx);
void ecb200_destroy(ecb200_ctx* ctx);
const char* ecb200_last_error(const ecb200_ctx* ctx);
const char* ecb200_version(void);
/* Number of kernels this context has launched so far (bench.py's gpu_launches evidence). */
uint64_t ecb200_launch_count(const ecb200_ctx* ctx);
/* Verification statistics since the context was created: rows verified on per-key window tables and tables built.
 * ecb200_ecdsa_verify* / ecb200_sm2dsa_verify* group the rows of a call by public key; when keys repeat (at least 4 rows per
 * distinct key) each key's window multiples v * 2^(W w) * Q are computed once per call and every row needs additions only - no doublings.
 * Calls whose keys do not repeat take the per-row path; results are identical either way.  Nothing is cached between calls.
 * The window width W follows the reuse: 4 bits below 20 rows per key, the curve's wide tables (6 bits on secp256k1, 5 on P-256)
 * from there on.  Environment (read by ecb200_init, for tests and A/B measurements): ECB200_KEYTAB=0 forces the per-row path,
 * ECB200_KT_WIDE=0/1 the narrow / wide tables, ECB200_GW=<bits> the window width of the fixed-base table of u1*G (default 20
 * bits = 410 MB per 256-bit curve, built on the first verification).  A `_dev` verification call on the table path
 * synchronises its stream once (the number of distinct keys comes back to the host). */
int ecb200_keytab_stats(const ecb200_ctx* ctx, uint64_t* rows, uint64_t* tables);
/* Block until everything enqueued on the context's stream has finished. */
int ecb200_sync(ecb200_ctx* ctx);
/* Measurement hook (bench.py roofline): while enabled, every launch of the dominant kernel of an operation (the main
 * kernel of ecb200_ecdsa_verify*, the scalar-multiplication kernel of ecb200_mul_var* / ecb200_mul_gen*) is bracketed by
 * CUDA events on the stream it is launched on; _read waits for them and returns the summed duration and the number of
 * launches since the last read. */
int ecb200_kernel_timing(ecb200_ctx* ctx, int enable);
int ecb200_kernel_timing_read(ecb200_ctx* ctx, double* total_ms, uint64_t* launches);

/* out[i] = SEC1(k[i] * G).  Replaces `ProjectivePoint::mul_by_generator(&k)` + `to_encoded_point`
 * (k256/src/arithmetic/mul.rs:415-440, primeorder/src/projective.rs:422-431; k256 affine.rs:272-284).
 * k: n x FB scalars, reduced mod n once like Reduce<Uint>::reduce_bytes (k256 scalar.rs:700-713). */
int ecb200_mul_gen(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* k, uint8_t* out, uint32_t flags);

/* out[i] = SEC1(k[i] * P[i]).  Replaces `&P * &k` (k256 mul.rs:443-481 -> lincomb N=1;
 * primeorder/src/projective.rs:106-150) followed by batch normalisation.
 * pts: n x 2FB (x||y) affine, or n x 3FB with ECB200_FLAG_PROJ; inf: optional n bytes, non-zero marks
 * the affine identity (ignored with FLAG_PROJ, where Z = 0 is the identity); invalid: optional n bytes,
 * set to 1 where a point failed validation (coordinate >= p or off-curve; its output is the identity
 * slot) — the analogue of a failed `AffinePoint::from_encoded_point` (k256 affine.rs:241-270). */
int ecb200_mul_var(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* pts, const uint8_t* inf, const uint8_t* k,
                   uint8_t* out, uint8_t* invalid, uint32_t flags);

/* (X:Y:Z) -> (x, y).  Replaces `BatchNormalize<[ProjectivePoint]>::batch_normalize` /
 * `group::Curve::batch_normalize` (k256/src/arithmetic/projective.rs:325-379,519-525;
 * primeorder/src/projective.rs:346-413).  xyz: n x 3FB; xy: n x 2FB; inf: n bytes (1 where Z == 0,
 * xy zeroed = AffinePoint::IDENTITY). */
int ecb200_batch_normalize(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* xyz, uint8_t* xy, uint8_t* inf);

/* out_point = sum_i k[i] * P[i] (ONE point).  Replaces `LinearCombinationExt::lincomb_ext(&[(P, k)])`
 * and `LinearCombination::lincomb` (k256 mul.rs:313-393; primeorder/src/projective.rs:415-420).
 * pts as in ecb200_mul_var; with ECB200_FLAG_PROJ in `out_flags_proj` the result is written as
 * X||Y||Z (3FB bytes, a partial sum another rank can keep adding to), else as one SEC1 slot.
 * A term that is not a valid point fails the whole call with ECB200_ERR_POINT (nothing is written). */
int ecb200_lincomb(ecb200_ctx* ctx, int curve, size_t n_terms, const uint8_t* pts, const uint8_t* k, uint8_t* out_point,
                   uint32_t flags, uint32_t out_flags_proj);

/* out[i] = SEC1(k1[i] * P1[i] + k2[i] * P2[i]) - one result PER ROW.  Replaces `LinearCombination::lincomb(&x, &k, &y, &l)`
 * (k256/src/arithmetic/mul.rs:313-323 -> lincomb N=2; primeorder/src/projective.rs:415-420) called in a loop over
 * slices.  p1, p2: n x 2FB affine x||y, or n x 3FB X||Y||Z with ECB200_FLAG_PROJ (Z = 0 is the identity); k1, k2: n x FB;
 * invalid: optional n bytes, 1 where either point failed validation (that row's output is the identity slot).
 * ECB200_FLAG_CT selects the secret-scalar kernels for both products. */
int ecb200_lincomb2(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* p1, const uint8_t* k1, const uint8_t* p2,
                    const uint8_t* k2, uint8_t* out, uint8_t* invalid, uint32_t flags);

/* ok[i] = 1 iff the signature verifies.  Replaces `VerifyingKey::verify_prehash` after bits2field, i.e.
 * `<AffinePoint as VerifyPrimitive>::verify_prehashed(&Q, &z, &sig)` (k256/src/ecdsa.rs:200-209 with the
 * low-s rule; p256/src/ecdsa.rs:71-75, p384/src/ecdsa.rs default impl; body in ecdsa 0.16.9
 * hazmat::verify_prehashed).  q: n x 2FB public keys x||y; z: n x FB prehash after bits2field;
 * rs: n x 2FB r||s.  Keys off the curve / coordinates >= p and r, s outside [1, n-1] give ok = 0
 * (the reference rejects them when the key / signature is constructed). */
int ecb200_ecdsa_verify(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* q, const uint8_t* z, const uint8_t* rs,
                        uint8_t* ok);

/* Test hook: out[i] = a[i] (op) b[i] in the base field (which = 0) or scalar field (which = 1).
 * op: 0 add, 1 sub, 2 mul, 3 square, 4 negate, 5 invert (0 -> 0), 6 sqrt (base field).
 * ok[i] = 0 when an input is >= the modulus (from_repr failure) or the root does not exist.
 * Mirrors the per-op surface the risc0 backend replaces (field_8x32_risc0.rs:139-193) so the
 * reference's field KATs can be replayed on the device. */
int ecb200_field_op(ecb200_ctx* ctx, int curve, int which, int op, size_t n, const uint8_t* a, const uint8_t* b,
                    uint8_t* out, uint8_t* ok);

/* Device-pointer variants: same semantics, buffers already resident in HBM on the context's device,
 * work enqueued on `stream` (NULL = context stream), no synchronisation.  Scratch memory is owned by
 * the context and reused, so calls on one context must be stream-ordered. */
int ecb200_mul_gen_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_k, uint8_t* d_out, uint32_t flags,
                       void* stream);
int ecb200_mul_var_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_pts, const uint8_t* d_inf,
                       const uint8_t* d_k, uint8_t* d_out, uint8_t* d_invalid, uint32_t flags, void* stream);
int ecb200_batch_normalize_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_xyz, uint8_t* d_xy,
                               uint8_t* d_inf, void* stream);
int ecb200_ecdsa_verify_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_q, const uint8_t* d_z,
                            const uint8_t* d_rs, uint8_t* d_ok, void* stream);
int ecb200_lincomb2_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_p1, const uint8_t* d_k1,
                        const uint8_t* d_p2, const uint8_t* d_k2, uint8_t* d_out, uint8_t* d_invalid, uint32_t flags,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Callers and data formats either side of the path.  Same conventions;
 * every entry point has a `_dev` twin taking device pointers and a stream.
 */

/* decode modes of ecb200_decode_points */
#define ECB200_DECODE_SEC1 0u    /* tag 02/03 + x, tag 05 + x (compact), tag 04 + x + y, all-zero slot = identity */
#define ECB200_DECODE_COMPACT 1u /* x only: DecompactPoint::decompact - the even root on secp256k1 (BIP340 x-only keys,
                                    k256 affine.rs:204-211), the root with the smaller y on the primeorder curves
                                    (primeorder/src/affine.rs:66-77,148-156) */

/* xy[i] = the affine point encoded in slot i; status[i] = 1 point, 2 identity, 0 invalid (xy zeroed unless 1).
 * Replaces `AffinePoint::from_encoded_point` / `DecompressPoint::decompress` / `DecompactPoint::decompact`
 * (k256/src/arithmetic/affine.rs:184-211,241-270; primeorder/src/affine.rs:129-195) incl. the square root
 * (k256 field.rs:220-255, p256 field.rs:385-411, p384 field.rs:95-117).  enc: n slots of `stride` bytes
 * (stride >= 1+FB for compressed-only input, >= 1+2FB if tag 04 may occur; FB for ECB200_DECODE_COMPACT). */
int ecb200_decode_points(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* enc, size_t stride, uint32_t mode,
                         uint8_t* xy, uint8_t* status);

/* ecb200_ecdsa_verify with SEC1-encoded public keys (`VerifyingKey::from_sec1_bytes` + `verify_prehash`):
 * keys are decoded (decompressed) on the device; an undecodable or identity key gives ok = 0. */
int ecb200_ecdsa_verify_sec1(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* keys, size_t key_stride,
                             const uint8_t* z, const uint8_t* rs, uint8_t* ok);

/* keys[i] = SEC1 slot of the public key recovered from (z, r||s, recid); ok[i] = 0 where recovery fails.
 * Replaces `VerifyingKey::recover_from_prehash` (ecdsa 0.16.9 recovery.rs; call sites and vectors
 * k256/src/ecdsa.rs:113-140,278-343).  recid: one byte per row, bit 0 = y(R) odd, bit 1 = x(R) was reduced.
 * Failure cases as in the reference: r, s out of range, r + n overflow / >= p, no square root, identity key,
 * and (k256) a high-s signature, which the closing `verify_prehash` of the reference rejects. */
int ecb200_ecdsa_recover(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* z, const uint8_t* rs,
                         const uint8_t* recid, uint8_t* keys, uint8_t* ok, uint32_t flags);

/* BIP340 Schnorr verification over secp256k1 after hashing.  Replaces the arithmetic of
 * `schnorr::VerifyingKey::verify_prehash` (k256/src/schnorr/verifying.rs:63-89) plus key / signature parsing
 * (verifying.rs:35-45, schnorr.rs:143-160).  pk: n x 32 x-only keys; e: n x 32 challenge digests
 * tagged_hash("BIP0340/challenge", r || pk || msg) (hashing stays with the caller, like the ECDSA prehash;
 * reduced mod n on the device); sig: n x 64 r||s.  ok = 1 iff lift_x(pk) exists, 0 < r < p, 1 <= s < n and
 * R = s*G - e*P is finite with even y and x(R) = r. */
int ecb200_schnorr_verify(ecb200_ctx* ctx, size_t n, const uint8_t* pk, const uint8_t* e, const uint8_t* sig,
                          uint8_t* ok);

/* SM2DSA verification after hashing.  Replaces `sm2::dsa::VerifyingKey::verify_prehash`
 * (sm2/src/dsa/verifying.rs:130-168): q: n x 64 keys x||y; e: n x 32 digests SM3(Z_A || M); rs: n x 64.
 * t = r + s (t = 0 rejected), (x1, y1) = s*G + t*P, accept iff r == e + x1 (mod n). */
int ecb200_sm2dsa_verify(ecb200_ctx* ctx, size_t n, const uint8_t* q, const uint8_t* e, const uint8_t* rs,
                         uint8_t* ok);

/* ECDSA signing with caller-supplied nonces (RFC 6979 derivation is HMAC work and stays on the host).
 * Replaces `SignPrimitive::try_sign_prehashed` (ecdsa 0.16.9 hazmat::sign_prehashed; k256/src/ecdsa.rs:181-198
 * adds low-s normalisation and the parity flip of the recovery id).  d, k: n x FB secret scalars; z: n x FB
 * prehash after bits2field; rs: n x 2FB; recid: n bytes; ok[i] = 0 (outputs zeroed) when d or k is 0 or >= n, or
 * r or s is 0.  Constant-time discipline: fixed-base multiplication with full table scans, complete formulas,
 * masked selections; no branch or address depends on d or k. */
int ecb200_ecdsa_sign(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d, const uint8_t* k, const uint8_t* z,
                      uint8_t* rs, uint8_t* recid, uint8_t* ok);

int ecb200_decode_points_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_enc, size_t stride,
                             uint32_t mode, uint8_t* d_xy, uint8_t* d_status, void* stream);
int ecb200_ecdsa_verify_sec1_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_keys, size_t key_stride,
                                 const uint8_t* d_z, const uint8_t* d_rs, uint8_t* d_ok, void* stream);
int ecb200_ecdsa_recover_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_z, const uint8_t* d_rs,
                             const uint8_t* d_recid, uint8_t* d_keys, uint8_t* d_ok, uint32_t flags, void* stream);
int ecb200_schnorr_verify_dev(ecb200_ctx* ctx, size_t n, const uint8_t* d_pk, const uint8_t* d_e,
                              const uint8_t* d_sig, uint8_t* d_ok, void* stream);
int ecb200_sm2dsa_verify_dev(ecb200_ctx* ctx, size_t n, const uint8_t* d_q, const uint8_t* d_e, const uint8_t* d_rs,
                             uint8_t* d_ok, void* stream);
int ecb200_ecdsa_sign_dev(ecb200_ctx* ctx, int curve, size_t n, const uint8_t* d_d, const uint8_t* d_k,
                          const uint8_t* d_z, uint8_t* d_rs, uint8_t* d_recid, uint8_t* d_ok, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECB200_H */
