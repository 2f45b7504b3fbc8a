/* osslref.c — TEST / BENCH INFRASTRUCTURE: batch drivers over OpenSSL libcrypto (EC_POINT_mul, ECDSA_do_verify) with
 * OpenMP over the host cores.  Second, independent CPU oracle and the "production CPU library" figure of BASELINE.md §3;
 * never linked into or loaded by the product.  Semantics aligned with the reference: scalars reduced once mod n,
 * r, s outside [1, n-1] rejected, optional low-s rule (k256/src/ecdsa.rs:203-205), keys that fail the curve check rejected. */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/ecdsa.h>
#include <openssl/err.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ossl_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ossl_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* out[i] = SEC1 slot (1+fb compressed / 1+2fb) of k[i]*G (pts == NULL) or k[i]*P[i]; identity or invalid point -> zero slot */
int ossl_mul(int nid, int fb, size_t n, const uint8_t* pts, const uint8_t* k, int compress, uint8_t* out) {
    EC_GROUP* g = EC_GROUP_new_by_curve_name(nid);
    if (!g) return -1;
    const size_t slot = 1 + (size_t)(compress ? fb : 2 * fb);
    const BIGNUM* order = EC_GROUP_get0_order(g);
#pragma omp parallel
    {
        BN_CTX* ctx = BN_CTX_new();
        EC_POINT* P = EC_POINT_new(g);
        EC_POINT* R = EC_POINT_new(g);
        BIGNUM *bk = BN_new(), *bx = BN_new(), *by = BN_new();
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            uint8_t* o = out + (size_t)i * slot;
            memset(o, 0, slot);
            BN_bin2bn(k + (size_t)i * fb, fb, bk);
            if (BN_cmp(bk, order) >= 0) BN_sub(bk, bk, order);
            int ok;
            if (!pts) ok = EC_POINT_mul(g, R, bk, NULL, NULL, ctx);
            else {
                BN_bin2bn(pts + (size_t)i * 2 * fb, fb, bx);
                BN_bin2bn(pts + (size_t)i * 2 * fb + fb, fb, by);
                ok = EC_POINT_set_affine_coordinates(g, P, bx, by, ctx) && EC_POINT_mul(g, R, NULL, P, bk, ctx);
            }
            if (!ok) { ERR_clear_error(); continue; }
            if (EC_POINT_is_at_infinity(g, R)) continue;
            EC_POINT_point2oct(g, R, compress ? POINT_CONVERSION_COMPRESSED : POINT_CONVERSION_UNCOMPRESSED, o, slot, ctx);
        }
        BN_free(bk); BN_free(bx); BN_free(by);
        EC_POINT_free(P); EC_POINT_free(R);
        BN_CTX_free(ctx);
    }
    EC_GROUP_free(g);
    return 0;
}

/* ok[i] = verify_prehashed(Q_i, z_i, (r_i, s_i)) */
int ossl_verify(int nid, int fb, size_t n, const uint8_t* q, const uint8_t* z, const uint8_t* rs, int low_s, uint8_t* ok) {
    EC_GROUP* g = EC_GROUP_new_by_curve_name(nid);
    if (!g) return -1;
    const BIGNUM* order = EC_GROUP_get0_order(g);
    BIGNUM* half = BN_dup(order);
    BN_rshift1(half, half);
#pragma omp parallel
    {
        BN_CTX* ctx = BN_CTX_new();
        EC_POINT* P = EC_POINT_new(g);
        BIGNUM *bx = BN_new(), *by = BN_new();
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            ok[i] = 0;
            BIGNUM* r = BN_bin2bn(rs + (size_t)i * 2 * fb, fb, NULL);
            BIGNUM* s = BN_bin2bn(rs + (size_t)i * 2 * fb + fb, fb, NULL);
            int good = !BN_is_zero(r) && !BN_is_zero(s) && BN_cmp(r, order) < 0 && BN_cmp(s, order) < 0 && !(low_s && BN_cmp(s, half) > 0);
            if (good) {
                BN_bin2bn(q + (size_t)i * 2 * fb, fb, bx);
                BN_bin2bn(q + (size_t)i * 2 * fb + fb, fb, by);
                good = EC_POINT_set_affine_coordinates(g, P, bx, by, ctx);
            }
            if (!good) { ERR_clear_error(); BN_free(r); BN_free(s); continue; }
            EC_KEY* key = EC_KEY_new();
            EC_KEY_set_group(key, g);
            EC_KEY_set_public_key(key, P);
            ECDSA_SIG* sig = ECDSA_SIG_new();
            ECDSA_SIG_set0(sig, r, s);   /* takes ownership */
            int v = ECDSA_do_verify(z + (size_t)i * fb, fb, sig, key);
            if (v != 1) ERR_clear_error();
            ok[i] = v == 1;
            ECDSA_SIG_free(sig);
            EC_KEY_free(key);
        }
        BN_free(bx); BN_free(by);
        EC_POINT_free(P);
        BN_CTX_free(ctx);
    }
    BN_free(half);
    EC_GROUP_free(g);
    return 0;
}
