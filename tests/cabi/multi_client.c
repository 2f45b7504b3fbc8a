/* multi_client.c — a plain C caller driving SEVERAL devices from one process through one context (ecb200_init_multi):
 * the replacement for a `for ... verify_prehash` loop (k256/benches/ecdsa.rs:57-63) that reaches every GPU of the box with one
 * call.  argv[1] = number of shards (default: every visible device; on a one-GPU box pass 2 to get two shards on device 0).
 * Signs n rows on the multi-device context, verifies them on it, and checks every result against a single-device context.
 * Built and run by tests/test_gpu_round2.py::test_plain_c_multi_device_client (GPU tier). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ecb200.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "check failed at line %d: %s (%s)\n", __LINE__, #c, ecb200_last_error(multi)); return 1; } } while (0)

int main(int argc, char** argv) {
    const size_t n = 300007;                       /* not a multiple of anything: uneven shards */
    int shards = argc > 1 ? atoi(argv[1]) : 0;
    int devs[64] = {0};
    ecb200_ctx *multi = NULL, *one = NULL;
    if (ecb200_init(0, &one) != ECB200_OK) { fprintf(stderr, "no CUDA device: the engine has no CPU fallback\n"); return 2; }
    int rc;
    if (shards <= 0) rc = ecb200_init_multi(0, NULL, &multi);            /* every visible device */
    else {
        /* spread the shards over the visible devices round-robin (two shards on device 0 when there is one GPU) */
        ecb200_ctx* probe = NULL;
        int ndev = 0;
        if (ecb200_init_multi(0, NULL, &probe) == ECB200_OK) { ndev = ecb200_device_count(probe); ecb200_destroy(probe); }
        if (ndev <= 0) ndev = 1;
        for (int i = 0; i < shards; i++) devs[i] = i % ndev;
        rc = ecb200_init_multi(shards, devs, &multi);
    }
    if (rc != ECB200_OK) { fprintf(stderr, "ecb200_init_multi failed: %d\n", rc); return 1; }
    const int g = ecb200_device_count(multi);
    uint8_t *d = malloc(32 * n), *k = malloc(32 * n), *z = malloc(32 * n), *rs = malloc(64 * n), *rs1 = malloc(64 * n);
    uint8_t *rid = malloc(n), *ok = malloc(n), *ok1 = malloc(n), *pub = malloc(65 * n), *q = malloc(64 * n), *rid1 = malloc(n);
    unsigned s = 12345;
    for (size_t i = 0; i < 32 * n; i++) { s = s * 1664525u + 1013904223u; d[i] = (uint8_t)(s >> 24); s = s * 1664525u + 1013904223u; k[i] = (uint8_t)(s >> 24); s = s * 1664525u + 1013904223u; z[i] = (uint8_t)(s >> 24); }
    for (size_t i = 0; i < n; i++) { d[32 * i] &= 0x7F; d[32 * i + 31] |= 1; k[32 * i] &= 0x7F; k[32 * i + 31] |= 1; }
    CHECK(ecb200_ecdsa_sign(multi, ECB200_K256, n, d, k, z, rs, rid, ok) == ECB200_OK);
    CHECK(ecb200_ecdsa_sign(one, ECB200_K256, n, d, k, z, rs1, rid1, ok1) == ECB200_OK);
    CHECK(memcmp(rs, rs1, 64 * n) == 0 && memcmp(rid, rid1, n) == 0 && memcmp(ok, ok1, n) == 0);
    CHECK(ecb200_mul_gen(multi, ECB200_K256, n, d, pub, ECB200_FLAG_CT | ECB200_FLAG_UNCOMPRESSED) == ECB200_OK);
    for (size_t i = 0; i < n; i++) memcpy(q + 64 * i, pub + 65 * i + 1, 64);
    for (size_t i = 0; i < n; i += 7) rs[64 * i + 63] ^= 1;                 /* corrupt every 7th signature */
    CHECK(ecb200_ecdsa_verify(multi, ECB200_K256, n, q, z, rs, ok) == ECB200_OK);
    CHECK(ecb200_ecdsa_verify(one, ECB200_K256, n, q, z, rs, ok1) == ECB200_OK);
    CHECK(memcmp(ok, ok1, n) == 0);
    size_t good = 0;
    for (size_t i = 0; i < n; i++) { good += ok[i]; CHECK(ok[i] == (i % 7 != 0)); }
    /* many-term linear combination: per-device partial sums, added on the first device */
    uint8_t sum_m[33], sum_1[33];
    CHECK(ecb200_lincomb(multi, ECB200_K256, 5000, q, k, sum_m, 0, 0) == ECB200_OK);
    CHECK(ecb200_lincomb(one, ECB200_K256, 5000, q, k, sum_1, 0, 0) == ECB200_OK);
    CHECK(memcmp(sum_m, sum_1, 33) == 0 && sum_m[0] != 0);
    /* device pointers belong to one device: the _dev entry points refuse the multi-device context */
    CHECK(ecb200_mul_gen_dev(multi, ECB200_K256, 1, d, pub, 0, NULL) == ECB200_ERR_ARG);
    CHECK(ecb200_launch_count(multi) > 0);
    printf("cabi multi client ok: %d shards, %zu rows, %zu accepted\n", g, n, good);
    ecb200_destroy(multi);
    ecb200_destroy(one);
    return 0;
}
