/* client.c — a plain C caller of the C ABI (include/ecb200.h, libecb200.so): what a non-Python host (the Rust -sys crate of
 * INTEGRATION.md, cgo, JNI ...) links against.  Checks a few known answers on the device and prints "cabi client ok".
 * Built and run by tests/test_gpu_next_rows.py::test_plain_c_client (GPU tier). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ecb200.h"

static void unhex(uint8_t* out, const char* hex, size_t n) {
    for (size_t i = 0; i < n; i++) { unsigned v; sscanf(hex + 2 * i, "%2x", &v); out[i] = (uint8_t)v; }
}
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "check failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(void) {
    ecb200_ctx* ctx = NULL;
    if (ecb200_init(0, &ctx) != ECB200_OK) { fprintf(stderr, "no CUDA device: the engine has no CPU fallback\n"); return 2; }
    /* k * G for k = 1, 2 on secp256k1: compressed SEC1 of G and 2G (k256/src/test_vectors/group.rs) */
    uint8_t k[64] = {0}, out[66], exp[33];
    k[31] = 1; k[63] = 2;
    CHECK(ecb200_mul_gen(ctx, ECB200_K256, 2, k, out, ECB200_FLAG_CT) == ECB200_OK);
    unhex(exp, "0279BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798", 33);
    CHECK(memcmp(out, exp, 33) == 0);
    unhex(exp, "02C6047F9441ED7D6D3045406E95C07CD85C778E4B8CEF3CA7ABAC09B95C709EE5", 33);
    CHECK(memcmp(out + 33, exp, 33) == 0);
    /* decode the compressed generator back to x||y */
    uint8_t xy[64], st[1];
    CHECK(ecb200_decode_points(ctx, ECB200_K256, 1, out, 33, ECB200_DECODE_SEC1, xy, st) == ECB200_OK);
    CHECK(st[0] == 1);
    unhex(exp, "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8", 32);
    CHECK(memcmp(xy + 32, exp, 32) == 0);
    /* sign with d = 1, k = 2 over z = 3, verify it, tamper, verify again, recover the key */
    uint8_t d[32] = {0}, kk[32] = {0}, z[32] = {0}, rs[64], rid[1], ok[2];
    d[31] = 1; kk[31] = 2; z[31] = 3;
    CHECK(ecb200_ecdsa_sign(ctx, ECB200_K256, 1, d, kk, z, rs, rid, ok) == ECB200_OK && ok[0] == 1);
    uint8_t q2[128], z2[64], rs2[128];
    memcpy(q2, xy, 64); memcpy(q2 + 64, xy, 64);
    memcpy(z2, z, 32); memcpy(z2 + 32, z, 32);
    memcpy(rs2, rs, 64); memcpy(rs2 + 64, rs, 64);
    rs2[127] ^= 1;
    CHECK(ecb200_ecdsa_verify(ctx, ECB200_K256, 2, q2, z2, rs2, ok) == ECB200_OK);
    CHECK(ok[0] == 1 && ok[1] == 0);
    uint8_t key[33];
    CHECK(ecb200_ecdsa_recover(ctx, ECB200_K256, 1, z, rs, rid, key, ok, 0) == ECB200_OK && ok[0] == 1);
    CHECK(memcmp(key, out, 33) == 0);
    /* misuse is a status code, not a crash */
    CHECK(ecb200_mul_gen(ctx, 9, 1, k, out, 0) == ECB200_ERR_ARG);
    CHECK(strlen(ecb200_last_error(ctx)) > 0);
    CHECK(ecb200_launch_count(ctx) > 0);
    ecb200_destroy(ctx);
    printf("cabi client ok\n");
    return 0;
}
