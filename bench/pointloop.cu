// pointloop.cu — microbenchmark of the public-input inner loop (4 Jacobian doublings + 1 mixed addition per
// round) under different call strategies, selected at compile time:
//   -DECB_POINT_FN='__device__ __noinline__'     point operations as real functions (arguments in local memory)
//   -DECB_POINT_FN='__device__ __forceinline__'  point operations inlined into the loop (accumulator in registers)
//   -DECB_FIELD_FN=...                           same choice for the field multiplier / squarer
// Build + run: see bench/run_pointloop.sh.  Prints ms and point-rounds/s per curve.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../rustcrypto-elliptic-curves_b200/csrc/jac.cuh"

using namespace ecb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#ifndef MINCTAS
#define MINCTAS 4
#endif

template <class C> __global__ void __launch_bounds__(128, MINCTAS) k_loop(u32* io, int rounds) {
    typedef Jac<C> JJ;
    constexpr int L = C::L;
    typename JJ::J acc;
    typename JJ::A e;
    u32* p = io + (size_t)(blockIdx.x * 128 + threadIdx.x) * 5 * L;
#pragma unroll
    for (int i = 0; i < L; i++) { acc.X.v[i] = p[i]; acc.Y.v[i] = p[L + i]; acc.Z.v[i] = p[2 * L + i]; e.x.v[i] = p[3 * L + i]; e.y.v[i] = p[4 * L + i]; }
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) JJ::dbl(acc, acc);
        JJ::madd(acc, acc, e, nullptr);
    }
#pragma unroll
    for (int i = 0; i < L; i++) { p[i] = acc.X.v[i]; p[L + i] = acc.Y.v[i]; p[2 * L + i] = acc.Z.v[i]; }
}

template <class C> void run(const char* name, int rounds) {
    constexpr int L = C::L;
    const int ctas = 148 * MINCTAS * 4, n = ctas * 128;
    u32* d;
    size_t words = (size_t)n * 5 * L;
    u32* h = (u32*)malloc(words * 4);
    unsigned s = 12345;
    for (size_t i = 0; i < words; i++) { s = s * 1664525u + 1013904223u; h[i] = s >> 1; }   // < 2^31 per limb: below every modulus
    CK(cudaMalloc(&d, words * 4));
    CK(cudaMemcpy(d, h, words * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_loop<C><<<ctas, 128>>>(d, 2);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_loop<C><<<ctas, 128>>>(d, rounds);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
    printf("%-6s rounds=%d threads=%d: %.3f ms  %.2f G rounds/s  (x=%08x)\n", name, rounds, n, best, (double)n * rounds / best / 1e6, h[0]);
    cudaFree(d);
    free(h);
}

int main(int argc, char** argv) {
    int rounds = argc > 1 ? atoi(argv[1]) : 64;
    run<CurveK256>("k256", rounds);
    run<CurveP256>("p256", rounds);
    return 0;
}
