// imad_peak.cu — measures the INT32 multiply-add issue peaks the roofline is quoted against
// (MEASURED_PEAKS.json has no integer figure).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/imad_peak bench/imad_peak.cu
//   ./bench/imad_peak > peaks_int.json
//
// For every instruction mix: one CTA of 1024 threads per SM (grid = #SMs x 2 waves resident as
// 2 CTAs of 512), UNROLL independent chains per thread, per-SM cycles from clock64, whole-chip
// time from CUDA events.  Reported: thread-level ops / clk / SM and chip-wide Gops/s.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../rustcrypto-elliptic-curves_b200/csrc/fp_k256.cuh"

using ecb::u32;
typedef unsigned long long u64;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;
constexpr int CH = 8;   // independent chains per thread

enum Mix { IMAD_LO, IMAD_HI, IMAD_WIDE, IMAD_WIDE_X, IADD3_, WIDE_PLUS_IADD3, LOHI_PAIR, FFMA_, DFMA_, WIDE_X_PLUS_2IADD3, WIDE_X_SHARED_B, MULWIDE8, FPMUL_K256, NMIX };
static const char* MIXNAME[NMIX] = {"imad_lo", "imad_hi", "imad_wide", "imad_wide_carry_chain", "iadd3", "imad_wide+iadd3",
                                    "imad_lo+imad_hi", "ffma", "dfma", "imad_wide_carry+2iadd3", "imad_wide_carry_chain_shared_multiplicand",
                                    "mul_wide8_macs(64/iter)", "fp_k256_mul_macs(73/iter)"};
// "ops" counted per loop body per chain
static const int MIXOPS[NMIX] = {1, 1, 1, 1, 1, 2, 2, 1, 1, 3, 1, 8, 0};

template <int MIX>
__global__ void __launch_bounds__(512, 2) k_mix(u32* sink, u64* cycles, u32 seed) {
    u32 a[CH], b[CH], c[CH], d[CH];
    u64 w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + 1 + threadIdx.x * 17; c[i] = seed * 11 + i + threadIdx.x; d[i] = seed * 13 + i + 9 + threadIdx.x * 29; w[i] = ((u64)c[i] << 32) | a[i]; }
    float f[CH]; double g[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { f[i] = (float)a[i]; g[i] = (double)a[i]; }
    __syncthreads();
    u64 t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH && MIX != MULWIDE8 && MIX != FPMUL_K256; i++) {
            if (MIX == IMAD_LO) a[i] = a[i] * b[i] + c[i];
            if (MIX == IMAD_HI) a[i] = __umulhi(a[i], b[i]) + c[i];
            if (MIX == IMAD_WIDE) w[i] = (u64)(u32)w[i] * b[i] + w[i];
            if (MIX == IMAD_WIDE_X) {   // the pattern the field multiplier emits: IMAD.WIDE.U32.X chains
                if (i == 0) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; mad.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
                else if (i < CH - 1) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
                else asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
            }
            if (MIX == WIDE_X_SHARED_B) {   // same multiplicand for the whole row, as in one row of the field multiplier
                if (i == 0) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; mad.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(d[i]), "r"(b[0]));
                else if (i < CH - 1) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(d[i]), "r"(b[0]));
                else asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(d[i]), "r"(b[0]));
            }
            if (MIX == IADD3_) a[i] = a[i] + b[i] + c[i];
            if (MIX == WIDE_PLUS_IADD3) {
                w[i] = (u64)(u32)w[i] * b[i] + w[i];
                d[i] = d[i] + b[i] + i;
            }
            if (MIX == LOHI_PAIR) {
                u32 t = a[i];
                a[i] = t * b[i] + c[i];
                c[i] = __umulhi(t, b[i]) + d[i];
            }
            if (MIX == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(1.0001f));
            if (MIX == DFMA_) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(g[i]) : "d"(1.0001));
            if (MIX == WIDE_X_PLUS_2IADD3) {
                if (i == 0) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; mad.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
                else if (i < CH - 1) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.cc.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
                else asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi}, %0; madc.lo.cc.u32 lo, %1, %2, lo; madc.hi.u32 hi, %1, %2, hi; mov.b64 %0, {lo,hi};}" : "+l"(w[i]) : "r"(b[i]), "r"(d[i]));
            }
        }
        if (MIX == MULWIDE8) {
            u32 r[16];
            ecb::mul_wide<8>(r, a, b);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = r[i] ^ r[i + 8];
        }
        if (MIX == FPMUL_K256) {
            ecb::FpK256::E x, y, z;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
            z = ecb::FpK256::mul_fn(x, y);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = z.v[i];
        }
        if (MIX == WIDE_X_PLUS_2IADD3) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(d[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(d[i]) : "r"(b[i]));
            }
        }
    }
    u64 t1 = clock64();
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= (u32)w[i] ^ (u32)(w[i] >> 32) ^ a[i] ^ c[i] ^ d[i] ^ b[i] ^ __float_as_uint(f[i]) ^ (u32)__double_as_longlong(g[i]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
static void run(int nsm, u32* sink, u64* cyc_d, double clk_mhz, bool last) {
    int grid = nsm * 2, block = 512;
    // warm up long enough for the SM clock to reach its boost state (a cold 0.3 ms launch runs at ~1.3 GHz)
    for (int r = 0; r < 100; r++) k_mix<MIX><<<grid, block>>>(sink, cyc_d, 1);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    const int REP = 300;   // >= 50 ms per mix: event-timed chip-wide rate at the sustained clock
    for (int r = 0; r < REP; r++) k_mix<MIX><<<grid, block>>>(sink, cyc_d, r + 2);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= REP;
    u64* cyc = (u64*)malloc(sizeof(u64) * grid);
    CK(cudaMemcpy(cyc, cyc_d, sizeof(u64) * grid, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)cyc[i]; avg /= grid;
    free(cyc);
    double ops_per_thread = (MIX == FPMUL_K256) ? (double)ITERS * 73 : (double)ITERS * CH * MIXOPS[MIX];
    double ops_per_sm = ops_per_thread * block * 2;      // 2 CTAs resident per SM
    double per_clk_sm = ops_per_sm / avg;
    double gops = ops_per_thread * block * (double)grid / (ms * 1e-3) / 1e9;
    printf("  \"%s\": {\"thread_ops_per_clk_per_sm\": %.2f, \"chip_gops\": %.1f, \"ms\": %.4f, \"implied_mhz\": %.0f}%s\n",
           MIXNAME[MIX], per_clk_sm, gops, ms, gops * 1e9 / (per_clk_sm * nsm) / 1e6, last ? "" : ",");
    (void)clk_mhz;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    u32* sink; u64* cyc;
    CK(cudaMalloc(&sink, sizeof(u32) * nsm * 2 * 512));
    CK(cudaMalloc(&cyc, sizeof(u64) * nsm * 2));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\n \"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f,\n \"note\": \"ops = per-thread instructions (or instruction pairs as named); imad_wide* = one 32x32->64 multiply-accumulate\",\n \"mixes\": {\n",
           p.name, nsm, clk_khz / 1e3);
    run<IMAD_LO>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<IMAD_HI>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<IMAD_WIDE>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<IMAD_WIDE_X>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<IADD3_>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<WIDE_PLUS_IADD3>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<LOHI_PAIR>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<FFMA_>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<DFMA_>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<WIDE_X_PLUS_2IADD3>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<WIDE_X_SHARED_B>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<MULWIDE8>(nsm, sink, cyc, clk_khz / 1e3, false);
    run<FPMUL_K256>(nsm, sink, cyc, clk_khz / 1e3, true);
    printf(" }\n}\n");
    return 0;
}
