// abi.cu — host runtime behind include/ecb200.h: context, device scratch, fixed-base table
// construction, host<->device staging and the extern "C" entry points.  No arithmetic happens on
// the host; every entry point fails loudly when CUDA is unavailable.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <thread>

#include "../../include/ecb200.h"
#include "launch.h"

namespace ecb {
thread_local uint64_t* tl_launch_counter = nullptr;
}
using namespace ecb;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// scratch that lives for one function call: freed on every return path (the CU() macro returns early on errors)
struct ScopedDev {
    void* p = nullptr;
    ~ScopedDev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <class T> T* as() const { return (T*)p; }
    void* release() { void* q = p; p = nullptr; return q; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Elements per pipelined piece of the host entry points.  One-row-per-thread kernels run 148 SMs x {6, 4, 3} resident CTAs x
// 128 threads per wave (113 664 / 75 776 / 56 832 rows); pieces that are whole multiples of all three (148 x 1536 = 227 328)
// do not end in a mostly empty last wave - a 2^20-row piece is 9.23 waves of the verify kernel.  ECB200_CHUNK overrides (tests).
#ifndef ECB200_CHUNK
#define ECB200_CHUNK (5 * 227328)
#endif
#ifndef ECB200_FIRST_CHUNK
#define ECB200_FIRST_CHUNK 227328
#endif
constexpr size_t CHUNK = (size_t)ECB200_CHUNK;
constexpr int NCURVE = 6;                   // K256, P256, P384, SM2, P192, P224 (ecb200_curve)
constexpr size_t MAX_ROWS = 0x7FFFFFFF;      // rows per device-pointer call (kernels index rows with int); host entry points chunk
constexpr int NSLOT = 2;                    // double buffering: H2D / kernel / D2H of adjacent chunks overlap

}  // namespace

struct ecb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;           // compute stream of the host entry points
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[NSLOT] = {}, ev_done[NSLOT] = {}, ev_out[NSLOT] = {};
    const CurveLaunch* cl[NCURVE] = {};
    uint32_t* gtab[NCURVE] = {};                  // affine multiples 1..ngtab of G (internal limbs)
    uint32_t* gentab[NCURVE] = {};                // fixed-base window tables, 4-bit windows (k_mul_gen_smem)
    uint32_t* gentab2[NCURVE] = {};               // fixed-base window tables of the split path (k_gen_half)
    bool use_gen2 = true;                    // ECB200_GEN2=0: one thread per scalar, complete additions (A/B comparisons)
    uint32_t* gbig[NCURVE] = {};                  // big fixed-base tables of the public-input fast path (built on first use)
    int gw = 0;                              // window width of gbig forced by ECB200_GW (2 .. 24 bits); 0 = the per-curve default, gw_of()
    bool verify_v1 = false;                  // ECB200_VERIFY_V1=1: complete-formula verify kernel (A/B comparisons)
    bool use_wintab = true;                  // ECB200_WINTAB=0: per-thread Jacobian window tables on the primeorder curves (A/B comparisons)
    DevBuf prep, aff;                        // verify_prep scratch; affine limbs of normalised projective inputs
    DevBuf wtab;                             // per-row affine window tables of the primeorder public-input path (23 L words per row)
    DevBuf kxy, kst;                         // decoded keys (x||y) and their status bytes; also xy / identity flags of the Schnorr epilogue
    DevBuf proj;                             // projective scratch (n x 3L limbs) — shared by all entry points
    DevBuf partial, one_point;
    DevBuf d_in[NSLOT][4], d_out[NSLOT][3];  // staging for host-pointer entry points
    PinBuf h_in[NSLOT][4], h_out[NSLOT][3];
    DevBuf proj2, inv2;                      // second product and its invalid flags (per-row two-term lincomb)
    // Per-key window tables of the verify path (kernels.cuh): state of ONE call - the rows of its chunks are grouped by public
    // key incrementally, tables are built once per distinct key and dropped when the call returns (nothing is cached between calls)
    struct KeyTab {
        DevBuf htab, gkeys, tab, kvalid, jac, rep, rep_slot, newgid, gid, counter;
        size_t cap = 0, alloc_words = 0, total_rows = 0;   // alloc_words: u32 words of `tab` (the width, hence the words per key, can change from call to call)
        uint32_t hmask = 0;
        int built = 0;
        int wide = 1;                        // table width of this call: 1 = the curve's default, 0 = the narrow one (few rows per key)
        bool active = false, disabled = true, first = true;
    } kt;
    int kt_force_wide = -1;                  // ECB200_KT_WIDE = 0 / 1 forces the width (tests, A/B); -1 = by rows per key
    bool use_keytab = true;                  // ECB200_KEYTAB=0: always the per-row path (A/B comparisons, tests of both paths)
    uint64_t kt_rows = 0, kt_groups = 0;     // statistics: rows verified on per-key tables / tables built (ecb200_keytab_stats)
    std::string err;
    uint64_t launches = 0;                   // kernels launched by this context (launch.h count_launch)
    // ecb200_init_multi: the parent owns one child context per device and no device state of its own; host entry points
    // shard [0, n) by contiguous index range over the children, one host thread per device
    std::vector<ecb200_ctx*> kids;
    bool held_secrets = false;               // a signing / FLAG_CT call staged secret scalars in this context
    size_t in_used[NSLOT][4] = {};           // bytes of each staging buffer written by the current call (for wiping)
    // optional per-launch timing of the dominant kernel (verify_main) with CUDA events on its own stream
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
};

namespace {

int fail(ecb200_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}
#define CU(ctx, call)                                                         \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail(ctx, ECB200_ERR_CUDA, #call, e_);  \
    } while (0)

// points the launch counter of this host thread at the calling context for the duration of one entry point
struct Enter {
    explicit Enter(ecb200_ctx* c) { tl_launch_counter = c ? &c->launches : nullptr; }
    ~Enter() { tl_launch_counter = nullptr; }
};

const CurveLaunch* curve_of(ecb200_ctx* c, int curve) {
    if (!c || curve < 0 || curve >= NCURVE) return nullptr;
    return c->cl[curve];
}
bool resolve_compress(const CurveLaunch* cl, uint32_t flags) {
    if (flags & ECB200_FLAG_COMPRESSED) return true;
    if (flags & ECB200_FLAG_UNCOMPRESSED) return false;
    return cl->compress_default;
}
size_t slot_bytes(const CurveLaunch* cl, uint32_t flags) { return 1 + (resolve_compress(cl, flags) ? cl->FB : 2 * cl->FB); }

// Build affine tables of multiples of G on the device with the ordinary kernels:
// entry e = scalar[e] * G, scalars given as big-endian bytes.
int build_table(ecb200_ctx* c, const CurveLaunch* cl, const std::vector<uint8_t>& scalars, int entries, uint32_t** out_tab,
                const std::vector<uint8_t>& gxy) {
    const int FB = cl->FB, L = cl->L;
    std::vector<uint8_t> pts((size_t)entries * 2 * FB);
    for (int e = 0; e < entries; e++) memcpy(&pts[(size_t)e * 2 * FB], gxy.data(), 2 * FB);
    ScopedDev pts_b, k_b, proj_b, tab_b;
    CU(c, pts_b.alloc(pts.size()));
    CU(c, k_b.alloc(scalars.size()));
    CU(c, proj_b.alloc((size_t)entries * 3 * L * 4));
    CU(c, tab_b.alloc((size_t)entries * 2 * L * 4));
    uint8_t *d_pts = pts_b.as<uint8_t>(), *d_k = k_b.as<uint8_t>();
    uint32_t* d_proj = proj_b.as<uint32_t>();
    CU(c, cudaMemcpyAsync(d_pts, pts.data(), pts.size(), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d_k, scalars.data(), scalars.size(), cudaMemcpyHostToDevice, c->stream));
    cl->mul_var(c->stream, false, entries, 0, d_pts, nullptr, d_k, d_proj, nullptr);
    cl->normalize(c->stream, entries, d_proj, 2 /*NORM_AFF_LIMBS*/, 0, nullptr, nullptr, tab_b.as<uint32_t>());
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));
    *out_tab = (uint32_t*)tab_b.release();   // the table outlives the call; everything else is freed here
    return 0;
}

// generator coordinates as canonical bytes
const char* GX[NCURVE] = {
    "79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798",
    "6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296",
    "AA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7",
    "32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7",
    "188DA80EB03090F67CBF20EB43A18800F4FF0AFD82FF1012",
    "B70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21"};
const char* GY[NCURVE] = {
    "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8",
    "4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5",
    "3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F",
    "BC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0",
    "07192B95FFC8DA78631011ED6B24CDD573F977A11E794811",
    "BD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34"};

void hex_to(uint8_t* dst, const char* hex, int nbytes) {
    for (int i = 0; i < nbytes; i++) {
        unsigned v;
        sscanf(hex + 2 * i, "%2x", &v);
        dst[i] = (uint8_t)v;
    }
}

int build_tables(ecb200_ctx* c) {
    for (int id = 0; id < NCURVE; id++) {
        const CurveLaunch* cl = c->cl[id];
        const int FB = cl->FB;
        std::vector<uint8_t> gxy(2 * FB);
        hex_to(gxy.data(), GX[id], FB);
        hex_to(gxy.data() + FB, GY[id], FB);
        {   // small table: j*G, j = 1..ngtab
            std::vector<uint8_t> sc((size_t)cl->ngtab * FB, 0);
            for (int j = 0; j < cl->ngtab; j++) sc[(size_t)j * FB + FB - 1] = (uint8_t)(j + 1);
            int r = build_table(c, cl, sc, cl->ngtab, &c->gtab[id], gxy);
            if (r) return r;
        }
        if (cl->gen_windows) {   // (j+1) * 16^i * G, i < 8L + 1, j < 8
            const int ne = cl->gen_windows * cl->gen_entries;
            std::vector<uint8_t> sc((size_t)ne * FB, 0);
            for (int e = 0; e < ne; e++) {
                int i = e / cl->gen_entries, j = e % cl->gen_entries;
                uint8_t* s = &sc[(size_t)e * FB];
                int bit = 4 * i;
                unsigned v = (unsigned)(j + 1);
                if (bit < 8 * FB) {
                    int byte = bit / 8, sh = bit % 8;
                    unsigned w = v << sh;
                    s[FB - 1 - byte] |= (uint8_t)w;
                    if (byte + 1 < FB) s[FB - 2 - byte] |= (uint8_t)(w >> 8);
                } else {
                    // (j+1) * (2^(8FB) mod n): a small multiple of a value far below n (129 bits for k256, < 2^225
                    // for P-256 / SM2, < 2^191 for P-384), so no reduction is needed
                    unsigned carry = 0;
                    for (int b = FB - 1; b >= 0; b--) {
                        unsigned t = (unsigned)cl->r_mod_n[b] * v + carry;
                        s[b] = (uint8_t)t;
                        carry = t >> 8;
                    }
                }
            }
            int r = build_table(c, cl, sc, ne, &c->gentab[id], gxy);
            if (r) return r;
        }
        if (cl->gen2_windows) {   // split fixed-base path: v * 2^(W w) * G, w < gen2_windows, 1 <= v <= gen2_entries
            const int W = cl->gen2_w, NW = cl->gen2_windows, E = cl->gen2_entries;
            const int top_bits = 8 * FB - W * (NW - 1);            // bits of the scalar that fall into the (unsigned) top window
            std::vector<uint8_t> sc((size_t)NW * E * FB, 0);
            for (int w = 0; w < NW; w++)
                for (int v = 1; v <= E; v++) {
                    uint8_t* s = &sc[((size_t)w * E + v - 1) * FB];
                    if (w == NW - 1 && v == (1 << top_bits)) { memcpy(s, cl->r_mod_n, FB); continue; }   // 2^(8 FB) mod n
                    if (w == NW - 1 && v > (1 << top_bits)) { s[FB - 1] = 1; continue; }                 // never selected
                    for (int b = 0; b < W; b++)
                        if ((v >> b) & 1) { const int pos = W * w + b; s[FB - 1 - pos / 8] |= (uint8_t)(1u << (pos % 8)); }
                }
            int r = build_table(c, cl, sc, NW * E, &c->gentab2[id], gxy);
            if (r) return r;
        }
    }
    return 0;
}

cudaStream_t pick(ecb200_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->stream; }

// Big fixed-base table for u1*G on the public-input path (jac.cuh add_fixed_base): signed gw-bit windows below the
// top one, entry (w << (gw-1)) + v - 1 = v * 2^(gw*w) * G for 1 <= v <= 2^(gw-1); the top window is unsigned with
// v up to 2^gw (v = 2^gw is 2^(8 FB) * G).  Affine internal limbs (34 MiB for a 256-bit curve at gw = 16; it lives in
// HBM/L2 and is gathered 64-96 B at a time).
// Built on first use with the engine's own kernels: scalar v << (gw*w), point G, Jacobian fast path, normalise.
// Window width of the table.  Measured on the B200 at 2^22 rows / 2^16 keys (profiles/r02_ab_fixed_base_width.txt), M verifies/s at
// gw = 16 / 18 / 20 / 22: secp256k1 118.5 / 120.1 / 123.4 / 124.6, P-256 107.0 / 108.3 / 110.9 / 112.2, per-row secp256k1 53.8 /
// 54.1 / 54.7 / 55.1, P-384 (2^20 rows) 21.2 / 21.4 / 21.6 / 21.8.  The 256-bit and smaller curves take 20 bits (13 additions
// instead of 16 for 410 MB of HBM per curve, built in ~0.15 s on first use; 22 bits would buy one more addition for 1.5 GB);
// P-384 keeps 16 (78 MiB; 20 bits would cost 956 MB for +2 %).
int gw_of(const ecb200_ctx* c, const CurveLaunch* cl) { return c->gw ? c->gw : (cl->FB > 32 ? 16 : 20); }
int ensure_gbig(ecb200_ctx* c, const CurveLaunch* cl) {
    if (c->gbig[cl->id]) return 0;
    const int FB = cl->FB, L = cl->L, gw = gw_of(c, cl);
    const int nwin = (8 * FB + gw - 1) / gw, tb = 8 * FB - gw * (nwin - 1);      // windows; bits of the (unsigned) top window
    const size_t per = (size_t)1 << (gw - 1), top_base = (size_t)(nwin - 1) * per, ne = top_base + ((size_t)1 << tb);
    std::vector<uint8_t> gxy(2 * FB);
    hex_to(gxy.data(), GX[cl->id], FB);
    hex_to(gxy.data() + FB, GY[cl->id], FB);
    ScopedDev big_b, pts_b, k_b, proj_b;      // the table is published in c->gbig only when it is complete
    CU(c, big_b.alloc(ne * 2 * L * 4));
    const size_t chunk = std::min<size_t>(ne, (size_t)1 << 18);
    std::vector<uint8_t> pts(chunk * 2 * FB), sc(chunk * FB);
    for (size_t e = 0; e < chunk; e++) memcpy(&pts[e * 2 * FB], gxy.data(), 2 * FB);
    CU(c, pts_b.alloc(pts.size()));
    CU(c, k_b.alloc(sc.size()));
    CU(c, proj_b.alloc(chunk * 3 * L * 4));
    uint8_t *d_pts = pts_b.as<uint8_t>(), *d_k = k_b.as<uint8_t>();
    uint32_t *d_proj = proj_b.as<uint32_t>(), *d_big = big_b.as<uint32_t>();
    CU(c, cudaMemcpyAsync(d_pts, pts.data(), pts.size(), cudaMemcpyHostToDevice, c->stream));
    for (size_t off = 0; off < ne; off += chunk) {
        size_t cnt = std::min(chunk, ne - off);
        std::fill(sc.begin(), sc.end(), 0);
        for (size_t e = 0; e < cnt; e++) {
            size_t idx = off + e;
            uint8_t* s = &sc[e * FB];
            if (idx == ne - 1) { memcpy(s, cl->r_mod_n, FB); continue; }   // v = 2^tb in the top window: 2^(8 FB) mod n
            size_t w = idx < top_base ? idx >> (gw - 1) : (size_t)(nwin - 1), v = idx - w * per + 1;
            int bit = (int)w * gw;
            for (int b = 0; b < gw; b++)
                if ((v >> b) & 1) { int pos = bit + b; s[FB - 1 - pos / 8] |= (uint8_t)(1u << (pos % 8)); }
        }
        CU(c, cudaMemcpyAsync(d_k, sc.data(), cnt * FB, cudaMemcpyHostToDevice, c->stream));
        cl->mul_var_fast(c->stream, (int)cnt, d_pts, nullptr, nullptr, d_k, d_proj, nullptr, nullptr);
        cl->normalize(c->stream, (int)cnt, d_proj, 2 /*NORM_AFF_LIMBS*/, 0, nullptr, nullptr, d_big + off * 2 * L);
        CU(c, cudaGetLastError());
        CU(c, cudaStreamSynchronize(c->stream));
    }
    c->gbig[cl->id] = (uint32_t*)big_b.release();
    return 0;
}

// -------------------------------------------------------------------------------------------
// device-pointer cores (enqueue only)

// ecb200_kernel_timing: bracket the dominant kernel of an operation with CUDA events on the stream it is launched on
struct TimedLaunch {
    ecb200_ctx* c;
    cudaStream_t s;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TimedLaunch(ecb200_ctx* c_, cudaStream_t s_) : c(c_), s(s_) {
        if (!c->timing) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, s);
    }
    ~TimedLaunch() {
        if (!e0) return;
        cudaEventRecord(e1, s);
        c->timed.emplace_back(e0, e1);
    }
};

// k*G for n scalars, normalised into SEC1 slots (mode 0, d_out) or affine limbs (mode 2, d_limbs).  Default: the split
// form - two threads per scalar sum one half of the windows each (Jacobian mixed additions), the halves are added with the
// complete formula inside the normalisation kernel.  ECB200_GEN2=0: one thread per scalar, complete mixed additions.
int gen_points(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_k, bool ct, int mode, int compress, uint8_t* d_out,
               uint32_t* d_limbs, cudaStream_t s) {
    if (c->use_gen2 && c->gentab2[cl->id]) {
        const size_t half = n * 3 * (size_t)cl->L;
        CU(c, c->proj.reserve(2 * half * 4));
        uint32_t* part = (uint32_t*)c->proj.p;
        {
            TimedLaunch t(c, s);
            cl->mul_gen2(s, ct, (int)n, d_k, c->gentab2[cl->id], part);
        }
        cl->sum_normalize(s, (int)n, part, part + half, mode, compress, d_out, nullptr, d_limbs);
    } else {
        CU(c, c->proj.reserve(n * 3 * cl->L * 4));
        {
            TimedLaunch t(c, s);
            cl->mul_gen(s, ct, (int)n, d_k, c->gentab[cl->id], (uint32_t*)c->proj.p);
        }
        cl->normalize(s, (int)n, (const uint32_t*)c->proj.p, mode, compress, d_out, nullptr, d_limbs);
    }
    return 0;
}
int mul_gen_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_k, uint8_t* d_out, uint32_t flags, cudaStream_t s) {
    int r = gen_points(c, cl, n, d_k, (flags & ECB200_FLAG_CT) != 0, 0, resolve_compress(cl, flags) ? 1 : 0, d_out, nullptr, s);
    if (r) return r;
    CU(c, cudaGetLastError());
    return 0;
}
// Primeorder curves: make the window tables {1..8}P of all n rows affine with one inversion per 16 rows x 7 entries
// (k_wintab), so that the window loop of the public-input kernels runs on mixed additions.  *out stays NULL on secp256k1
// (shared-Z table inside the kernel, no inversion needed) and with ECB200_WINTAB=0 (A/B: per-thread Jacobian tables).
int window_tables(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_pts, const uint32_t* d_aff, cudaStream_t s, uint32_t** out) {
    *out = nullptr;
    if (cl->id == ECB200_K256 || !c->use_wintab || n == 0) return 0;
    CU(c, c->wtab.reserve(n * 23 * (size_t)cl->L * 4));
    cl->wintab(s, (int)n, d_pts, d_aff, (uint32_t*)c->wtab.p);
    *out = (uint32_t*)c->wtab.p;
    return 0;
}
int mul_var_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_pts, const uint8_t* d_inf, const uint8_t* d_k,
                 uint8_t* d_out, uint8_t* d_invalid, uint32_t flags, cudaStream_t s) {
    CU(c, c->proj.reserve(n * 3 * cl->L * 4));
    uint32_t* proj = (uint32_t*)c->proj.p;
    if (flags & ECB200_FLAG_CT) {
        // secret scalars: complete formulas, fixed windows, full table scans
        TimedLaunch t(c, s);
        cl->mul_var(s, true, (int)n, flags, d_pts, (flags & ECB200_FLAG_PROJ) ? nullptr : d_inf, d_k, proj, d_invalid);
    } else if (flags & ECB200_FLAG_PROJ) {
        // public scalars, projective inputs: normalise once (Montgomery trick), then the Jacobian fast path
        CU(c, c->aff.reserve(n * 2 * cl->L * 4));
        cl->load_proj(s, (int)n, d_pts, proj, nullptr);
        cl->normalize(s, (int)n, proj, 2 /*NORM_AFF_LIMBS*/, 0, nullptr, nullptr, (uint32_t*)c->aff.p);
        uint32_t* wt = nullptr;
        int r = window_tables(c, cl, n, nullptr, (const uint32_t*)c->aff.p, s, &wt);
        if (r) return r;
        TimedLaunch t(c, s);
        cl->mul_var_fast(s, (int)n, nullptr, (const uint32_t*)c->aff.p, nullptr, d_k, proj, d_invalid, wt);
    } else {
        uint32_t* wt = nullptr;
        int r = window_tables(c, cl, n, d_pts, nullptr, s, &wt);
        if (r) return r;
        TimedLaunch t(c, s);
        cl->mul_var_fast(s, (int)n, d_pts, nullptr, d_inf, d_k, proj, d_invalid, wt);
    }
    cl->normalize(s, (int)n, proj, 0, resolve_compress(cl, flags) ? 1 : 0, d_out, nullptr, nullptr);
    CU(c, cudaGetLastError());
    return 0;
}
int batch_normalize_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_xyz, uint8_t* d_xy, uint8_t* d_inf, cudaStream_t s) {
    CU(c, c->proj.reserve(n * 3 * cl->L * 4));
    cl->load_proj(s, (int)n, d_xyz, (uint32_t*)c->proj.p, nullptr);
    cl->normalize(s, (int)n, (const uint32_t*)c->proj.p, 1, 0, d_xy, d_inf, nullptr);
    CU(c, cudaGetLastError());
    return 0;
}
enum { VM_ECDSA = 0, VM_SM2DSA = 1, VM_SCHNORR = 2, VM_RECOVER = 3, DEC_SEC1 = 0, DEC_COMPACT = 1, FIN_SCHNORR = 0, FIN_RECOVER = 1 };   // = kernels.cuh

// ---- per-key window tables (kernels.cuh).  Policy: a call is eligible when it is large enough to amortise anything
// (KT_MIN_ROWS); tables are used while the call shows at least KT_MIN_REUSE rows per distinct key (a narrow table costs ~3
// verifications, saves ~0.5 per row) and the distinct keys fit the table memory; the first chunk of a multi-chunk host call
// must already show two rows per key, so a call with all-distinct keys pays only the grouping kernels (~0.4 % of a step).
// Table width: the curve's wide default (6 bits on secp256k1, 5 on P-256) when the call has at least KT_WIDE_REUSE rows per
// distinct key, else 4 bits - a table is paid per key, its additions per row (kernels.cuh).  Decided once per call, on the
// first chunk: total rows of the call / distinct keys seen so far.
// Measured (profiles/r02_ab_table_width_by_reuse.txt, M verifies/s narrow / wide at 8, 16, 32, 64 rows per key): secp256k1 67.7 / 57.7,
// 82.7 / 82.5, 93.8 / 106.2, 100.0 / 123.4; P-256 52.5 / 47.6, 71.6 / 70.4, 88.1 / 93.2, 99.3 / 111.0 - the widths cross at ~16 rows per
// key, and the narrow tables already beat the per-row path (54.7 / 34.1 M/s) well below 8: their cost per key equals ~3 rows' worth.
constexpr size_t KT_MIN_ROWS = 4096, KT_MIN_REUSE = 4, KT_WIDE_REUSE = 20;
constexpr size_t KT_MEM_CAP = (size_t)12 << 30;      // bytes of tables per context (180 GB of HBM per GPU)

int kt_begin(ecb200_ctx* c, const CurveLaunch* cl, size_t total_rows, size_t max_chunk, cudaStream_t s) {
    ecb200_ctx::KeyTab& k = c->kt;
    k.active = true;
    k.first = true;
    k.built = 0;
    k.total_rows = total_rows;
    k.disabled = !c->use_keytab || c->verify_v1 || total_rows < KT_MIN_ROWS || total_rows >= ((size_t)1 << 30);
    if (k.disabled) return 0;
    const size_t key_bytes = (size_t)std::max(cl->kt_key_words[0], cl->kt_key_words[1]) * 4;
    k.cap = std::min(total_rows / KT_MIN_REUSE, KT_MEM_CAP / key_bytes);
    if (k.cap < 1) { k.disabled = true; return 0; }
    size_t hs = 1024;
    while (hs < 2 * (k.cap + max_chunk)) hs <<= 1;
    k.hmask = (uint32_t)(hs - 1);
    CU(c, k.htab.reserve(hs * 4));
    CU(c, k.counter.reserve(16));
    CU(c, k.gkeys.reserve(k.cap * (size_t)cl->kt_kbw * 4));
    CU(c, k.kvalid.reserve(k.cap));
    CU(c, cudaMemsetAsync(k.htab.p, 0xFF, hs * 4, s));
    CU(c, cudaMemsetAsync(k.counter.p, 0, 16, s));
    return 0;
}
void kt_end(ecb200_ctx* c) { c->kt.active = false; c->kt.disabled = true; }

// groups the n rows of a chunk by key, decides, and builds the tables of the keys first seen in this chunk; *use = verify on tables.
// (Measured and not kept: the chunk's verify_prep on a side stream next to the table construction - the step stayed at 34.0 ms,
// k_kt_fill keeps the multiplier pipe 60 % busy on its own.)
int kt_chunk(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_q, cudaStream_t s, bool* use) {
    ecb200_ctx::KeyTab& k = c->kt;
    *use = false;
    CU(c, k.gid.reserve(n * 4));
    CU(c, k.rep.reserve(n * 4));
    CU(c, k.rep_slot.reserve(n * 4));
    CU(c, k.newgid.reserve(n * 4));
    cl->kt_group(s, (int)n, (const uint32_t*)d_q, (int*)k.htab.p, k.hmask, (uint32_t*)k.gkeys.p, (int*)k.gid.p, (int*)k.rep.p, (int*)k.rep_slot.p,
                 (int*)k.newgid.p, (int*)k.counter.p, (int)k.cap);
    int D = 0;
    CU(c, cudaMemcpyAsync(&D, k.counter.p, 4, cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));      // the one host round trip of this path: the number of distinct keys decides what is launched next
    const bool first = k.first;
    k.first = false;
    if ((size_t)D > k.cap || (size_t)D * KT_MIN_REUSE > k.total_rows || (first && n < k.total_rows && (size_t)D * 2 > n)) {
        k.disabled = true;                // little reuse (or more keys than tables): this chunk and the rest of the call take the per-row path
        return 0;
    }
    if (first) k.wide = c->kt_force_wide >= 0 ? c->kt_force_wide : ((size_t)D * KT_WIDE_REUSE <= k.total_rows ? 1 : 0);
    if (D > k.built) {
        const size_t kw = (size_t)cl->kt_key_words[k.wide];
        if ((size_t)D * kw > k.alloc_words) {  // grow the table store (position independent: old tables are copied over)
            size_t want = std::min(k.cap, std::max<size_t>((size_t)2 * D, 65536));
            DevBuf bigger;
            CU(c, bigger.reserve(want * kw * 4));
            if (k.built > 0) CU(c, cudaMemcpyAsync(bigger.p, k.tab.p, (size_t)k.built * kw * 4, cudaMemcpyDeviceToDevice, s));
            CU(c, cudaStreamSynchronize(s));
            k.tab.release();
            k.tab = bigger;
            k.alloc_words = want * kw;
        }
        const int cnt = D - k.built;
        CU(c, k.jac.reserve((size_t)cnt * cl->kt_windows[k.wide] * 3 * cl->L * 4));
        cl->kt_build(s, k.wide, k.built, cnt, (const uint32_t*)k.gkeys.p, (uint32_t*)k.jac.p, (uint8_t*)k.kvalid.p, (uint32_t*)k.tab.p);

        c->kt_groups += (uint64_t)cnt;
        k.built = D;
    }
    *use = true;
    return 0;
}

int verify_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_q, const uint8_t* d_z, const uint8_t* d_rs, uint8_t* d_ok, cudaStream_t s,
                int mode = VM_ECDSA) {
    if (c->verify_v1 && mode == VM_ECDSA) {
        cl->verify(s, (int)n, d_q, d_z, d_rs, c->gtab[cl->id], d_ok);
    } else {
        int r = ensure_gbig(c, cl);
        if (r) return r;
        CU(c, c->prep.reserve(n * (size_t)cl->prep_words * 4));
        if (c->kt.active && !c->kt.disabled && (mode == VM_ECDSA || mode == VM_SM2DSA) && ((uintptr_t)d_q & 3) == 0) {
            bool use = false;
            r = kt_chunk(c, cl, n, d_q, s, &use);
            if (r) return r;
            if (use) {     // rows of repeated keys: no doublings, gathered additions from the per-key tables
                cl->verify_prep(s, (int)n, mode, d_z, d_rs, (uint32_t*)c->prep.p);
                {
                    TimedLaunch t(c, s);
                    cl->verify_keytab(s, c->kt.wide, (int)n, mode, d_rs, d_z, (const uint32_t*)c->prep.p, (const int*)c->kt.gid.p, (const uint8_t*)c->kt.kvalid.p,
                                      (const uint32_t*)c->kt.tab.p, c->gbig[cl->id], gw_of(c, cl), d_ok);
                }
                c->kt_rows += n;
                CU(c, cudaGetLastError());
                return 0;
            }
        }
        cl->verify_prep(s, (int)n, mode, d_z, d_rs, (uint32_t*)c->prep.p);
        uint32_t* wt = nullptr;
        r = window_tables(c, cl, n, d_q, nullptr, s, &wt);
        if (r) return r;
        TimedLaunch t(c, s);
        cl->verify_main(s, (int)n, mode, d_q, d_rs, d_z, nullptr, (const uint32_t*)c->prep.p, c->gbig[cl->id], gw_of(c, cl), d_ok, nullptr, wt);
    }
    CU(c, cudaGetLastError());
    return 0;
}
// SEC1 / compact decoding into x||y + status (f1)
int decode_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_enc, size_t stride, int mode, uint8_t* d_xy, uint8_t* d_status, cudaStream_t s) {
    cl->decode(s, (int)n, mode, d_enc, (int)stride, d_xy, d_status);
    CU(c, cudaGetLastError());
    return 0;
}
// ECDSA verification with SEC1-encoded keys: decode on the device, then the ordinary pipeline (an undecodable key
// becomes x||y = 0, which is off every supported curve, so the row is rejected by the key check of the main kernel)
int verify_sec1_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_keys, size_t stride, const uint8_t* d_z, const uint8_t* d_rs,
                     uint8_t* d_ok, cudaStream_t s) {
    CU(c, c->kxy.reserve(n * 2 * (size_t)cl->FB));
    CU(c, c->kst.reserve(n));
    cl->decode(s, (int)n, DEC_SEC1, d_keys, (int)stride, (uint8_t*)c->kxy.p, (uint8_t*)c->kst.p);
    return verify_core(c, cl, n, (const uint8_t*)c->kxy.p, d_z, d_rs, d_ok, s);
}
// BIP340: prep (mod n) -> main (lift_x, s*G - e*P) -> normalise -> byte-level epilogue (R finite, y even, x == r)
int schnorr_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_pk, const uint8_t* d_e, const uint8_t* d_sig, uint8_t* d_ok, cudaStream_t s) {
    int r = ensure_gbig(c, cl);
    if (r) return r;
    CU(c, c->prep.reserve(n * (size_t)cl->prep_words * 4));
    CU(c, c->proj.reserve(n * 3 * (size_t)cl->L * 4));
    CU(c, c->kxy.reserve(n * 2 * (size_t)cl->FB));
    CU(c, c->kst.reserve(n));
    cl->verify_prep(s, (int)n, VM_SCHNORR, d_e, d_sig, (uint32_t*)c->prep.p);
    cl->verify_main(s, (int)n, VM_SCHNORR, d_pk, d_sig, d_e, nullptr, (const uint32_t*)c->prep.p, c->gbig[cl->id], gw_of(c, cl), d_ok, (uint32_t*)c->proj.p, nullptr);
    cl->normalize(s, (int)n, (const uint32_t*)c->proj.p, 1 /*NORM_XY_BYTES*/, 0, (uint8_t*)c->kxy.p, (uint8_t*)c->kst.p, nullptr);
    cl->finish(s, (int)n, FIN_SCHNORR, (const uint8_t*)c->kxy.p, 0, (const uint8_t*)c->kst.p, d_sig, d_ok);
    CU(c, cudaGetLastError());
    return 0;
}
// public-key recovery: prep (r^-1) -> main (decompress R, u1*G + u2*R) -> normalise to SEC1 -> ok &= key != identity
int recover_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_z, const uint8_t* d_rs, const uint8_t* d_recid, uint8_t* d_keys,
                 uint8_t* d_ok, uint32_t flags, cudaStream_t s) {
    int r = ensure_gbig(c, cl);
    if (r) return r;
    CU(c, c->prep.reserve(n * (size_t)cl->prep_words * 4));
    CU(c, c->proj.reserve(n * 3 * (size_t)cl->L * 4));
    cl->verify_prep(s, (int)n, VM_RECOVER, d_z, d_rs, (uint32_t*)c->prep.p);
    cl->verify_main(s, (int)n, VM_RECOVER, nullptr, d_rs, d_z, d_recid, (const uint32_t*)c->prep.p, c->gbig[cl->id], gw_of(c, cl), d_ok, (uint32_t*)c->proj.p, nullptr);
    cl->normalize(s, (int)n, (const uint32_t*)c->proj.p, 0, resolve_compress(cl, flags) ? 1 : 0, d_keys, nullptr, nullptr);
    cl->finish(s, (int)n, FIN_RECOVER, d_keys, (int)slot_bytes(cl, flags), nullptr, nullptr, d_ok);
    CU(c, cudaGetLastError());
    return 0;
}
// signing: constant-time fixed-base k*G -> normalise -> (r, s, recid)
int sign_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_d, const uint8_t* d_k, const uint8_t* d_z, uint8_t* d_rs, uint8_t* d_recid,
              uint8_t* d_ok, cudaStream_t s) {
    CU(c, c->aff.reserve(n * 2 * (size_t)cl->L * 4));
    int r = gen_points(c, cl, n, d_k, true, 2 /*NORM_AFF_LIMBS*/, 0, nullptr, (uint32_t*)c->aff.p, s);
    if (r) return r;
    cl->sign_finish(s, (int)n, d_d, d_k, d_z, (const uint32_t*)c->aff.p, d_rs, d_recid, d_ok);
    CU(c, cudaGetLastError());
    return 0;
}

// Staging copy between pageable caller memory and the context's pinned buffers.  One host thread moves ~10 GB/s, a third
// of what the 2^22-row verify pipeline consumes (671 MB per 34 ms), so large copies are cut over a few threads
// (measured with bench.py's pageable-buffer leg: 72.6 M verifies/s single-threaded against 117.9 M/s from page-locked buffers).
void staged_copy(void* dst, const void* src, size_t bytes) {
    constexpr size_t MIN_PER_THREAD = (size_t)4 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t t = std::min<size_t>({(size_t)8, hw ? (size_t)std::max(1u, hw / 2) : (size_t)1, bytes / MIN_PER_THREAD});
    if (t <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    th.reserve(t - 1);
    const size_t per = ((bytes + t - 1) / t + 63) & ~(size_t)63;
    for (size_t i = 1; i < t; i++) {
        const size_t lo = std::min(bytes, i * per), hi = std::min(bytes, lo + per);
        if (hi > lo) th.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, per));
    for (auto& x : th) x.join();
}

// -------------------------------------------------------------------------------------------
// Host-pointer pipeline: the batch is cut into CHUNK-element pieces; piece i is copied into pinned
// memory and sent H2D on the copy-in stream while piece i-1 computes and piece i-2 drains D2H.
// in[k]/out[k]: host arrays with per-element byte sizes in_sz[k]/out_sz[k] (0 = unused, NULL allowed).
template <class Fn>
int run_pipeline_pieces(ecb200_ctx* c, size_t n, int n_in, const uint8_t* const* in, const size_t* in_sz, int n_out, uint8_t* const* out,
                        const size_t* out_sz, Fn&& enqueue) {
    if (n == 0) return 0;
    const size_t FIRST = (size_t)ECB200_FIRST_CHUNK;   // the first piece is short, so the (unoverlapped) first H2D is short
    auto piece = [&](size_t ch, size_t& off, size_t& cnt) {
        off = ch == 0 ? 0 : FIRST + (ch - 1) * CHUNK;
        cnt = std::min(ch == 0 ? FIRST : CHUNK, n - off);
    };
    size_t nchunks = n <= FIRST ? 1 : 1 + (n - FIRST + CHUNK - 1) / CHUNK;
    const size_t cap = std::min(n, CHUNK);   // staging capacity in elements
    // caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister) are copied from / to directly;
    // pageable ones go through the context's pinned staging buffers
    bool in_pinned[4] = {}, out_pinned[3] = {};
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    for (int k = 0; k < n_in; k++) in_pinned[k] = in[k] && in_sz[k] && is_pinned(in[k]);
    for (int k = 0; k < n_out; k++) out_pinned[k] = out[k] && out_sz[k] && is_pinned(out[k]);
    for (size_t ch = 0; ch < nchunks + 1; ch++) {
        if (ch < nchunks) {
            int slot = (int)(ch % NSLOT);
            size_t off, cnt;
            piece(ch, off, cnt);
            // slot reuse: its previous D2H must have drained (host copy-out below waits on it) — ensured
            // because chunk ch-NSLOT was fully retired in an earlier iteration.
            for (int k = 0; k < n_in; k++) {
                if (!in[k] || !in_sz[k]) continue;
                size_t bytes = cnt * in_sz[k];
                CU(c, c->d_in[slot][k].reserve(cap * in_sz[k]));
                const void* src = in[k] + off * in_sz[k];
                if (!in_pinned[k]) {
                    CU(c, c->h_in[slot][k].reserve(cap * in_sz[k]));
                    staged_copy(c->h_in[slot][k].p, src, bytes);
                    src = c->h_in[slot][k].p;
                }
                c->in_used[slot][k] = std::max(c->in_used[slot][k], bytes);
                CU(c, cudaMemcpyAsync(c->d_in[slot][k].p, src, bytes, cudaMemcpyHostToDevice, c->copy_in));
            }
            for (int k = 0; k < n_out; k++) {
                if (!out[k] || !out_sz[k]) continue;
                CU(c, c->d_out[slot][k].reserve(cap * out_sz[k]));
                if (!out_pinned[k]) CU(c, c->h_out[slot][k].reserve(cap * out_sz[k]));
            }
            CU(c, cudaEventRecord(c->ev_in[slot], c->copy_in));
            CU(c, cudaStreamWaitEvent(c->stream, c->ev_in[slot], 0));
            const uint8_t* di[4];
            uint8_t* dout_[3];
            for (int k = 0; k < 4; k++) di[k] = (k < n_in && in[k] && in_sz[k]) ? (const uint8_t*)c->d_in[slot][k].p : nullptr;
            for (int k = 0; k < 3; k++) dout_[k] = (k < n_out && out[k] && out_sz[k]) ? (uint8_t*)c->d_out[slot][k].p : nullptr;
            int r = enqueue(cnt, di, dout_, c->stream);
            if (r) return r;
            CU(c, cudaEventRecord(c->ev_done[slot], c->stream));
            CU(c, cudaStreamWaitEvent(c->copy_out, c->ev_done[slot], 0));
            for (int k = 0; k < n_out; k++) {
                if (!out[k] || !out_sz[k]) continue;
                void* dst = out_pinned[k] ? (void*)(out[k] + off * out_sz[k]) : c->h_out[slot][k].p;
                CU(c, cudaMemcpyAsync(dst, c->d_out[slot][k].p, cnt * out_sz[k], cudaMemcpyDeviceToHost, c->copy_out));
            }
            CU(c, cudaEventRecord(c->ev_out[slot], c->copy_out));
        }
        if (ch >= 1) {   // retire chunk ch-1: wait for its D2H and copy out of pinned memory
            size_t pc = ch - 1;
            int slot = (int)(pc % NSLOT);
            size_t off, cnt;
            piece(pc, off, cnt);
            CU(c, cudaEventSynchronize(c->ev_out[slot]));
            for (int k = 0; k < n_out; k++) {
                if (!out[k] || !out_sz[k] || out_pinned[k]) continue;
                staged_copy(out[k] + off * out_sz[k], c->h_out[slot][k].p, cnt * out_sz[k]);
            }
        }
    }
    return 0;
}

// The caller owns its buffers again the moment an entry point returns - also when it returns an error: a failure in the middle
// of the pipeline (allocation, launch, copy) must not leave DMA transfers from / into caller memory in flight.
template <class Fn>
int run_pipeline(ecb200_ctx* c, size_t n, int n_in, const uint8_t* const* in, const size_t* in_sz, int n_out, uint8_t* const* out,
                 const size_t* out_sz, Fn&& enqueue) {
    const int r = run_pipeline_pieces(c, n, n_in, in, in_sz, n_out, out, out_sz, enqueue);
    if (r) {
        cudaStreamSynchronize(c->copy_in);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_out);
        cudaGetLastError();      // the error that ended the pipeline is already in c->err
    }
    return r;
}

// Secret scalars (signing keys, nonces, FLAG_CT scalars) pass through the context's staging buffers, which live until
// ecb200_destroy and are reused by later calls.  The reference zeroizes such values (Zeroizing / ZeroizeOnDrop on
// SigningKey and the nonce); here the staged copies are overwritten as soon as the call that brought them has drained:
// device staging by cudaMemsetAsync, pinned host staging by memset.  mask: bit k = input k of the pipeline is secret.
int wipe_secret_inputs(ecb200_ctx* c, unsigned mask) {
    c->held_secrets = true;
    for (int slot = 0; slot < NSLOT; slot++)
        for (int k = 0; k < 4; k++) {
            const size_t used = c->in_used[slot][k];
            c->in_used[slot][k] = 0;
            if (!((mask >> k) & 1u) || !used) continue;
            if (c->d_in[slot][k].p) CU(c, cudaMemsetAsync(c->d_in[slot][k].p, 0, std::min(used, c->d_in[slot][k].cap), c->stream));
            if (c->h_in[slot][k].p) memset(c->h_in[slot][k].p, 0, std::min(used, c->h_in[slot][k].cap));
        }
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}
void forget_staged(ecb200_ctx* c) { memset(c->in_used, 0, sizeof(c->in_used)); }

// ecb200_init_multi: contiguous index shards [floor(i n / g), floor((i+1) n / g)) over the g child contexts,
// one host thread per device for the duration of the call; f(child, index, lo, count) runs the single-device entry point.
template <class F> int multi_run(ecb200_ctx* c, size_t n, bool call_empty, F&& f) {
    const size_t g = c->kids.size();
    std::vector<int> rc(g, 0);
    std::vector<std::thread> th;
    th.reserve(g);
    for (size_t i = 0; i < g; i++) {
        const size_t lo = i * n / g, hi = (i + 1) * n / g;
        if (hi == lo && !call_empty) continue;
        th.emplace_back([&rc, &f, c, i, lo, hi] { rc[i] = f(c->kids[i], i, lo, hi - lo); });
    }
    for (auto& t : th) t.join();
    for (size_t i = 0; i < g; i++)
        if (rc[i]) {
            c->err = "device " + std::to_string(c->kids[i]->device) + ": " + c->kids[i]->err;
            return rc[i];
        }
    return 0;
}
template <class T> T* at(T* p, size_t bytes) { return p ? p + bytes : nullptr; }

// per-row two-term linear combination: two products on the ordinary kernels, one complete addition, normalisation
int lincomb2_core(ecb200_ctx* c, const CurveLaunch* cl, size_t n, const uint8_t* d_p1, const uint8_t* d_k1, const uint8_t* d_p2, const uint8_t* d_k2,
                  uint8_t* d_out, uint8_t* d_invalid, uint32_t flags, cudaStream_t s) {
    CU(c, c->proj.reserve(n * 3 * (size_t)cl->L * 4));
    CU(c, c->proj2.reserve(n * 3 * (size_t)cl->L * 4));
    CU(c, c->inv2.reserve(2 * n));
    uint32_t *pa = (uint32_t*)c->proj.p, *pb = (uint32_t*)c->proj2.p;
    uint8_t* inv_a = d_invalid ? d_invalid : (uint8_t*)c->inv2.p + n;
    uint8_t* inv_b = (uint8_t*)c->inv2.p;
    const bool proj = (flags & ECB200_FLAG_PROJ) != 0;
    const uint8_t* pts[2] = {d_p1, d_p2};
    const uint8_t* ks[2] = {d_k1, d_k2};
    uint32_t* dst[2] = {pa, pb};
    uint8_t* inv[2] = {inv_a, inv_b};
    for (int t = 0; t < 2; t++) {
        if (flags & ECB200_FLAG_CT) {
            cl->mul_var(s, true, (int)n, proj ? ECB200_FLAG_PROJ : 0, pts[t], nullptr, ks[t], dst[t], inv[t]);
            continue;
        }
        // public scalars: the Jacobian fast path (projective inputs are normalised first, as in mul_var_core); the
        // context's affine / window-table scratch is reused by the second product, stream-ordered after the first
        const uint32_t* aff = nullptr;
        if (proj) {
            CU(c, c->aff.reserve(n * 2 * (size_t)cl->L * 4));
            cl->load_proj(s, (int)n, pts[t], dst[t], nullptr);
            cl->normalize(s, (int)n, dst[t], 2 /*NORM_AFF_LIMBS*/, 0, nullptr, nullptr, (uint32_t*)c->aff.p);
            aff = (const uint32_t*)c->aff.p;
        }
        uint32_t* wt = nullptr;
        int r = window_tables(c, cl, n, proj ? nullptr : pts[t], aff, s, &wt);
        if (r) return r;
        cl->mul_var_fast(s, (int)n, proj ? nullptr : pts[t], aff, nullptr, ks[t], dst[t], inv[t], wt);
    }
    cl->add_proj(s, (int)n, pa, pb, pa, inv_a, inv_b);
    cl->normalize(s, (int)n, pa, 0, resolve_compress(cl, flags) ? 1 : 0, d_out, nullptr, nullptr);
    CU(c, cudaGetLastError());
    return 0;
}

}  // namespace

// ===============================================================================================
extern "C" {

size_t ecb200_field_bytes(int curve) {
    return curve == ECB200_P384 ? 48 : curve == ECB200_P192 ? 24 : curve == ECB200_P224 ? 28 : (curve >= 0 && curve <= 3 ? 32 : 0);
}
size_t ecb200_point_slot_bytes(int curve, uint32_t flags) {
    size_t fb = ecb200_field_bytes(curve);
    if (!fb) return 0;
    bool comp = (flags & ECB200_FLAG_COMPRESSED) ? true : (flags & ECB200_FLAG_UNCOMPRESSED) ? false : (curve == ECB200_K256);
    return 1 + (comp ? fb : 2 * fb);
}
const char* ecb200_version(void) { return "ecb200 0.2 (sm_100a)"; }

static void set_launchers(ecb200_ctx* c) {
    c->cl[0] = launch_k256();
    c->cl[1] = launch_p256();
    c->cl[2] = launch_p384();
    c->cl[3] = launch_sm2();
    c->cl[4] = launch_p192();
    c->cl[5] = launch_p224();
}

int ecb200_init(int device, ecb200_ctx** out) {
    if (!out) return ECB200_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ECB200_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ECB200_ERR_CUDA;
    ecb200_ctx* c = new ecb200_ctx();
    Enter enter_(c);
    c->device = device;
    set_launchers(c);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < NSLOT; i++)
        ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
    if (const char* e = getenv("ECB200_GW")) { int g = atoi(e); if (g >= 2 && g <= 24) c->gw = g; }
    if (const char* e = getenv("ECB200_VERIFY_V1")) c->verify_v1 = atoi(e) != 0;
    if (const char* e = getenv("ECB200_WINTAB")) c->use_wintab = atoi(e) != 0;
    if (const char* e = getenv("ECB200_GEN2")) c->use_gen2 = atoi(e) != 0;
    if (const char* e = getenv("ECB200_KEYTAB")) c->use_keytab = atoi(e) != 0;
    if (const char* e = getenv("ECB200_KT_WIDE")) c->kt_force_wide = atoi(e) != 0 ? 1 : 0;
    if (!ok || build_tables(c) != 0) {
        fprintf(stderr, "ecb200_init failed: %s (%s)\n", c->err.c_str(), cudaGetErrorString(cudaGetLastError()));
        ecb200_destroy(c);
        return ECB200_ERR_CUDA;
    }
    c->launches = 0;     // table construction is not the caller's work
    *out = c;
    return 0;
}

int ecb200_init_multi(int n_dev, const int* devices, ecb200_ctx** out) {
    if (!out) return ECB200_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return ECB200_ERR_CUDA;
    if (n_dev <= 0) { if (devices) return ECB200_ERR_ARG; n_dev = ndev; }
    if (n_dev > 64) return ECB200_ERR_ARG;
    for (int i = 0; devices && i < n_dev; i++)     // an ordinal may repeat: two shards (two child contexts) on one device
        if (devices[i] < 0 || devices[i] >= ndev) return ECB200_ERR_ARG;
    if (!devices && n_dev > ndev) return ECB200_ERR_CUDA;
    ecb200_ctx* c = new ecb200_ctx();
    c->device = -1;
    set_launchers(c);
    for (int i = 0; i < n_dev; i++) {
        ecb200_ctx* kid = nullptr;
        int r = ecb200_init(devices ? devices[i] : i, &kid);
        if (r) { ecb200_destroy(c); return r; }
        c->kids.push_back(kid);
    }
    *out = c;
    return 0;
}
int ecb200_device_count(const ecb200_ctx* c) { return !c ? 0 : c->kids.empty() ? 1 : (int)c->kids.size(); }

void ecb200_destroy(ecb200_ctx* c) {
    if (!c) return;
    if (!c->kids.empty() || c->device < 0) {
        for (ecb200_ctx* k : c->kids) ecb200_destroy(k);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->held_secrets)   // belt and braces: secret staging was wiped after each call (wipe_secret_inputs)
        for (int s = 0; s < NSLOT; s++)
            for (int k = 0; k < 4; k++) {
                if (c->d_in[s][k].p) cudaMemset(c->d_in[s][k].p, 0, c->d_in[s][k].cap);
                if (c->h_in[s][k].p) memset(c->h_in[s][k].p, 0, c->h_in[s][k].cap);
            }
    for (int i = 0; i < NCURVE; i++) {
        if (c->gtab[i]) cudaFree(c->gtab[i]);
        if (c->gentab[i]) cudaFree(c->gentab[i]);
        if (c->gentab2[i]) cudaFree(c->gentab2[i]);
        if (c->gbig[i]) cudaFree(c->gbig[i]);
    }
    c->prep.release();
    c->aff.release();
    c->wtab.release();
    c->kxy.release();
    c->kst.release();
    c->proj.release();
    c->proj2.release();
    c->inv2.release();
    for (DevBuf* b : {&c->kt.htab, &c->kt.gkeys, &c->kt.tab, &c->kt.kvalid, &c->kt.jac, &c->kt.rep, &c->kt.rep_slot, &c->kt.newgid, &c->kt.gid, &c->kt.counter}) b->release();
    c->partial.release();
    c->one_point.release();
    for (int s = 0; s < NSLOT; s++) {
        for (int k = 0; k < 4; k++) { c->d_in[s][k].release(); c->h_in[s][k].release(); }
        for (int k = 0; k < 3; k++) { c->d_out[s][k].release(); c->h_out[s][k].release(); }
        if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
        if (c->ev_done[s]) cudaEventDestroy(c->ev_done[s]);
        if (c->ev_out[s]) cudaEventDestroy(c->ev_out[s]);
    }
    for (auto& pr : c->timed) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    delete c;
}

const char* ecb200_last_error(const ecb200_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t ecb200_launch_count(const ecb200_ctx* c) {
    if (!c) return 0;
    uint64_t t = c->launches;
    for (const ecb200_ctx* k : c->kids) t += k->launches;
    return t;
}
int ecb200_keytab_stats(const ecb200_ctx* c, uint64_t* rows, uint64_t* tables) {
    if (!c || !rows || !tables) return ECB200_ERR_ARG;
    *rows = c->kt_rows;
    *tables = c->kt_groups;
    for (const ecb200_ctx* k : c->kids) { *rows += k->kt_rows; *tables += k->kt_groups; }
    return 0;
}
int ecb200_sync(ecb200_ctx* c) {
    if (!c) return ECB200_ERR_ARG;
    for (ecb200_ctx* k : c->kids) { int r = ecb200_sync(k); if (r) return r; }
    if (!c->kids.empty()) return 0;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ecb200_kernel_timing(ecb200_ctx* c, int enable) {
    if (!c) return ECB200_ERR_ARG;
    c->timing = enable != 0;
    for (ecb200_ctx* k : c->kids) k->timing = enable != 0;
    return 0;
}
int ecb200_kernel_timing_read(ecb200_ctx* c, double* total_ms, uint64_t* launches) {
    if (!c || !total_ms || !launches) return ECB200_ERR_ARG;
    if (!c->kids.empty()) {   // sum over the devices (each device times its own launches)
        *total_ms = 0; *launches = 0;
        for (ecb200_ctx* k : c->kids) {
            double ms = 0; uint64_t cnt = 0;
            int r = ecb200_kernel_timing_read(k, &ms, &cnt);
            if (r) return r;
            *total_ms += ms; *launches += cnt;
        }
        return 0;
    }
    CU(c, cudaSetDevice(c->device));
    double sum = 0;
    for (auto& pr : c->timed) {
        CU(c, cudaEventSynchronize(pr.second));
        float ms = 0;
        CU(c, cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    *total_ms = sum;
    *launches = c->timed.size();
    c->timed.clear();
    return 0;
}

// device pointers belong to one device: the _dev entry points refuse a multi-device context
#define DEV_ENTER(c, name)                                                                                     \
    if ((c) && !(c)->kids.empty()) return fail(c, ECB200_ERR_ARG, name ": device-pointer calls need a single-device context"); \
    Enter enter_(c)

// ---- device-pointer entry points
int ecb200_mul_gen_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_k, uint8_t* d_out, uint32_t flags, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_k || !d_out))) return fail(c, ECB200_ERR_ARG, "mul_gen_dev: bad argument");
    DEV_ENTER(c, "mul_gen_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return mul_gen_core(c, cl, n, d_k, d_out, flags, pick(c, stream));
}
int ecb200_mul_var_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_pts, const uint8_t* d_inf, const uint8_t* d_k,
                       uint8_t* d_out, uint8_t* d_invalid, uint32_t flags, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_pts || !d_k || !d_out))) return fail(c, ECB200_ERR_ARG, "mul_var_dev: bad argument");
    DEV_ENTER(c, "mul_var_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return mul_var_core(c, cl, n, d_pts, d_inf, d_k, d_out, d_invalid, flags, pick(c, stream));
}
int ecb200_batch_normalize_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_xyz, uint8_t* d_xy, uint8_t* d_inf, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_xyz || !d_xy))) return fail(c, ECB200_ERR_ARG, "batch_normalize_dev: bad argument");
    DEV_ENTER(c, "batch_normalize_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return batch_normalize_core(c, cl, n, d_xyz, d_xy, d_inf, pick(c, stream));
}
int ecb200_ecdsa_verify_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_q, const uint8_t* d_z, const uint8_t* d_rs,
                            uint8_t* d_ok, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_q || !d_z || !d_rs || !d_ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_verify_dev: bad argument");
    DEV_ENTER(c, "ecdsa_verify_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    int r = kt_begin(c, cl, n, n, pick(c, stream));
    if (!r) r = verify_core(c, cl, n, d_q, d_z, d_rs, d_ok, pick(c, stream));
    kt_end(c);
    return r;
}
int ecb200_lincomb2_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_p1, const uint8_t* d_k1, const uint8_t* d_p2, const uint8_t* d_k2,
                        uint8_t* d_out, uint8_t* d_invalid, uint32_t flags, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_p1 || !d_k1 || !d_p2 || !d_k2 || !d_out)))
        return fail(c, ECB200_ERR_ARG, "lincomb2_dev: bad argument");
    DEV_ENTER(c, "lincomb2_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return lincomb2_core(c, cl, n, d_p1, d_k1, d_p2, d_k2, d_out, d_invalid, flags, pick(c, stream));
}

// ---- host-pointer entry points
int ecb200_mul_gen(ecb200_ctx* c, int curve, size_t n, const uint8_t* k, uint8_t* out, uint32_t flags) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!k || !out))) return fail(c, ECB200_ERR_ARG, "mul_gen: bad argument");
    const size_t FB = cl->FB, SB = slot_bytes(cl, flags);
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_mul_gen(kid, curve, cnt, k + lo * FB, out + lo * SB, flags);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[1] = {k};
    size_t in_sz[1] = {FB};
    uint8_t* o[1] = {out};
    size_t o_sz[1] = {SB};
    forget_staged(c);
    int r = run_pipeline(c, n, 1, in, in_sz, 1, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return mul_gen_core(c, cl, cnt, di[0], dout[0], flags, s);
    });
    if ((flags & ECB200_FLAG_CT) && n) { int w = wipe_secret_inputs(c, 1u); if (!r) r = w; }
    return r;
}
int ecb200_mul_var(ecb200_ctx* c, int curve, size_t n, const uint8_t* pts, const uint8_t* inf, const uint8_t* k, uint8_t* out,
                   uint8_t* invalid, uint32_t flags) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!pts || !k || !out))) return fail(c, ECB200_ERR_ARG, "mul_var: bad argument");
    const bool proj = (flags & ECB200_FLAG_PROJ) != 0;
    const size_t FB = cl->FB, PB = FB * (proj ? 3 : 2), SB = slot_bytes(cl, flags);
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_mul_var(kid, curve, cnt, pts + lo * PB, at(inf, lo), k + lo * FB, out + lo * SB, at(invalid, lo), flags);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {pts, k, proj ? nullptr : inf};
    size_t in_sz[3] = {PB, FB, 1};
    uint8_t* o[2] = {out, invalid};
    size_t o_sz[2] = {SB, 1};
    forget_staged(c);
    int r = run_pipeline(c, n, 3, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return mul_var_core(c, cl, cnt, di[0], di[2], di[1], dout[0], dout[1], flags, s);
    });
    if ((flags & ECB200_FLAG_CT) && n) { int w = wipe_secret_inputs(c, 2u); if (!r) r = w; }
    return r;
}
int ecb200_batch_normalize(ecb200_ctx* c, int curve, size_t n, const uint8_t* xyz, uint8_t* xy, uint8_t* inf) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!xyz || !xy))) return fail(c, ECB200_ERR_ARG, "batch_normalize: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_batch_normalize(kid, curve, cnt, xyz + lo * 3 * FB, xy + lo * 2 * FB, at(inf, lo));
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[1] = {xyz};
    size_t in_sz[1] = {FB * 3};
    uint8_t* o[2] = {xy, inf};
    size_t o_sz[2] = {FB * 2, 1};
    return run_pipeline(c, n, 1, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return batch_normalize_core(c, cl, cnt, di[0], dout[0], dout[1], s);
    });
}
int ecb200_ecdsa_verify(ecb200_ctx* c, int curve, size_t n, const uint8_t* q, const uint8_t* z, const uint8_t* rs, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!q || !z || !rs || !ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_verify: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_ecdsa_verify(kid, curve, cnt, q + lo * 2 * FB, z + lo * FB, rs + lo * 2 * FB, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {q, z, rs};
    size_t in_sz[3] = {FB * 2, FB, FB * 2};
    uint8_t* o[1] = {ok};
    size_t o_sz[1] = {1};
    int r = kt_begin(c, cl, n, std::min(n, CHUNK), c->stream);
    if (!r) r = run_pipeline(c, n, 3, in, in_sz, 1, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return verify_core(c, cl, cnt, di[0], di[1], di[2], dout[0], s);
    });
    kt_end(c);
    return r;
}
int ecb200_field_op(ecb200_ctx* c, int curve, int which, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || which < 0 || which > 1 || op < 0 || op > 7 || (n && (!a || !out || !ok))) return fail(c, ECB200_ERR_ARG, "field_op: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_field_op(kid, curve, which, op, cnt, a + lo * FB, at(b, lo * FB), out + lo * FB, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[2] = {a, b};
    size_t in_sz[2] = {FB, FB};
    uint8_t* o[2] = {out, ok};
    size_t o_sz[2] = {FB, 1};
    return run_pipeline(c, n, 2, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        cl->field_op(s, (int)cnt, which, op, di[0], di[1], dout[0], dout[1]);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : fail(c, ECB200_ERR_CUDA, "field_op launch", e);
    });
}
int ecb200_lincomb2(ecb200_ctx* c, int curve, size_t n, const uint8_t* p1, const uint8_t* k1, const uint8_t* p2, const uint8_t* k2, uint8_t* out,
                    uint8_t* invalid, uint32_t flags) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!p1 || !k1 || !p2 || !k2 || !out))) return fail(c, ECB200_ERR_ARG, "lincomb2: bad argument");
    const size_t FB = cl->FB, SB = slot_bytes(cl, flags), PB = FB * ((flags & ECB200_FLAG_PROJ) ? 3 : 2);
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_lincomb2(kid, curve, cnt, p1 + lo * PB, k1 + lo * FB, p2 + lo * PB, k2 + lo * FB, out + lo * SB, at(invalid, lo), flags);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[4] = {p1, k1, p2, k2};
    size_t in_sz[4] = {PB, FB, PB, FB};
    uint8_t* o[2] = {out, invalid};
    size_t o_sz[2] = {SB, 1};
    forget_staged(c);
    int r = run_pipeline(c, n, 4, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return lincomb2_core(c, cl, cnt, di[0], di[1], di[2], di[3], dout[0], dout[1], flags, s);
    });
    if ((flags & ECB200_FLAG_CT) && n) { int w = wipe_secret_inputs(c, 2u | 8u); if (!r) r = w; }
    return r;
}
int ecb200_lincomb(ecb200_ctx* c, int curve, size_t n_terms, const uint8_t* pts, const uint8_t* k, uint8_t* out_point, uint32_t flags,
                   uint32_t out_flags_proj) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || !out_point || (n_terms && (!pts || !k))) return fail(c, ECB200_ERR_ARG, "lincomb: bad argument");
    const int L = cl->L, FB = cl->FB;
    const bool proj = (flags & ECB200_FLAG_PROJ) != 0;
    const size_t psz = (size_t)FB * (proj ? 3 : 2);
    if (!c->kids.empty()) {
        // every device reduces its index shard of the terms to ONE projective partial; the <= 8 partials (3 FB bytes each)
        // come back to the host and the first device adds them (scalars = 1) - the only exchange on this path
        const size_t g = c->kids.size();
        std::vector<uint8_t> partials(g * 3 * FB), ones(g * FB, 0);
        for (size_t i = 0; i < g; i++) ones[i * FB + FB - 1] = 1;
        int r = multi_run(c, n_terms, true, [&](ecb200_ctx* kid, size_t i, size_t lo, size_t cnt) {
            return ecb200_lincomb(kid, curve, cnt, at(pts, lo * psz), at(k, lo * FB), &partials[i * 3 * FB], flags, ECB200_FLAG_PROJ);
        });
        if (r) return r;
        r = ecb200_lincomb(c->kids[0], curve, g, partials.data(), ones.data(), out_point, (flags & ~ECB200_FLAG_CT) | ECB200_FLAG_PROJ, out_flags_proj);
        if (r) c->err = c->kids[0]->err;
        return r;
    }
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CU(c, c->proj.reserve(std::max<size_t>(n_terms, 1) * 3 * L * 4));
    CU(c, c->partial.reserve((size_t)cl->sum_blocks * 3 * L * 4));
    CU(c, c->one_point.reserve((size_t)3 * L * 4 + 3 * FB + 256));
    CU(c, c->inv2.reserve(std::max<size_t>(n_terms, 1)));
    // terms are processed in CHUNK pieces; each piece's products land in c->proj at its offset
    for (size_t off = 0; off < n_terms; off += CHUNK) {
        size_t cnt = std::min(CHUNK, n_terms - off);
        CU(c, c->d_in[0][0].reserve(cnt * psz));
        CU(c, c->d_in[0][1].reserve(cnt * FB));
        CU(c, cudaMemcpyAsync(c->d_in[0][0].p, pts + off * psz, cnt * psz, cudaMemcpyHostToDevice, s));
        CU(c, cudaMemcpyAsync(c->d_in[0][1].p, k + off * FB, cnt * FB, cudaMemcpyHostToDevice, s));
        uint8_t* inv = (uint8_t*)c->inv2.p + off;
        if ((flags & ECB200_FLAG_CT) || proj)
            cl->mul_var(s, (flags & ECB200_FLAG_CT) != 0, (int)cnt, flags, (const uint8_t*)c->d_in[0][0].p, nullptr,
                        (const uint8_t*)c->d_in[0][1].p, (uint32_t*)c->proj.p + off * 3 * L, inv);
        else
            cl->mul_var_fast(s, (int)cnt, (const uint8_t*)c->d_in[0][0].p, nullptr, nullptr, (const uint8_t*)c->d_in[0][1].p,
                             (uint32_t*)c->proj.p + off * 3 * L, inv, nullptr);
        if (flags & ECB200_FLAG_CT) CU(c, cudaMemsetAsync(c->d_in[0][1].p, 0, cnt * FB, s));   // secret scalars do not outlive their piece
        CU(c, cudaStreamSynchronize(s));   // staging buffers are reused by the next piece
    }
    // A term that fails validation (coordinate >= p, affine point off the curve) cannot be represented in the reference at
    // all; dropping it silently would return a wrong sum with status 0, so the whole call fails instead.
    if (n_terms) {
        std::vector<uint8_t> inv(n_terms);
        CU(c, cudaMemcpyAsync(inv.data(), c->inv2.p, n_terms, cudaMemcpyDeviceToHost, s));
        CU(c, cudaStreamSynchronize(s));
        for (size_t i = 0; i < n_terms; i++)
            if (inv[i]) return fail(c, ECB200_ERR_POINT, ("lincomb: term " + std::to_string(i) + " is not a valid point").c_str());
    }
    uint32_t* d_sum = (uint32_t*)c->one_point.p;
    uint8_t* d_bytes = (uint8_t*)c->one_point.p + 3 * L * 4;
    cl->sum(s, (int)n_terms, (const uint32_t*)c->proj.p, (uint32_t*)c->partial.p, d_sum);
    size_t out_bytes;
    if (out_flags_proj & ECB200_FLAG_PROJ) {
        cl->proj_to_bytes(s, 1, d_sum, d_bytes);
        out_bytes = (size_t)3 * FB;
    } else {
        cl->normalize(s, 1, d_sum, 0, resolve_compress(cl, flags) ? 1 : 0, d_bytes, nullptr, nullptr);
        out_bytes = slot_bytes(cl, flags);
    }
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(out_point, d_bytes, out_bytes, cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    return 0;
}


// ---- SEC1 decoding, verification with encoded keys, recovery, BIP340, SM2DSA, signing
int ecb200_decode_points_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_enc, size_t stride, uint32_t mode, uint8_t* d_xy, uint8_t* d_status,
                             void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || mode > 1 || (n && (!d_enc || !d_xy || !d_status)) || stride < (size_t)cl->FB + (mode == DEC_SEC1 ? 1 : 0))
        return fail(c, ECB200_ERR_ARG, "decode_points_dev: bad argument");
    DEV_ENTER(c, "decode_points_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return decode_core(c, cl, n, d_enc, stride, (int)mode, d_xy, d_status, pick(c, stream));
}
int ecb200_decode_points(ecb200_ctx* c, int curve, size_t n, const uint8_t* enc, size_t stride, uint32_t mode, uint8_t* xy, uint8_t* status) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || mode > 1 || (n && (!enc || !xy || !status)) || stride < (size_t)cl->FB + (mode == DEC_SEC1 ? 1 : 0))
        return fail(c, ECB200_ERR_ARG, "decode_points: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_decode_points(kid, curve, cnt, enc + lo * stride, stride, mode, xy + lo * 2 * FB, status + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[1] = {enc};
    size_t in_sz[1] = {stride};
    uint8_t* o[2] = {xy, status};
    size_t o_sz[2] = {FB * 2, 1};
    return run_pipeline(c, n, 1, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return decode_core(c, cl, cnt, di[0], stride, (int)mode, dout[0], dout[1], s);
    });
}
int ecb200_ecdsa_verify_sec1_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_keys, size_t key_stride, const uint8_t* d_z, const uint8_t* d_rs,
                                 uint8_t* d_ok, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_keys || !d_z || !d_rs || !d_ok)) || key_stride < (size_t)cl->FB + 1) return fail(c, ECB200_ERR_ARG, "ecdsa_verify_sec1_dev: bad argument");
    DEV_ENTER(c, "ecdsa_verify_sec1_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    int r = kt_begin(c, cl, n, n, pick(c, stream));
    if (!r) r = verify_sec1_core(c, cl, n, d_keys, key_stride, d_z, d_rs, d_ok, pick(c, stream));
    kt_end(c);
    return r;
}
int ecb200_ecdsa_verify_sec1(ecb200_ctx* c, int curve, size_t n, const uint8_t* keys, size_t key_stride, const uint8_t* z, const uint8_t* rs, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!keys || !z || !rs || !ok)) || key_stride < (size_t)cl->FB + 1) return fail(c, ECB200_ERR_ARG, "ecdsa_verify_sec1: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_ecdsa_verify_sec1(kid, curve, cnt, keys + lo * key_stride, key_stride, z + lo * FB, rs + lo * 2 * FB, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {keys, z, rs};
    size_t in_sz[3] = {key_stride, FB, FB * 2};
    uint8_t* o[1] = {ok};
    size_t o_sz[1] = {1};
    int r = kt_begin(c, cl, n, std::min(n, CHUNK), c->stream);
    if (!r) r = run_pipeline(c, n, 3, in, in_sz, 1, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return verify_sec1_core(c, cl, cnt, di[0], key_stride, di[1], di[2], dout[0], s);
    });
    kt_end(c);
    return r;
}
int ecb200_ecdsa_recover_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_z, const uint8_t* d_rs, const uint8_t* d_recid, uint8_t* d_keys,
                             uint8_t* d_ok, uint32_t flags, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_z || !d_rs || !d_recid || !d_keys || !d_ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_recover_dev: bad argument");
    DEV_ENTER(c, "ecdsa_recover_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return recover_core(c, cl, n, d_z, d_rs, d_recid, d_keys, d_ok, flags, pick(c, stream));
}
int ecb200_ecdsa_recover(ecb200_ctx* c, int curve, size_t n, const uint8_t* z, const uint8_t* rs, const uint8_t* recid, uint8_t* keys, uint8_t* ok,
                         uint32_t flags) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!z || !rs || !recid || !keys || !ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_recover: bad argument");
    const size_t FB = cl->FB, SB = slot_bytes(cl, flags);
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_ecdsa_recover(kid, curve, cnt, z + lo * FB, rs + lo * 2 * FB, recid + lo, keys + lo * SB, ok + lo, flags);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {z, rs, recid};
    size_t in_sz[3] = {FB, FB * 2, 1};
    uint8_t* o[2] = {keys, ok};
    size_t o_sz[2] = {SB, 1};
    return run_pipeline(c, n, 3, in, in_sz, 2, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return recover_core(c, cl, cnt, di[0], di[1], di[2], dout[0], dout[1], flags, s);
    });
}
int ecb200_schnorr_verify_dev(ecb200_ctx* c, size_t n, const uint8_t* d_pk, const uint8_t* d_e, const uint8_t* d_sig, uint8_t* d_ok, void* stream) {
    const CurveLaunch* cl = curve_of(c, ECB200_K256);
    if (!cl || n > MAX_ROWS || (n && (!d_pk || !d_e || !d_sig || !d_ok))) return fail(c, ECB200_ERR_ARG, "schnorr_verify_dev: bad argument");
    DEV_ENTER(c, "schnorr_verify_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return schnorr_core(c, cl, n, d_pk, d_e, d_sig, d_ok, pick(c, stream));
}
int ecb200_schnorr_verify(ecb200_ctx* c, size_t n, const uint8_t* pk, const uint8_t* e, const uint8_t* sig, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, ECB200_K256);
    if (!cl || (n && (!pk || !e || !sig || !ok))) return fail(c, ECB200_ERR_ARG, "schnorr_verify: bad argument");
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_schnorr_verify(kid, cnt, pk + lo * 32, e + lo * 32, sig + lo * 64, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {pk, e, sig};
    size_t in_sz[3] = {32, 32, 64};
    uint8_t* o[1] = {ok};
    size_t o_sz[1] = {1};
    return run_pipeline(c, n, 3, in, in_sz, 1, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return schnorr_core(c, cl, cnt, di[0], di[1], di[2], dout[0], s);
    });
}
int ecb200_sm2dsa_verify_dev(ecb200_ctx* c, size_t n, const uint8_t* d_q, const uint8_t* d_e, const uint8_t* d_rs, uint8_t* d_ok, void* stream) {
    const CurveLaunch* cl = curve_of(c, ECB200_SM2);
    if (!cl || n > MAX_ROWS || (n && (!d_q || !d_e || !d_rs || !d_ok))) return fail(c, ECB200_ERR_ARG, "sm2dsa_verify_dev: bad argument");
    DEV_ENTER(c, "sm2dsa_verify_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    int r = kt_begin(c, cl, n, n, pick(c, stream));
    if (!r) r = verify_core(c, cl, n, d_q, d_e, d_rs, d_ok, pick(c, stream), VM_SM2DSA);
    kt_end(c);
    return r;
}
int ecb200_sm2dsa_verify(ecb200_ctx* c, size_t n, const uint8_t* q, const uint8_t* e, const uint8_t* rs, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, ECB200_SM2);
    if (!cl || (n && (!q || !e || !rs || !ok))) return fail(c, ECB200_ERR_ARG, "sm2dsa_verify: bad argument");
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_sm2dsa_verify(kid, cnt, q + lo * 64, e + lo * 32, rs + lo * 64, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {q, e, rs};
    size_t in_sz[3] = {64, 32, 64};
    uint8_t* o[1] = {ok};
    size_t o_sz[1] = {1};
    int r = kt_begin(c, cl, n, std::min(n, CHUNK), c->stream);
    if (!r) r = run_pipeline(c, n, 3, in, in_sz, 1, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return verify_core(c, cl, cnt, di[0], di[1], di[2], dout[0], s, VM_SM2DSA);
    });
    kt_end(c);
    return r;
}
int ecb200_ecdsa_sign_dev(ecb200_ctx* c, int curve, size_t n, const uint8_t* d_d, const uint8_t* d_k, const uint8_t* d_z, uint8_t* d_rs, uint8_t* d_recid,
                          uint8_t* d_ok, void* stream) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || n > MAX_ROWS || (n && (!d_d || !d_k || !d_z || !d_rs || !d_recid || !d_ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_sign_dev: bad argument");
    DEV_ENTER(c, "ecdsa_sign_dev");
    if (!n) return 0;
    CU(c, cudaSetDevice(c->device));
    return sign_core(c, cl, n, d_d, d_k, d_z, d_rs, d_recid, d_ok, pick(c, stream));
}
int ecb200_ecdsa_sign(ecb200_ctx* c, int curve, size_t n, const uint8_t* d, const uint8_t* k, const uint8_t* z, uint8_t* rs, uint8_t* recid, uint8_t* ok) {
    const CurveLaunch* cl = curve_of(c, curve);
    if (!cl || (n && (!d || !k || !z || !rs || !recid || !ok))) return fail(c, ECB200_ERR_ARG, "ecdsa_sign: bad argument");
    const size_t FB = cl->FB;
    if (!c->kids.empty())
        return multi_run(c, n, false, [&](ecb200_ctx* kid, size_t, size_t lo, size_t cnt) {
            return ecb200_ecdsa_sign(kid, curve, cnt, d + lo * FB, k + lo * FB, z + lo * FB, rs + lo * 2 * FB, recid + lo, ok + lo);
        });
    Enter enter_(c);
    CU(c, cudaSetDevice(c->device));
    const uint8_t* in[3] = {d, k, z};
    size_t in_sz[3] = {FB, FB, FB};
    uint8_t* o[3] = {rs, recid, ok};
    size_t o_sz[3] = {FB * 2, 1, 1};
    forget_staged(c);
    int r = run_pipeline(c, n, 3, in, in_sz, 3, o, o_sz, [&](size_t cnt, const uint8_t* const* di, uint8_t* const* dout, cudaStream_t s) {
        return sign_core(c, cl, cnt, di[0], di[1], di[2], dout[0], dout[1], dout[2], s);
    });
    if (n) { int w = wipe_secret_inputs(c, 1u | 2u); if (!r) r = w; }   // d and k (Zeroizing / ZeroizeOnDrop in the reference)
    return r;
}

}  // extern "C"
