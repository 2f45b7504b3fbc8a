// ecb200.hpp — C++17 host-side mirror of the reference's trait surface over the C ABI (ecb200.h).
//
// The reference (risc0 fork of RustCrypto `elliptic-curves`) is Rust; this image has no cargo / rustc, so the host
// side above the C ABI is written in C++ (the reference is compiled code) with the reference's own names, argument
// meaning and error behaviour for the batch-able hot path.  Header-only; link with -lecb200.
//
//   reference item (file:line)                                                  here
//   --------------------------------------------------------------------------  -----------------------------------------
//   k256::Secp256k1 / p256::NistP256 / p384::NistP384 / sm2::Sm2 / p192::NistP192 curve markers (Curve::ORDER, FieldBytesSize)
//     (k256/src/lib.rs:76-111, p256/src/lib.rs:74-120, p384/src/lib.rs:50-76, sm2/src/lib.rs:60-85, p192/src/lib.rs:42-66)
//   Scalar: PrimeField::{from_repr,to_repr}, Reduce::reduce_bytes, IsHigh        Scalar<C>
//     (k256/src/arithmetic/scalar.rs:340-377,519-523,700-713)
//   AffinePoint {x, y, infinity}, ToEncodedPoint, AffineCoordinates              AffinePoint<C>, EncodedPoint<C>
//     (k256/src/arithmetic/affine.rs:37-75,110-120,272-284; primeorder/src/affine.rs:38-47,233-243)
//   FromEncodedPoint / DecompressPoint / DecompactPoint                          AffinePoint<C>::from_encoded_points, decompact_batch
//     (k256 affine.rs:184-211,241-270; primeorder affine.rs:129-195)
//   ProjectivePoint {x, y, z}, Group::{identity, generator is on the device}     ProjectivePoint<C>
//     (k256/src/arithmetic/projective.rs:38-42; primeorder/src/projective.rs:37-41)
//   MulByGenerator::mul_by_generator (k256 mul.rs:415-440; primeorder            ProjectivePoint<C>::mul_by_generator_batch
//     projective.rs:422-431)
//   Mul<&Scalar> for &ProjectivePoint (k256 mul.rs:443-481; primeorder           ProjectivePoint<C>::mul_batch (constant time),
//     projective.rs:106-150)                                                       mul_batch_vartime (public scalars)
//   BatchNormalize<[ProjectivePoint]> / group::Curve::batch_normalize            ProjectivePoint<C>::batch_normalize
//     (k256 projective.rs:325-379,519-525; primeorder projective.rs:346-413)
//   LinearCombinationExt::lincomb_ext, LinearCombination::lincomb                ProjectivePoint<C>::lincomb_ext, lincomb
//     (k256 mul.rs:313-393; primeorder projective.rs:415-420)
//   ecdsa::Signature::{from_scalars, from_slice, normalize_s} (ecdsa 0.16.9)     ecdsa::Signature<C>
//   ecdsa::VerifyingKey::{from_affine, from_sec1_bytes, verify_prehash,          ecdsa::VerifyingKey<C>::{..., verify_prehash_batch,
//     recover_from_prehash} (k256/src/ecdsa.rs:113-140,200-209,278-343;            recover_from_prehash_batch}
//     p256/src/ecdsa.rs:71-75)
//   SignPrimitive::try_sign_prehashed (k256/src/ecdsa.rs:181-198)                ecdsa::try_sign_prehashed_batch
//   schnorr::VerifyingKey (k256/src/schnorr/verifying.rs:35-89)                  schnorr::VerifyingKey::verify_raw_batch
//   sm2::dsa::VerifyingKey::verify_prehash (sm2/src/dsa/verifying.rs:130-168)    sm2dsa::verify_prehash_batch
//
// Single-element calls stay on the reference CPU types; everything here is the *batch* form ("swap a slice-of-inputs
// loop for one call").  Error behaviour: `CtOption<T>` -> std::optional<T>, `Result<(), signature::Error>` ->
// ecb200::Result (opaque error, no variants — exactly what signature::Error carries), a failing engine (no CUDA device,
// launch error) -> ecb200::Error exception.  There is NO CPU fallback: the only arithmetic on the host is byte
// comparison / subtraction against the group order for `from_repr` / `reduce_bytes` / `is_high`, which the reference
// also does when a `Scalar` or `Signature` is *constructed*, before the hot path starts.
#ifndef ECB200_HPP
#define ECB200_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ecb200.h"

namespace ecb200 {

// ---------------------------------------------------------------------------------------------------------------
// errors

// engine-level failure (ecb200_status != 0): bad argument, CUDA error, out of memory
class Error : public std::runtime_error {
public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }
private:
    int status_;
};

// Result<(), signature::Error>: the error type of the `signature` crate is opaque (no variants)
class Result {
public:
    static Result Ok() { return Result(true); }
    static Result Err() { return Result(false); }
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    explicit operator bool() const { return ok_; }
    bool operator==(const Result& o) const { return ok_ == o.ok_; }
private:
    explicit Result(bool ok) : ok_(ok) {}
    bool ok_;
};

namespace detail {
constexpr int hexval(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c - 'A' + 10; }
template <size_t N> constexpr std::array<uint8_t, N> from_hex(const char* s) {
    std::array<uint8_t, N> r{};
    for (size_t i = 0; i < N; i++) r[i] = (uint8_t)(hexval(s[2 * i]) * 16 + hexval(s[2 * i + 1]));
    return r;
}
template <size_t N> inline int cmp_be(const std::array<uint8_t, N>& a, const std::array<uint8_t, N>& b) {
    int c = std::memcmp(a.data(), b.data(), N);
    return c < 0 ? -1 : c > 0 ? 1 : 0;
}
template <size_t N> inline bool is_zero(const std::array<uint8_t, N>& a) {
    uint8_t t = 0;
    for (uint8_t v : a) t |= v;
    return t == 0;
}
template <size_t N> inline std::array<uint8_t, N> sub_be(const std::array<uint8_t, N>& a, const std::array<uint8_t, N>& b) {
    std::array<uint8_t, N> r{};
    int borrow = 0;
    for (size_t i = N; i-- > 0;) {
        int d = (int)a[i] - (int)b[i] - borrow;
        borrow = d < 0;
        r[i] = (uint8_t)(d + (borrow << 8));
    }
    return r;
}
template <size_t N> inline std::array<uint8_t, N> shr1_be(const std::array<uint8_t, N>& a) {
    std::array<uint8_t, N> r{};
    uint8_t carry = 0;
    for (size_t i = 0; i < N; i++) {
        r[i] = (uint8_t)((a[i] >> 1) | (carry << 7));
        carry = a[i] & 1;
    }
    return r;
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------
// curve markers: `elliptic_curve::Curve` (FieldBytesSize, ORDER) + the SEC1 default of `PointCompression`

struct Secp256k1 {   // k256/src/lib.rs:76-111 (ORDER :99-100, compress-by-default :108-111); low-s rule k256/src/ecdsa.rs:203-205
    static constexpr int ID = ECB200_K256;
    static constexpr size_t FB = 32;
    static constexpr bool COMPRESS_POINTS = true;
    static constexpr bool LOW_S_ONLY = true;
    static constexpr const char* NAME = "k256";
    static std::array<uint8_t, 32> order() { return detail::from_hex<32>("FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141"); }
};
struct NistP256 {    // p256/src/lib.rs:74-120 (ORDER :91-92)
    static constexpr int ID = ECB200_P256;
    static constexpr size_t FB = 32;
    static constexpr bool COMPRESS_POINTS = false;
    static constexpr bool LOW_S_ONLY = false;
    static constexpr const char* NAME = "p256";
    static std::array<uint8_t, 32> order() { return detail::from_hex<32>("FFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551"); }
};
struct NistP384 {    // p384/src/lib.rs:50-76
    static constexpr int ID = ECB200_P384;
    static constexpr size_t FB = 48;
    static constexpr bool COMPRESS_POINTS = false;
    static constexpr bool LOW_S_ONLY = false;
    static constexpr const char* NAME = "p384";
    static std::array<uint8_t, 48> order() {
        return detail::from_hex<48>("FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973");
    }
};
struct Sm2 {         // sm2/src/lib.rs:60-85
    static constexpr int ID = ECB200_SM2;
    static constexpr size_t FB = 32;
    static constexpr bool COMPRESS_POINTS = false;
    static constexpr bool LOW_S_ONLY = false;
    static constexpr const char* NAME = "sm2";
    static std::array<uint8_t, 32> order() { return detail::from_hex<32>("FFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123"); }
};

struct NistP192 {    // p192/src/lib.rs:42-66 (the primeorder template on 24-byte fields)
    static constexpr int ID = ECB200_P192;
    static constexpr size_t FB = 24;
    static constexpr bool COMPRESS_POINTS = false;
    static constexpr bool LOW_S_ONLY = false;
    static constexpr const char* NAME = "p192";
    static std::array<uint8_t, 24> order() { return detail::from_hex<24>("FFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831"); }
};

struct NistP224 {    // p224/src/lib.rs (28-byte fields)
    static constexpr int ID = ECB200_P224;
    static constexpr size_t FB = 28;
    static constexpr bool COMPRESS_POINTS = false;
    static constexpr bool LOW_S_ONLY = false;
    static constexpr const char* NAME = "p224";
    static std::array<uint8_t, 28> order() { return detail::from_hex<28>("FFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D"); }
};

template <class C> using FieldBytes = std::array<uint8_t, C::FB>;   // big-endian, `elliptic_curve::FieldBytes<C>`

// ---------------------------------------------------------------------------------------------------------------
// the engine: one context = one GPU (one process per GPU); thread-compatible, not thread-safe

class Engine {
public:
    explicit Engine(int device = 0) {
        int rc = ecb200_init(device, &ctx_);
        if (rc != 0 || !ctx_)
            throw Error(rc, "ecb200_init(device=" + std::to_string(device) + ") failed with " + std::to_string(rc) +
                                ": no usable CUDA device (there is no CPU fallback)");
    }
    // one context over several GPUs of the box (ecb200_init_multi): every batch entry point shards its rows by contiguous
    // index range over `devices` inside one call; an empty list means every visible device
    explicit Engine(const std::vector<int>& devices) {
        int rc = ecb200_init_multi((int)devices.size(), devices.empty() ? nullptr : devices.data(), &ctx_);
        if (rc != 0 || !ctx_)
            throw Error(rc, "ecb200_init_multi over " + std::to_string(devices.size()) + " devices failed with " + std::to_string(rc) +
                                ": no usable CUDA device (there is no CPU fallback)");
    }
    int device_count() const { return ecb200_device_count(ctx_); }
    ~Engine() { if (ctx_) ecb200_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    Engine(Engine&& o) noexcept : ctx_(o.ctx_) { o.ctx_ = nullptr; }
    Engine& operator=(Engine&& o) noexcept {
        if (this != &o) { if (ctx_) ecb200_destroy(ctx_); ctx_ = o.ctx_; o.ctx_ = nullptr; }
        return *this;
    }
    ecb200_ctx* raw() const { return ctx_; }
    uint64_t launch_count() const { return ecb200_launch_count(ctx_); }
    void sync() { check(ecb200_sync(ctx_), "sync"); }
    void check(int rc, const char* what) const {
        if (rc != 0) throw Error(rc, std::string(what) + " failed (" + std::to_string(rc) + "): " + ecb200_last_error(ctx_));
    }
private:
    ecb200_ctx* ctx_ = nullptr;
};

// Contiguous index shard [lo, hi) of rank `rank` out of `world`: floor(i*n/g) boundaries.
inline std::pair<size_t, size_t> shard_range(size_t n, size_t rank, size_t world) {
    return {(size_t)((unsigned __int128)rank * n / world), (size_t)((unsigned __int128)(rank + 1) * n / world)};
}

// ---------------------------------------------------------------------------------------------------------------
// Scalar

template <class C> class Scalar {
public:
    Scalar() = default;   // ZERO
    static Scalar zero() { return Scalar(); }
    static Scalar one() { return from_u64(1); }
    static Scalar from_u64(uint64_t v) {
        Scalar s;
        for (int i = 0; i < 8; i++) s.repr_[C::FB - 1 - i] = (uint8_t)(v >> (8 * i));
        return s;
    }
    // PrimeField::from_repr: None when the integer is >= n
    static std::optional<Scalar> from_repr(const FieldBytes<C>& b) {
        if (detail::cmp_be(b, C::order()) >= 0) return std::nullopt;
        Scalar s;
        s.repr_ = b;
        return s;
    }
    // Reduce<Uint>::reduce_bytes: one conditional subtraction of n (k256 scalar.rs:700-713)
    static Scalar reduce_bytes(const FieldBytes<C>& b) {
        Scalar s;
        s.repr_ = detail::cmp_be(b, C::order()) >= 0 ? detail::sub_be(b, C::order()) : b;
        return s;
    }
    const FieldBytes<C>& to_repr() const { return repr_; }
    const FieldBytes<C>& to_bytes() const { return repr_; }
    bool is_zero() const { return detail::is_zero(repr_); }
    // IsHigh::is_high: s > n / 2 (k256 scalar.rs:519-523)
    bool is_high() const { return detail::cmp_be(repr_, detail::shr1_be(C::order())) > 0; }
    // -s mod n (used by Signature::normalize_s)
    Scalar negate() const {
        if (is_zero()) return *this;
        Scalar s;
        s.repr_ = detail::sub_be(C::order(), repr_);
        return s;
    }
    bool operator==(const Scalar& o) const { return repr_ == o.repr_; }
    bool operator!=(const Scalar& o) const { return !(*this == o); }
private:
    FieldBytes<C> repr_{};
};

// ---------------------------------------------------------------------------------------------------------------
// SEC1 EncodedPoint (sec1 crate; tags: 00 identity, 02/03 compressed, 04 uncompressed, 05 compact)

template <class C> class EncodedPoint {
public:
    static constexpr size_t MAX = 1 + 2 * C::FB;
    EncodedPoint() { bytes_.fill(0); len_ = 1; }   // identity
    static EncodedPoint identity() { return EncodedPoint(); }
    // EncodedPoint::from_bytes: length must match the tag
    static std::optional<EncodedPoint> from_bytes(const uint8_t* p, size_t n) {
        if (n == 0 || n > MAX) return std::nullopt;
        size_t want = p[0] == 0 ? 1 : (p[0] == 2 || p[0] == 3 || p[0] == 5) ? 1 + C::FB : p[0] == 4 ? 1 + 2 * C::FB : 0;
        if (want != n) return std::nullopt;
        EncodedPoint e;
        std::memcpy(e.bytes_.data(), p, n);
        e.len_ = n;
        return e;
    }
    static std::optional<EncodedPoint> from_bytes(const std::vector<uint8_t>& v) { return from_bytes(v.data(), v.size()); }
    // one fixed-size output slot of the C ABI (identity = all-zero slot)
    static EncodedPoint from_slot(const uint8_t* p, size_t slot) {
        EncodedPoint e;
        if (p[0] != 0) { std::memcpy(e.bytes_.data(), p, slot); e.len_ = slot; }
        return e;
    }
    uint8_t tag() const { return bytes_[0]; }
    bool is_identity() const { return bytes_[0] == 0; }
    bool is_compressed() const { return bytes_[0] == 2 || bytes_[0] == 3; }
    const uint8_t* as_bytes() const { return bytes_.data(); }
    size_t len() const { return len_; }
    std::vector<uint8_t> to_vec() const { return std::vector<uint8_t>(bytes_.begin(), bytes_.begin() + len_); }
    bool operator==(const EncodedPoint& o) const { return len_ == o.len_ && std::memcmp(bytes_.data(), o.bytes_.data(), len_) == 0; }
    bool operator!=(const EncodedPoint& o) const { return !(*this == o); }
    const std::array<uint8_t, MAX>& padded() const { return bytes_; }   // zero-padded to MAX (decode stride)
private:
    std::array<uint8_t, MAX> bytes_;
    size_t len_;
};

// ---------------------------------------------------------------------------------------------------------------
// AffinePoint

template <class C> struct AffinePoint {
    FieldBytes<C> x{};
    FieldBytes<C> y{};
    bool infinity = true;   // AffinePoint::IDENTITY = {0, 0, infinity: 1} (k256 affine.rs:63-68)

    static AffinePoint identity() { return AffinePoint(); }
    static AffinePoint from_coordinates_unchecked(const FieldBytes<C>& x, const FieldBytes<C>& y) {
        AffinePoint p;
        p.x = x; p.y = y; p.infinity = false;
        return p;
    }
    bool is_identity() const { return infinity; }
    bool y_is_odd() const { return (y[C::FB - 1] & 1) != 0; }   // AffineCoordinates::y_is_odd
    // ToEncodedPoint::to_encoded_point (k256 affine.rs:272-284; primeorder affine.rs:340-358)
    EncodedPoint<C> to_encoded_point(bool compress) const {
        if (infinity) return EncodedPoint<C>::identity();
        uint8_t buf[1 + 2 * C::FB];
        buf[0] = compress ? (y_is_odd() ? 3 : 2) : 4;
        std::memcpy(buf + 1, x.data(), C::FB);
        if (!compress) std::memcpy(buf + 1 + C::FB, y.data(), C::FB);
        return *EncodedPoint<C>::from_bytes(buf, compress ? 1 + C::FB : 1 + 2 * C::FB);
    }
    EncodedPoint<C> to_encoded_point() const { return to_encoded_point(C::COMPRESS_POINTS); }
    bool operator==(const AffinePoint& o) const {
        return infinity == o.infinity && (infinity || (x == o.x && y == o.y));
    }
    bool operator!=(const AffinePoint& o) const { return !(*this == o); }

    // FromEncodedPoint::from_encoded_point over a slice: decompression / validation runs on the device
    // (ecb200_decode_points).  None where the reference's CtOption is none (x >= p, no square root, off-curve,
    // unknown tag); the identity encoding decodes to IDENTITY.
    static std::vector<std::optional<AffinePoint>> from_encoded_points(Engine& eng, const std::vector<EncodedPoint<C>>& enc) {
        const size_t n = enc.size(), stride = EncodedPoint<C>::MAX;
        std::vector<uint8_t> in(n * stride), xy(n * 2 * C::FB), st(n);
        for (size_t i = 0; i < n; i++) std::memcpy(&in[i * stride], enc[i].padded().data(), stride);
        eng.check(ecb200_decode_points(eng.raw(), C::ID, n, in.data(), stride, ECB200_DECODE_SEC1, xy.data(), st.data()), "decode_points");
        return unpack(xy, st);
    }
    // DecompactPoint::decompact over a slice of x coordinates (k256: even root; primeorder: smaller y)
    static std::vector<std::optional<AffinePoint>> decompact_batch(Engine& eng, const std::vector<FieldBytes<C>>& xs) {
        const size_t n = xs.size();
        std::vector<uint8_t> xy(n * 2 * C::FB), st(n);
        eng.check(ecb200_decode_points(eng.raw(), C::ID, n, reinterpret_cast<const uint8_t*>(xs.data()), C::FB, ECB200_DECODE_COMPACT,
                                       xy.data(), st.data()), "decode_points");
        return unpack(xy, st);
    }
private:
    static std::vector<std::optional<AffinePoint>> unpack(const std::vector<uint8_t>& xy, const std::vector<uint8_t>& st) {
        std::vector<std::optional<AffinePoint>> out(st.size());
        for (size_t i = 0; i < st.size(); i++) {
            if (st[i] == 2) out[i] = AffinePoint::identity();
            else if (st[i] == 1) {
                AffinePoint p;
                std::memcpy(p.x.data(), &xy[i * 2 * C::FB], C::FB);
                std::memcpy(p.y.data(), &xy[i * 2 * C::FB + C::FB], C::FB);
                p.infinity = false;
                out[i] = p;
            }
        }
        return out;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// ProjectivePoint (homogeneous X : Y : Z, x = X/Z — the coordinates of both k256 and primeorder)

template <class C> struct ProjectivePoint {
    FieldBytes<C> x{};
    FieldBytes<C> y{};
    FieldBytes<C> z{};

    // ProjectivePoint::IDENTITY = (0 : 1 : 0) (k256 projective.rs:44-50)
    static ProjectivePoint identity() {
        ProjectivePoint p;
        p.y[C::FB - 1] = 1;
        return p;
    }
    // From<AffinePoint> (k256 projective.rs:381-390)
    static ProjectivePoint from_affine(const AffinePoint<C>& a) {
        if (a.infinity) return identity();
        ProjectivePoint p;
        p.x = a.x; p.y = a.y; p.z[C::FB - 1] = 1;
        return p;
    }
    bool is_identity() const { return detail::is_zero(z); }   // Group::is_identity: Z == 0

    // MulByGenerator::mul_by_generator over a slice of (secret) scalars -> affine points.  Constant-time path.
    static std::vector<AffinePoint<C>> mul_by_generator_batch(Engine& eng, const std::vector<Scalar<C>>& ks) {
        const size_t n = ks.size(), slot = 1 + 2 * C::FB;
        std::vector<uint8_t> out(n * slot);
        eng.check(ecb200_mul_gen(eng.raw(), C::ID, n, scalar_bytes(ks), out.data(), ECB200_FLAG_CT | ECB200_FLAG_UNCOMPRESSED), "mul_gen");
        return slots_to_affine(out, n);
    }
    // same, SEC1-encoded with the curve's default compression (`to_encoded_point(C::COMPRESS_POINTS)`)
    static std::vector<EncodedPoint<C>> mul_by_generator_batch_encoded(Engine& eng, const std::vector<Scalar<C>>& ks) {
        const size_t n = ks.size(), slot = ecb200_point_slot_bytes(C::ID, 0);
        std::vector<uint8_t> out(n * slot);
        eng.check(ecb200_mul_gen(eng.raw(), C::ID, n, scalar_bytes(ks), out.data(), ECB200_FLAG_CT), "mul_gen");
        std::vector<EncodedPoint<C>> r(n);
        for (size_t i = 0; i < n; i++) r[i] = EncodedPoint<C>::from_slot(&out[i * slot], slot);
        return r;
    }
    // [&P_i * &k_i] then batch_normalize.  `Mul` in the reference is constant time, so this is the CT path.
    static std::vector<AffinePoint<C>> mul_batch(Engine& eng, const std::vector<std::pair<ProjectivePoint, Scalar<C>>>& terms) {
        return mul_impl(eng, terms, ECB200_FLAG_CT);
    }
    // public scalars and points only (variable-time windows)
    static std::vector<AffinePoint<C>> mul_batch_vartime(Engine& eng, const std::vector<std::pair<ProjectivePoint, Scalar<C>>>& terms) {
        return mul_impl(eng, terms, 0);
    }
    // BatchNormalize<[ProjectivePoint]>::batch_normalize -> Vec<AffinePoint>
    static std::vector<AffinePoint<C>> batch_normalize(Engine& eng, const std::vector<ProjectivePoint>& pts) {
        static_assert(sizeof(ProjectivePoint) == 3 * C::FB, "ProjectivePoint must be three packed FieldBytes");
        const size_t n = pts.size();
        std::vector<uint8_t> xy(n * 2 * C::FB), inf(n);
        eng.check(ecb200_batch_normalize(eng.raw(), C::ID, n, reinterpret_cast<const uint8_t*>(pts.data()), xy.data(), inf.data()),
                  "batch_normalize");
        std::vector<AffinePoint<C>> r(n);
        for (size_t i = 0; i < n; i++) {
            if (inf[i]) continue;
            std::memcpy(r[i].x.data(), &xy[i * 2 * C::FB], C::FB);
            std::memcpy(r[i].y.data(), &xy[i * 2 * C::FB + C::FB], C::FB);
            r[i].infinity = false;
        }
        return r;
    }
    // LinearCombinationExt::lincomb_ext(&[(P, k)]) -> one ProjectivePoint (constant time like the reference)
    static ProjectivePoint lincomb_ext(Engine& eng, const std::vector<std::pair<ProjectivePoint, Scalar<C>>>& terms) {
        std::vector<uint8_t> pts, ks;
        pack(terms, pts, ks);
        ProjectivePoint r;
        static_assert(sizeof(ProjectivePoint) == 3 * C::FB, "ProjectivePoint must be three packed FieldBytes");
        eng.check(ecb200_lincomb(eng.raw(), C::ID, terms.size(), pts.data(), ks.data(), reinterpret_cast<uint8_t*>(&r),
                                 ECB200_FLAG_CT | ECB200_FLAG_PROJ, ECB200_FLAG_PROJ), "lincomb");
        return r;
    }
    // LinearCombination::lincomb(x, k, y, l) = x*k + y*l
    static ProjectivePoint lincomb(Engine& eng, const ProjectivePoint& x, const Scalar<C>& k, const ProjectivePoint& y, const Scalar<C>& l) {
        return lincomb_ext(eng, {{x, k}, {y, l}});
    }
    // LinearCombination::lincomb(&x, &k, &y, &l) over slices: out[i] = x_i*k_i + y_i*l_i, one point per row (constant time
    // like the reference; k256/src/arithmetic/mul.rs:313-323, primeorder/src/projective.rs:415-420)
    struct LincombTerm { ProjectivePoint x; Scalar<C> k; ProjectivePoint y; Scalar<C> l; };
    static std::vector<AffinePoint<C>> lincomb_batch(Engine& eng, const std::vector<LincombTerm>& terms, bool vartime = false) {
        const size_t n = terms.size(), slot = 1 + 2 * C::FB;
        std::vector<uint8_t> p1(n * 3 * C::FB + 1), p2(n * 3 * C::FB + 1), k1(n * C::FB + 1), k2(n * C::FB + 1), out(n * slot + 1);
        for (size_t i = 0; i < n; i++) {
            std::memcpy(&p1[i * 3 * C::FB], &terms[i].x, 3 * C::FB);
            std::memcpy(&p2[i * 3 * C::FB], &terms[i].y, 3 * C::FB);
            std::memcpy(&k1[i * C::FB], terms[i].k.to_repr().data(), C::FB);
            std::memcpy(&k2[i * C::FB], terms[i].l.to_repr().data(), C::FB);
        }
        eng.check(ecb200_lincomb2(eng.raw(), C::ID, n, p1.data(), k1.data(), p2.data(), k2.data(), out.data(), nullptr,
                                  (vartime ? 0u : ECB200_FLAG_CT) | ECB200_FLAG_PROJ | ECB200_FLAG_UNCOMPRESSED), "lincomb2");
        return slots_to_affine(out, n);
    }
    // group::Curve::to_affine for one point (a batch of one)
    AffinePoint<C> to_affine(Engine& eng) const { return batch_normalize(eng, {*this})[0]; }

private:
    static const uint8_t* scalar_bytes(const std::vector<Scalar<C>>& ks) {
        static_assert(sizeof(Scalar<C>) == C::FB, "Scalar must be one packed FieldBytes");
        return reinterpret_cast<const uint8_t*>(ks.data());
    }
    static void pack(const std::vector<std::pair<ProjectivePoint, Scalar<C>>>& terms, std::vector<uint8_t>& pts, std::vector<uint8_t>& ks) {
        const size_t n = terms.size();
        pts.resize(n * 3 * C::FB + 1);
        ks.resize(n * C::FB + 1);
        for (size_t i = 0; i < n; i++) {
            std::memcpy(&pts[i * 3 * C::FB], &terms[i].first, 3 * C::FB);
            std::memcpy(&ks[i * C::FB], terms[i].second.to_repr().data(), C::FB);
        }
    }
    static std::vector<AffinePoint<C>> slots_to_affine(const std::vector<uint8_t>& out, size_t n) {
        const size_t slot = 1 + 2 * C::FB;
        std::vector<AffinePoint<C>> r(n);
        for (size_t i = 0; i < n; i++) {
            if (out[i * slot] != 4) continue;   // identity slot
            std::memcpy(r[i].x.data(), &out[i * slot + 1], C::FB);
            std::memcpy(r[i].y.data(), &out[i * slot + 1 + C::FB], C::FB);
            r[i].infinity = false;
        }
        return r;
    }
    static std::vector<AffinePoint<C>> mul_impl(Engine& eng, const std::vector<std::pair<ProjectivePoint, Scalar<C>>>& terms, uint32_t ct) {
        const size_t n = terms.size(), slot = 1 + 2 * C::FB;
        std::vector<uint8_t> pts, ks, out(n * slot + 1);
        pack(terms, pts, ks);
        eng.check(ecb200_mul_var(eng.raw(), C::ID, n, pts.data(), nullptr, ks.data(), out.data(), nullptr,
                                 ct | ECB200_FLAG_PROJ | ECB200_FLAG_UNCOMPRESSED), "mul_var");
        return slots_to_affine(out, n);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// ECDSA (the `ecdsa` crate surface the curve crates re-export: k256/src/ecdsa.rs:162-174, p256/src/ecdsa.rs:52-64)

namespace ecdsa {

// hazmat::bits2field (ecdsa 0.16.9): shorter than FB/2 -> Err; shorter than FB -> left-padded; longer -> leftmost FB bytes
template <class C> inline std::optional<FieldBytes<C>> bits2field(const uint8_t* prehash, size_t len) {
    if (len < C::FB / 2) return std::nullopt;
    FieldBytes<C> z{};
    if (len < C::FB) std::memcpy(z.data() + (C::FB - len), prehash, len);
    else std::memcpy(z.data(), prehash, C::FB);
    return z;
}

// RecoveryId::to_byte: bit 0 = y(R) odd, bit 1 = x(R) reduced
struct RecoveryId {
    uint8_t byte = 0;
    static std::optional<RecoveryId> from_byte(uint8_t b) { return b < 4 ? std::optional<RecoveryId>(RecoveryId{b}) : std::nullopt; }
    bool is_y_odd() const { return byte & 1; }
    bool is_x_reduced() const { return (byte & 2) != 0; }
    uint8_t to_byte() const { return byte; }
    bool operator==(const RecoveryId& o) const { return byte == o.byte; }
};

template <class C> class Signature {
public:
    // Signature::from_scalars: r and s must be in [1, n-1]
    static std::optional<Signature> from_scalars(const FieldBytes<C>& r, const FieldBytes<C>& s) {
        auto sr = Scalar<C>::from_repr(r), ss = Scalar<C>::from_repr(s);
        if (!sr || !ss || sr->is_zero() || ss->is_zero()) return std::nullopt;
        Signature g;
        g.r_ = *sr; g.s_ = *ss;
        return g;
    }
    // Signature::from_slice / TryFrom<&[u8]>: exactly 2*FB bytes r || s
    static std::optional<Signature> from_slice(const uint8_t* p, size_t len) {
        if (len != 2 * C::FB) return std::nullopt;
        FieldBytes<C> r, s;
        std::memcpy(r.data(), p, C::FB);
        std::memcpy(s.data(), p + C::FB, C::FB);
        return from_scalars(r, s);
    }
    const Scalar<C>& r() const { return r_; }
    const Scalar<C>& s() const { return s_; }
    std::array<uint8_t, 2 * C::FB> to_bytes() const {
        std::array<uint8_t, 2 * C::FB> b;
        std::memcpy(b.data(), r_.to_repr().data(), C::FB);
        std::memcpy(b.data() + C::FB, s_.to_repr().data(), C::FB);
        return b;
    }
    // Signature::normalize_s: Some(low-s twin) when s is high, None otherwise
    std::optional<Signature> normalize_s() const {
        if (!s_.is_high()) return std::nullopt;
        Signature g = *this;
        g.s_ = s_.negate();
        return g;
    }
    bool operator==(const Signature& o) const { return r_ == o.r_ && s_ == o.s_; }
private:
    Scalar<C> r_, s_;
};

template <class C> class VerifyingKey {
public:
    // VerifyingKey::from_affine: the identity is not a valid public key.  (Off-curve coordinates cannot be
    // represented by the reference's AffinePoint; here they are caught on the device: verification gives Err.)
    static std::optional<VerifyingKey> from_affine(const AffinePoint<C>& a) {
        if (a.infinity) return std::nullopt;
        VerifyingKey k;
        k.point_ = a;
        return k;
    }
    // VerifyingKey::from_sec1_bytes over a slice: decoded / decompressed on the device
    static std::vector<std::optional<VerifyingKey>> from_sec1_bytes_batch(Engine& eng, const std::vector<std::vector<uint8_t>>& keys) {
        std::vector<EncodedPoint<C>> enc;
        std::vector<size_t> idx;
        for (size_t i = 0; i < keys.size(); i++) {
            auto e = EncodedPoint<C>::from_bytes(keys[i]);
            if (e) { enc.push_back(*e); idx.push_back(i); }
        }
        auto pts = AffinePoint<C>::from_encoded_points(eng, enc);
        std::vector<std::optional<VerifyingKey>> out(keys.size());
        for (size_t j = 0; j < idx.size(); j++)
            if (pts[j]) out[idx[j]] = from_affine(*pts[j]);
        return out;
    }
    const AffinePoint<C>& as_affine() const { return point_; }
    EncodedPoint<C> to_encoded_point(bool compress) const { return point_.to_encoded_point(compress); }
    bool operator==(const VerifyingKey& o) const { return point_ == o.point_; }

    // PrehashVerifier::verify_prehash over slices: Vec<Result<(), Error>>.  keys[i] verifies (prehashes[i], sigs[i]).
    static std::vector<Result> verify_prehash_batch(Engine& eng, const std::vector<VerifyingKey>& keys,
                                                    const std::vector<std::vector<uint8_t>>& prehashes,
                                                    const std::vector<Signature<C>>& sigs) {
        const size_t n = keys.size();
        if (prehashes.size() != n || sigs.size() != n) throw Error(ECB200_ERR_ARG, "verify_prehash_batch: slice lengths differ");
        std::vector<uint8_t> q, z, rs;
        std::vector<size_t> idx;
        q.reserve(n * 2 * C::FB); z.reserve(n * C::FB); rs.reserve(n * 2 * C::FB);
        for (size_t i = 0; i < n; i++) {
            auto zi = bits2field<C>(prehashes[i].data(), prehashes[i].size());
            if (!zi) continue;   // Err: prehash too short
            idx.push_back(i);
            q.insert(q.end(), keys[i].point_.x.begin(), keys[i].point_.x.end());
            q.insert(q.end(), keys[i].point_.y.begin(), keys[i].point_.y.end());
            z.insert(z.end(), zi->begin(), zi->end());
            auto b = sigs[i].to_bytes();
            rs.insert(rs.end(), b.begin(), b.end());
        }
        std::vector<uint8_t> ok(idx.size() + 1);
        q.push_back(0); z.push_back(0); rs.push_back(0);   // non-null data() for empty batches
        eng.check(ecb200_ecdsa_verify(eng.raw(), C::ID, idx.size(), q.data(), z.data(), rs.data(), ok.data()), "ecdsa_verify");
        std::vector<Result> out(n, Result::Err());
        for (size_t j = 0; j < idx.size(); j++)
            if (ok[j]) out[idx[j]] = Result::Ok();
        return out;
    }
    // VerifyingKey::recover_from_prehash over slices (ecdsa 0.16.9 recovery.rs; k256/src/ecdsa.rs:113-140,278-343)
    static std::vector<std::optional<VerifyingKey>> recover_from_prehash_batch(Engine& eng, const std::vector<std::vector<uint8_t>>& prehashes,
                                                                               const std::vector<Signature<C>>& sigs,
                                                                               const std::vector<RecoveryId>& recids) {
        const size_t n = sigs.size(), slot = 1 + 2 * C::FB;
        if (prehashes.size() != n || recids.size() != n) throw Error(ECB200_ERR_ARG, "recover_from_prehash_batch: slice lengths differ");
        std::vector<uint8_t> z, rs, id;
        std::vector<size_t> idx;
        for (size_t i = 0; i < n; i++) {
            auto zi = bits2field<C>(prehashes[i].data(), prehashes[i].size());
            if (!zi) continue;
            idx.push_back(i);
            z.insert(z.end(), zi->begin(), zi->end());
            auto b = sigs[i].to_bytes();
            rs.insert(rs.end(), b.begin(), b.end());
            id.push_back(recids[i].to_byte());
        }
        std::vector<uint8_t> keys(idx.size() * slot + 1), ok(idx.size() + 1);
        z.push_back(0); rs.push_back(0); id.push_back(0);
        eng.check(ecb200_ecdsa_recover(eng.raw(), C::ID, idx.size(), z.data(), rs.data(), id.data(), keys.data(), ok.data(),
                                       ECB200_FLAG_UNCOMPRESSED), "ecdsa_recover");
        std::vector<std::optional<VerifyingKey>> out(n);
        for (size_t j = 0; j < idx.size(); j++) {
            if (!ok[j]) continue;
            AffinePoint<C> p;
            std::memcpy(p.x.data(), &keys[j * slot + 1], C::FB);
            std::memcpy(p.y.data(), &keys[j * slot + 1 + C::FB], C::FB);
            p.infinity = false;
            out[idx[j]] = from_affine(p);
        }
        return out;
    }
private:
    AffinePoint<C> point_;
};

// SignPrimitive::try_sign_prehashed over slices: d_i signs z_i with the caller's nonce k_i (RFC 6979 derivation is
// HMAC work and stays with the caller).  None where the reference returns Err (k = 0, r = 0 or s = 0).  Constant time.
template <class C>
inline std::vector<std::optional<std::pair<Signature<C>, RecoveryId>>> try_sign_prehashed_batch(
    Engine& eng, const std::vector<Scalar<C>>& d, const std::vector<Scalar<C>>& k, const std::vector<FieldBytes<C>>& z) {
    const size_t n = d.size();
    if (k.size() != n || z.size() != n) throw Error(ECB200_ERR_ARG, "try_sign_prehashed_batch: slice lengths differ");
    static_assert(sizeof(Scalar<C>) == C::FB, "Scalar must be one packed FieldBytes");
    std::vector<uint8_t> rs(n * 2 * C::FB + 1), rid(n + 1), ok(n + 1);
    eng.check(ecb200_ecdsa_sign(eng.raw(), C::ID, n, reinterpret_cast<const uint8_t*>(d.data()), reinterpret_cast<const uint8_t*>(k.data()),
                                reinterpret_cast<const uint8_t*>(z.data()), rs.data(), rid.data(), ok.data()), "ecdsa_sign");
    std::vector<std::optional<std::pair<Signature<C>, RecoveryId>>> out(n);
    for (size_t i = 0; i < n; i++) {
        if (!ok[i]) continue;
        auto sig = Signature<C>::from_slice(&rs[i * 2 * C::FB], 2 * C::FB);
        if (sig) out[i] = std::make_pair(*sig, RecoveryId{rid[i]});
    }
    return out;
}

}  // namespace ecdsa

// ---------------------------------------------------------------------------------------------------------------
// BIP340 Schnorr over secp256k1 (k256/src/schnorr/verifying.rs:35-89).  The tagged challenge hash is SHA-256 work and
// stays with the caller, like the ECDSA prehash: e_i = tagged_hash("BIP0340/challenge", r || pk || msg).

namespace schnorr {
struct VerifyingKey {
    std::array<uint8_t, 32> x{};   // x-only public key (VerifyingKey::to_bytes)
    // verify_raw over slices: Ok iff lift_x(pk) exists, 0 < r < p, 1 <= s < n and s*G - e*P is finite with even y and x = r
    static std::vector<Result> verify_raw_batch(Engine& eng, const std::vector<VerifyingKey>& keys, const std::vector<std::array<uint8_t, 32>>& e,
                                                const std::vector<std::array<uint8_t, 64>>& sigs) {
        const size_t n = keys.size();
        if (e.size() != n || sigs.size() != n) throw Error(ECB200_ERR_ARG, "schnorr verify: slice lengths differ");
        static_assert(sizeof(VerifyingKey) == 32, "x-only key must be 32 packed bytes");
        std::vector<uint8_t> ok(n + 1);
        eng.check(ecb200_schnorr_verify(eng.raw(), n, reinterpret_cast<const uint8_t*>(keys.data()), reinterpret_cast<const uint8_t*>(e.data()),
                                        reinterpret_cast<const uint8_t*>(sigs.data()), ok.data()), "schnorr_verify");
        std::vector<Result> out(n, Result::Err());
        for (size_t i = 0; i < n; i++) if (ok[i]) out[i] = Result::Ok();
        return out;
    }
};
}  // namespace schnorr

// SM2DSA (sm2/src/dsa/verifying.rs:130-168): e_i = SM3(Z_A || M) computed by the caller
namespace sm2dsa {
inline std::vector<Result> verify_prehash_batch(Engine& eng, const std::vector<AffinePoint<Sm2>>& keys, const std::vector<FieldBytes<Sm2>>& e,
                                                const std::vector<ecdsa::Signature<Sm2>>& sigs) {
    const size_t n = keys.size();
    if (e.size() != n || sigs.size() != n) throw Error(ECB200_ERR_ARG, "sm2dsa verify: slice lengths differ");
    std::vector<uint8_t> q(n * 64 + 1), rs(n * 64 + 1), ok(n + 1);
    for (size_t i = 0; i < n; i++) {
        std::memcpy(&q[i * 64], keys[i].x.data(), 32);
        std::memcpy(&q[i * 64 + 32], keys[i].y.data(), 32);
        auto b = sigs[i].to_bytes();
        std::memcpy(&rs[i * 64], b.data(), 64);
    }
    eng.check(ecb200_sm2dsa_verify(eng.raw(), n, q.data(), reinterpret_cast<const uint8_t*>(e.data()), rs.data(), ok.data()), "sm2dsa_verify");
    std::vector<Result> out(n, Result::Err());
    for (size_t i = 0; i < n; i++) if (ok[i]) out[i] = Result::Ok();
    return out;
}
}  // namespace sm2dsa

// ---------------------------------------------------------------------------------------------------------------
// per-crate aliases, named as in the reference

namespace k256 {
using Curve = Secp256k1;
using FieldBytes = ecb200::FieldBytes<Secp256k1>;
using Scalar = ecb200::Scalar<Secp256k1>;
using AffinePoint = ecb200::AffinePoint<Secp256k1>;
using ProjectivePoint = ecb200::ProjectivePoint<Secp256k1>;
using EncodedPoint = ecb200::EncodedPoint<Secp256k1>;
namespace ecdsa {
using Signature = ecb200::ecdsa::Signature<Secp256k1>;
using VerifyingKey = ecb200::ecdsa::VerifyingKey<Secp256k1>;
using RecoveryId = ecb200::ecdsa::RecoveryId;
}  // namespace ecdsa
namespace schnorr { using VerifyingKey = ecb200::schnorr::VerifyingKey; }
}  // namespace k256
namespace p256 {
using Curve = NistP256;
using FieldBytes = ecb200::FieldBytes<NistP256>;
using Scalar = ecb200::Scalar<NistP256>;
using AffinePoint = ecb200::AffinePoint<NistP256>;
using ProjectivePoint = ecb200::ProjectivePoint<NistP256>;
using EncodedPoint = ecb200::EncodedPoint<NistP256>;
namespace ecdsa {
using Signature = ecb200::ecdsa::Signature<NistP256>;
using VerifyingKey = ecb200::ecdsa::VerifyingKey<NistP256>;
}  // namespace ecdsa
}  // namespace p256
namespace p384 {
using Curve = NistP384;
using FieldBytes = ecb200::FieldBytes<NistP384>;
using Scalar = ecb200::Scalar<NistP384>;
using AffinePoint = ecb200::AffinePoint<NistP384>;
using ProjectivePoint = ecb200::ProjectivePoint<NistP384>;
using EncodedPoint = ecb200::EncodedPoint<NistP384>;
namespace ecdsa {
using Signature = ecb200::ecdsa::Signature<NistP384>;
using VerifyingKey = ecb200::ecdsa::VerifyingKey<NistP384>;
}  // namespace ecdsa
}  // namespace p384
namespace p192 {
using Curve = NistP192;
using FieldBytes = ecb200::FieldBytes<NistP192>;
using Scalar = ecb200::Scalar<NistP192>;
using AffinePoint = ecb200::AffinePoint<NistP192>;
using ProjectivePoint = ecb200::ProjectivePoint<NistP192>;
using EncodedPoint = ecb200::EncodedPoint<NistP192>;
namespace ecdsa {
using Signature = ecb200::ecdsa::Signature<NistP192>;
using VerifyingKey = ecb200::ecdsa::VerifyingKey<NistP192>;
}  // namespace ecdsa
}  // namespace p192
namespace p224 {
using Curve = NistP224;
using FieldBytes = ecb200::FieldBytes<NistP224>;
using Scalar = ecb200::Scalar<NistP224>;
using AffinePoint = ecb200::AffinePoint<NistP224>;
using ProjectivePoint = ecb200::ProjectivePoint<NistP224>;
using EncodedPoint = ecb200::EncodedPoint<NistP224>;
namespace ecdsa {
using Signature = ecb200::ecdsa::Signature<NistP224>;
using VerifyingKey = ecb200::ecdsa::VerifyingKey<NistP224>;
}  // namespace ecdsa
}  // namespace p224
namespace sm2 {
using Curve = Sm2;
using FieldBytes = ecb200::FieldBytes<Sm2>;
using Scalar = ecb200::Scalar<Sm2>;
using AffinePoint = ecb200::AffinePoint<Sm2>;
using ProjectivePoint = ecb200::ProjectivePoint<Sm2>;
using EncodedPoint = ecb200::EncodedPoint<Sm2>;
namespace dsa { using Signature = ecb200::ecdsa::Signature<Sm2>; }
}  // namespace sm2

}  // namespace ecb200
#endif  // ECB200_HPP
