// host_mirror.cpp — the reference's own hot-path tests, replayed through the C++ host mirror (include/ecb200.hpp).
//
// Test infrastructure.  Built by tests/test_host_mirror.py with g++ against include/ and libecb200.so:
//   host_mirror --host-only            host-side logic only (CPU tier; also checks that the engine refuses to start
//                                       without a CUDA device — there is no CPU fallback)
//   host_mirror <fixture.txt>          everything, on cuda:0 (GPU tier).  The fixture is one case per line, written by
//                                       the Python test from tests/golden/*.json (the reference's vectors) and the oracle.
// Each block is named after the reference test it mirrors:
//   test_vector_scalar_mult / test_vector_add_mixed / test_vector_double_generator
//                                       k256/src/arithmetic/projective.rs tests, primeorder/src/dev.rs impl_projective_arithmetic_tests!
//   lincomb, mul_by_generator           k256/src/arithmetic/mul.rs:493-512
//   batch_normalize_array / _slice      k256/src/arithmetic/projective.rs:773-834
//   ecdsa verify (FIPS, Wycheproof)     ecdsa_core::new_verification_test! / new_wycheproof_test! (p256/src/ecdsa.rs:98-197)
//   ecdsa signing KATs                  ecdsa_core::new_signing_test!
//   public_key_recovery                 k256/src/ecdsa.rs:278-343
//   bip340 verify vectors               k256/src/schnorr.rs:291-445
//   sm2dsa verify                       sm2/tests/sm2dsa.rs:16-32
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "ecb200.hpp"

using namespace ecb200;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...)                                                   \
    do {                                                                   \
        g_checks++;                                                        \
        if (!(cond)) {                                                     \
            g_fail++;                                                      \
            std::fprintf(stderr, "FAIL %s:%d: %s | ", __FILE__, __LINE__, #cond); \
            std::fprintf(stderr, __VA_ARGS__);                             \
            std::fprintf(stderr, "\n");                                    \
        }                                                                  \
    } while (0)

static std::vector<uint8_t> unhex(const std::string& s) {
    if (s == "-") return {};
    std::vector<uint8_t> r(s.size() / 2);
    for (size_t i = 0; i < r.size(); i++) r[i] = (uint8_t)(detail::hexval(s[2 * i]) * 16 + detail::hexval(s[2 * i + 1]));
    return r;
}
template <size_t N> static std::array<uint8_t, N> arr(const std::string& s) {
    std::array<uint8_t, N> a{};
    std::vector<uint8_t> v = unhex(s);
    if (v.size() != N) { std::fprintf(stderr, "fixture field has %zu bytes, wanted %zu\n", v.size(), N); std::exit(2); }
    std::memcpy(a.data(), v.data(), N);
    return a;
}
typedef std::vector<std::string> Row;
typedef std::map<std::string, std::vector<Row>> Cases;   // "kind curve" -> rows of hex fields

// ---------------------------------------------------------------------------------------------------- host-only logic
template <class C> static void host_logic() {
    typedef Scalar<C> S;
    auto n = C::order();
    CHECK(!S::from_repr(n), "from_repr(n) must be None (%s)", C::NAME);
    auto nm1 = n;
    nm1[C::FB - 1] -= 1;   // the orders used here end in an odd byte > 0
    CHECK(S::from_repr(nm1).has_value(), "from_repr(n-1)");
    CHECK(S::reduce_bytes(n) == S::zero(), "reduce_bytes(n) == 0");
    FieldBytes<C> ones;
    ones.fill(0xFF);
    CHECK(S::reduce_bytes(ones) == *S::from_repr(detail::sub_be(ones, n)), "reduce_bytes(2^b - 1)");
    CHECK(S::one().negate() == *S::from_repr(nm1), "-1 == n-1");
    CHECK(S::zero().negate() == S::zero(), "-0 == 0");
    CHECK(!S::one().is_high() && S::from_repr(nm1)->is_high(), "is_high");
    S half = *S::from_repr(detail::shr1_be(n));   // floor(n/2) is NOT high, floor(n/2)+1 is
    CHECK(!half.is_high(), "floor(n/2) is low");
    CHECK(half.negate().is_high(), "n - floor(n/2) is high");
    // Signature::from_scalars range rules and normalize_s
    typedef ecdsa::Signature<C> Sig;
    CHECK(!Sig::from_scalars(S::zero().to_repr(), S::one().to_repr()), "r = 0 rejected");
    CHECK(!Sig::from_scalars(S::one().to_repr(), S::zero().to_repr()), "s = 0 rejected");
    CHECK(!Sig::from_scalars(n, S::one().to_repr()), "r = n rejected");
    auto hs = Sig::from_scalars(S::one().to_repr(), nm1);
    CHECK(hs && hs->normalize_s() && hs->normalize_s()->s() == S::one(), "normalize_s(n-1) == 1");
    CHECK(hs && !Sig::from_scalars(S::one().to_repr(), S::one().to_repr())->normalize_s(), "low s stays");
    uint8_t buf[2 * C::FB + 1] = {0};
    CHECK(!Sig::from_slice(buf, 2 * C::FB + 1) && !Sig::from_slice(buf, 2 * C::FB - 1), "from_slice length");
    // bits2field (ecdsa hazmat): too short -> Err, short -> left pad, long -> truncate
    std::vector<uint8_t> h(C::FB + 7, 0xAB);
    CHECK(!ecdsa::bits2field<C>(h.data(), C::FB / 2 - 1), "prehash shorter than FB/2");
    auto z = ecdsa::bits2field<C>(h.data(), C::FB / 2);
    CHECK(z && (*z)[0] == 0 && (*z)[C::FB / 2 - 1] == 0 && (*z)[C::FB / 2] == 0xAB, "left padded");
    z = ecdsa::bits2field<C>(h.data(), h.size());
    CHECK(z && (*z)[0] == 0xAB && (*z)[C::FB - 1] == 0xAB, "truncated");
    // SEC1
    typedef EncodedPoint<C> EP;
    std::vector<uint8_t> e(1 + 2 * C::FB, 1);
    e[0] = 4;
    CHECK(EP::from_bytes(e).has_value(), "tag 04 full length");
    CHECK(!EP::from_bytes(e.data(), 1 + C::FB), "tag 04 short");
    e[0] = 2;
    CHECK(EP::from_bytes(e.data(), 1 + C::FB).has_value() && !EP::from_bytes(e), "tag 02 length");
    e[0] = 6;
    CHECK(!EP::from_bytes(e.data(), 1 + C::FB), "unknown tag");
    e[0] = 0;
    CHECK(EP::from_bytes(e.data(), 1).has_value() && EP::from_bytes(e.data(), 1)->is_identity() && !EP::from_bytes(e.data(), 2), "identity");
    AffinePoint<C> id;
    CHECK(id.is_identity() && id.to_encoded_point(true).len() == 1 && id.to_encoded_point(false).is_identity(), "identity encoding");
    FieldBytes<C> x{}, y{};
    x[0] = 7; y[C::FB - 1] = 3;
    auto P = AffinePoint<C>::from_coordinates_unchecked(x, y);
    CHECK(P.to_encoded_point(true).tag() == 3 && P.to_encoded_point(true).len() == 1 + C::FB, "odd y -> tag 03");
    CHECK(P.to_encoded_point(false).tag() == 4 && P.to_encoded_point().is_compressed() == C::COMPRESS_POINTS, "default compression");
    CHECK(ProjectivePoint<C>::identity().is_identity() && !ProjectivePoint<C>::from_affine(P).is_identity() &&
              ProjectivePoint<C>::from_affine(id).is_identity(), "projective identity");
    CHECK(!ecdsa::VerifyingKey<C>::from_affine(id) && ecdsa::VerifyingKey<C>::from_affine(P), "identity is not a key");
    CHECK(ecb200_field_bytes(C::ID) == C::FB, "field bytes agree with the C ABI");
    CHECK(ecb200_point_slot_bytes(C::ID, 0) == (C::COMPRESS_POINTS ? 1 + C::FB : 1 + 2 * C::FB), "default slot size agrees with the C ABI");
}

static int host_only() {
    host_logic<Secp256k1>();
    host_logic<NistP256>();
    host_logic<NistP384>();
    host_logic<Sm2>();
    host_logic<NistP192>();
    host_logic<NistP224>();
    // shard_range: contiguous, complete, floor boundaries
    for (size_t n : {0ul, 1ul, 7ul, 4194304ul, 4194305ul})
        for (size_t g : {1ul, 2ul, 3ul, 8ul}) {
            size_t prev = 0;
            for (size_t r = 0; r < g; r++) {
                auto s = shard_range(n, r, g);
                CHECK(s.first == prev && s.second >= s.first && s.second == (r + 1) * n / g, "shard %zu/%zu of %zu", r, g, n);
                prev = s.second;
            }
            CHECK(prev == n, "shards cover the batch");
        }
    auto rid = ecdsa::RecoveryId::from_byte(3);
    CHECK(rid && rid->is_y_odd() && rid->is_x_reduced() && !ecdsa::RecoveryId::from_byte(4), "RecoveryId");
    return 0;
}

// ---------------------------------------------------------------------------------------------------- device tests
template <class C> static const std::vector<Row>& rows(const Cases& cs, const char* kind) {
    static const std::vector<Row> none;
    auto it = cs.find(std::string(kind) + " " + C::NAME);
    return it == cs.end() ? none : it->second;
}
template <class C> static AffinePoint<C> pt(const std::string& x, const std::string& y) {
    return AffinePoint<C>::from_coordinates_unchecked(arr<C::FB>(x), arr<C::FB>(y));
}

template <class C> static void group_tests(Engine& eng, const Cases& cs) {
    typedef Scalar<C> S;
    typedef ProjectivePoint<C> PP;
    typedef AffinePoint<C> AP;
    const auto& add = rows<C>(cs, "add");   // add[i] = (i+1) * G
    const auto& mul = rows<C>(cs, "mul");   // (k, x, y): k * G
    if (add.empty()) return;
    const AP G = pt<C>(add[0][0], add[0][1]);
    const PP Gp = PP::from_affine(G);

    // test_vector_scalar_mult + mul_by_generator (mul.rs:504-512): G * k == mul_by_generator(k) == vector
    std::vector<S> ks;
    std::vector<std::pair<PP, S>> terms;
    for (size_t i = 0; i < add.size(); i++) ks.push_back(S::from_u64(i + 1));
    for (const Row& r : mul) ks.push_back(*S::from_repr(arr<C::FB>(r[0])));
    ks.push_back(S::zero());
    for (const S& k : ks) terms.push_back({Gp, k});
    auto a = PP::mul_by_generator_batch(eng, ks);
    auto b = PP::mul_batch(eng, terms);
    auto c = PP::mul_batch_vartime(eng, terms);
    auto enc = PP::mul_by_generator_batch_encoded(eng, ks);
    for (size_t i = 0; i < ks.size(); i++) {
        AP want = i < add.size() ? pt<C>(add[i][0], add[i][1]) : i < add.size() + mul.size() ? pt<C>(mul[i - add.size()][1], mul[i - add.size()][2]) : AP::identity();
        CHECK(a[i] == want, "%s mul_by_generator row %zu", C::NAME, i);
        CHECK(b[i] == want, "%s G * k (constant time) row %zu", C::NAME, i);
        CHECK(c[i] == want, "%s G * k (vartime) row %zu", C::NAME, i);
        CHECK(enc[i] == want.to_encoded_point(C::COMPRESS_POINTS), "%s encoded row %zu", C::NAME, i);
    }
    // test_vector_add_mixed: (i+1)G + G == (i+2)G;  test_vector_double_generator: 2 * (2^i G) chain
    std::vector<PP> sums;
    for (size_t i = 0; i + 1 < add.size(); i++) sums.push_back(PP::lincomb(eng, PP::from_affine(pt<C>(add[i][0], add[i][1])), S::one(), Gp, S::one()));
    PP d = Gp;
    std::vector<PP> dbl;
    for (size_t i = 2; i <= add.size(); i *= 2) {
        d = PP::lincomb_ext(eng, {{d, S::from_u64(2)}});
        dbl.push_back(d);
    }
    // batch_normalize_slice: one call normalises everything (the lincomb outputs carry non-unit Z)
    std::vector<PP> all = sums;
    all.insert(all.end(), dbl.begin(), dbl.end());
    all.push_back(PP::identity());
    all.push_back(PP::lincomb(eng, Gp, S::one(), Gp, S::one().negate()));   // G - G: the identity with whatever X, Y came out
    auto aff = PP::batch_normalize(eng, all);
    for (size_t i = 0; i < sums.size(); i++) CHECK(aff[i] == pt<C>(add[i + 1][0], add[i + 1][1]), "%s add row %zu", C::NAME, i);
    size_t j = sums.size();
    for (size_t i = 2; i <= add.size(); i *= 2, j++) CHECK(aff[j] == pt<C>(add[i - 1][0], add[i - 1][1]), "%s doubling to %zu G", C::NAME, i);
    CHECK(aff[j].is_identity() && aff[j + 1].is_identity(), "%s identity normalises to AffinePoint::IDENTITY", C::NAME);
    CHECK(all.back().is_identity(), "%s G - G has Z = 0", C::NAME);
    // batch_normalize_array: batch result == one-at-a-time to_affine
    if (sums.size() >= 2) CHECK(sums[0].to_affine(eng) == aff[0] && sums[1].to_affine(eng) == aff[1], "%s to_affine", C::NAME);
    // lincomb (mul.rs:493-501): lincomb(x, k, y, l) == x*k + y*l
    if (mul.size() >= 2) {
        S k = *S::from_repr(arr<C::FB>(mul[0][0])), l = *S::from_repr(arr<C::FB>(mul[1][0]));
        PP x = PP::from_affine(pt<C>(add[2][0], add[2][1])), y = PP::from_affine(pt<C>(add[4][0], add[4][1]));
        PP ref = PP::lincomb(eng, x, k, y, l);
        auto parts = PP::mul_batch(eng, {{x, k}, {y, l}});
        PP sum = PP::lincomb(eng, PP::from_affine(parts[0]), S::one(), PP::from_affine(parts[1]), S::one());
        CHECK(ref.to_affine(eng) == sum.to_affine(eng), "%s lincomb == x*k + y*l", C::NAME);
    }
}

template <class C> static void decode_tests(Engine& eng, const Cases& cs) {
    // FromEncodedPoint: rows (encoding, status, x, y)
    std::vector<EncodedPoint<C>> enc;
    std::vector<const Row*> keep;
    for (const Row& r : rows<C>(cs, "decode")) {
        auto e = EncodedPoint<C>::from_bytes(unhex(r[0]));
        if (!e) { CHECK(r[1] == "00", "%s host-rejected encoding must be invalid in the oracle too", C::NAME); continue; }
        enc.push_back(*e);
        keep.push_back(&r);
    }
    auto got = AffinePoint<C>::from_encoded_points(eng, enc);
    for (size_t i = 0; i < keep.size(); i++) {
        const Row& r = *keep[i];
        if (r[1] == "00") CHECK(!got[i], "%s decode row %zu must fail", C::NAME, i);
        else if (r[1] == "02") CHECK(got[i] && got[i]->is_identity(), "%s decode row %zu identity", C::NAME, i);
        else CHECK(got[i] && *got[i] == pt<C>(r[2], r[3]), "%s decode row %zu", C::NAME, i);
    }
    // to_encoded_point / from_encoded_point round trip in both encodings
    std::vector<EncodedPoint<C>> back;
    for (auto& g : got) if (g) { back.push_back(g->to_encoded_point(true)); back.push_back(g->to_encoded_point(false)); }
    auto again = AffinePoint<C>::from_encoded_points(eng, back);
    size_t k = 0;
    for (auto& g : got) if (g) { CHECK(again[k] && *again[k] == *g && again[k + 1] && *again[k + 1] == *g, "%s sec1 round trip", C::NAME); k += 2; }
}

template <class C> static void ecdsa_tests(Engine& eng, const Cases& cs) {
    typedef ecdsa::Signature<C> Sig;
    typedef ecdsa::VerifyingKey<C> VK;
    // verify rows: qx qy prehash r s expect
    std::vector<VK> keys;
    std::vector<std::vector<uint8_t>> hs;
    std::vector<Sig> sigs;
    std::vector<bool> exp;
    for (const Row& r : rows<C>(cs, "verify")) {
        auto sig = Sig::from_scalars(arr<C::FB>(r[3]), arr<C::FB>(r[4]));
        if (!sig) { CHECK(r[5] == "00", "%s signature rejected at construction must be invalid", C::NAME); continue; }
        keys.push_back(*VK::from_affine(pt<C>(r[0], r[1])));
        hs.push_back(unhex(r[2]));
        sigs.push_back(*sig);
        exp.push_back(r[5] == "01");
    }
    auto got = VK::verify_prehash_batch(eng, keys, hs, sigs);
    size_t accepted = 0;
    for (size_t i = 0; i < got.size(); i++) {
        CHECK(got[i].is_ok() == exp[i], "%s verify row %zu: got %d", C::NAME, i, (int)got[i].is_ok());
        accepted += got[i].is_ok();
    }
    if (!got.empty()) CHECK(accepted > 0 && accepted < got.size(), "%s verify rows cover both outcomes", C::NAME);
    // keys given as SEC1 bytes (VerifyingKey::from_sec1_bytes): same answers, both encodings
    std::vector<std::vector<uint8_t>> sec1;
    for (size_t i = 0; i < keys.size(); i++) sec1.push_back(keys[i].to_encoded_point(i % 2 == 0).to_vec());
    auto parsed = VK::from_sec1_bytes_batch(eng, sec1);
    std::vector<VK> k2;
    std::vector<std::vector<uint8_t>> h2;
    std::vector<Sig> s2;
    std::vector<bool> e2;
    for (size_t i = 0; i < parsed.size(); i++) {
        if (!parsed[i]) { CHECK(!exp[i], "%s undecodable key on an accepted row %zu", C::NAME, i); continue; }
        k2.push_back(*parsed[i]); h2.push_back(hs[i]); s2.push_back(sigs[i]); e2.push_back(exp[i]);
    }
    auto got2 = VK::verify_prehash_batch(eng, k2, h2, s2);
    for (size_t i = 0; i < got2.size(); i++) CHECK(got2[i].is_ok() == e2[i], "%s verify (sec1 key) row %zu", C::NAME, i);

    // signing KATs: d k z r s recid(or -)
    std::vector<Scalar<C>> d, k;
    std::vector<FieldBytes<C>> z;
    const auto& sr = rows<C>(cs, "sign");
    for (const Row& r : sr) {
        d.push_back(*Scalar<C>::from_repr(arr<C::FB>(r[0])));
        k.push_back(*Scalar<C>::from_repr(arr<C::FB>(r[1])));
        z.push_back(arr<C::FB>(r[2]));
    }
    auto signed_ = ecdsa::try_sign_prehashed_batch<C>(eng, d, k, z);
    std::vector<VK> vk;
    std::vector<std::vector<uint8_t>> vh;
    std::vector<Sig> vs;
    std::vector<ecdsa::RecoveryId> vid;
    auto pubs = ProjectivePoint<C>::mul_by_generator_batch(eng, d);
    for (size_t i = 0; i < sr.size(); i++) {
        CHECK(signed_[i].has_value(), "%s signing row %zu", C::NAME, i);
        if (!signed_[i]) continue;
        CHECK(signed_[i]->first.r().to_repr() == arr<C::FB>(sr[i][3]) && signed_[i]->first.s().to_repr() == arr<C::FB>(sr[i][4]),
              "%s signing KAT row %zu", C::NAME, i);
        if (sr[i][5] != "-") CHECK(signed_[i]->second.to_byte() == unhex(sr[i][5])[0], "%s recovery id row %zu", C::NAME, i);
        vk.push_back(*VK::from_affine(pubs[i]));
        vh.push_back(std::vector<uint8_t>(z[i].begin(), z[i].end()));
        vs.push_back(signed_[i]->first);
        vid.push_back(signed_[i]->second);
    }
    // sign -> verify -> recover round trip
    auto v = VK::verify_prehash_batch(eng, vk, vh, vs);
    auto rec = VK::recover_from_prehash_batch(eng, vh, vs, vid);
    for (size_t i = 0; i < v.size(); i++) {
        CHECK(v[i].is_ok(), "%s own signature verifies, row %zu", C::NAME, i);
        CHECK(rec[i] && *rec[i] == vk[i], "%s recovered key equals the signer's, row %zu", C::NAME, i);
    }
    // public_key_recovery vectors: prehash r s recid expected-sec1 (or -)
    std::vector<std::vector<uint8_t>> rh;
    std::vector<Sig> rsg;
    std::vector<ecdsa::RecoveryId> rid;
    std::vector<std::string> want;
    for (const Row& r : rows<C>(cs, "recover")) {
        auto sig = Sig::from_scalars(arr<C::FB>(r[1]), arr<C::FB>(r[2]));
        auto id = ecdsa::RecoveryId::from_byte(unhex(r[3])[0]);
        if (!sig || !id) { CHECK(r[4] == "-", "%s recovery row rejected at construction must fail in the oracle", C::NAME); continue; }
        rh.push_back(unhex(r[0])); rsg.push_back(*sig); rid.push_back(*id); want.push_back(r[4]);
    }
    auto rk = VK::recover_from_prehash_batch(eng, rh, rsg, rid);
    for (size_t i = 0; i < rk.size(); i++) {
        if (want[i] == "-") CHECK(!rk[i], "%s recovery row %zu must fail", C::NAME, i);
        else CHECK(rk[i] && rk[i]->to_encoded_point(true).to_vec() == unhex(want[i]), "%s recovery row %zu", C::NAME, i);
    }
}

static void schnorr_tests(Engine& eng, const Cases& cs) {
    auto it = cs.find("schnorr k256");
    if (it == cs.end()) return;
    std::vector<schnorr::VerifyingKey> keys;
    std::vector<std::array<uint8_t, 32>> e;
    std::vector<std::array<uint8_t, 64>> sigs;
    for (const Row& r : it->second) {
        keys.push_back(schnorr::VerifyingKey{arr<32>(r[0])});
        e.push_back(arr<32>(r[1]));
        sigs.push_back(arr<64>(r[2]));
    }
    auto got = schnorr::VerifyingKey::verify_raw_batch(eng, keys, e, sigs);
    for (size_t i = 0; i < got.size(); i++) CHECK(got[i].is_ok() == (it->second[i][3] == "01"), "bip340 row %zu", i);
}

static void sm2dsa_tests(Engine& eng, const Cases& cs) {
    auto it = cs.find("sm2dsa sm2");
    if (it == cs.end()) return;
    std::vector<AffinePoint<Sm2>> keys;
    std::vector<FieldBytes<Sm2>> e;
    std::vector<ecdsa::Signature<Sm2>> sigs;
    std::vector<bool> exp;
    for (const Row& r : it->second) {
        auto sig = ecdsa::Signature<Sm2>::from_scalars(arr<32>(r[3]), arr<32>(r[4]));
        if (!sig) { CHECK(r[5] == "00", "sm2dsa signature rejected at construction must be invalid"); continue; }
        keys.push_back(pt<Sm2>(r[0], r[1]));
        e.push_back(arr<32>(r[2]));
        sigs.push_back(*sig);
        exp.push_back(r[5] == "01");
    }
    auto got = sm2dsa::verify_prehash_batch(eng, keys, e, sigs);
    for (size_t i = 0; i < got.size(); i++) CHECK(got[i].is_ok() == exp[i], "sm2dsa row %zu", i);
}

template <class C> static void curve_tests(Engine& eng, const Cases& cs) {
    int before = g_fail;
    group_tests<C>(eng, cs);
    decode_tests<C>(eng, cs);
    ecdsa_tests<C>(eng, cs);
    std::printf("%s: %s\n", C::NAME, g_fail == before ? "ok" : "FAILED");
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: host_mirror --host-only | <fixture.txt>\n"); return 2; }
    host_only();
    if (std::string(argv[1]) == "--host-only") {
        // no CUDA device in the CPU tier: construction must fail loudly instead of falling back to a CPU path
        if (argc > 2 && std::string(argv[2]) == "--expect-no-device") {
            bool threw = false;
            try { Engine e(0); } catch (const Error& err) { threw = true; std::printf("engine refused: %s\n", err.what()); }
            CHECK(threw, "Engine() must throw without a CUDA device");
        }
        std::printf("host mirror host-only: %d checks, %d failures\n", g_checks, g_fail);
        return g_fail ? 1 : 0;
    }
    std::ifstream in(argv[1]);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    Cases cs;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string kind, curve, f;
        ss >> kind >> curve;
        Row r;
        while (ss >> f) r.push_back(f);
        cs[kind + " " + curve].push_back(r);
    }
    try {
        Engine eng(0);
        curve_tests<Secp256k1>(eng, cs);
        curve_tests<NistP256>(eng, cs);
        curve_tests<NistP384>(eng, cs);
        curve_tests<Sm2>(eng, cs);
        curve_tests<NistP192>(eng, cs);
        curve_tests<NistP224>(eng, cs);
        schnorr_tests(eng, cs);
        sm2dsa_tests(eng, cs);
        CHECK(eng.launch_count() > 0, "kernels were launched");
        // error behaviour: engine-level failures are exceptions carrying the ecb200_status
        bool threw = false;
        try { eng.check(ecb200_mul_gen(eng.raw(), 99, 1, nullptr, nullptr, 0), "mul_gen"); } catch (const Error& e) { threw = e.status() == ECB200_ERR_ARG; }
        CHECK(threw, "bad curve id -> Error(ECB200_ERR_ARG)");
    } catch (const Error& e) {
        std::fprintf(stderr, "engine error: %s\n", e.what());
        return 3;
    }
    std::printf("host mirror: %d checks, %d failures\n", g_checks, g_fail);
    if (!g_fail) std::printf("host mirror ok\n");
    return g_fail ? 1 : 0;
}
