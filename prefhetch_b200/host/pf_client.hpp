// pf_client.hpp — the CLIENT side of the encrypted coarse search: what the reference's
// client does around `get_coarse_scores` (ref: src/client/client_lib.cpp:83-156) once the query is sent
// encrypted — the step include/client/client_lib.h:33-35 leaves as a commented-out prototype
// (`compute_encrypted_coarse_query`).  Header-only C++17, CPU only, no SEAL and no CUDA: a client has no GPU.
//
//   key generation      ternary secret key; SEAL-format GaloisKeys for the R-1 row rotations the server hoists
//   query encryption    BatchEncoder::encode of the query replicated with period d_pad/m, symmetric BFV
//                       encryption, written as SEAL streams (full, or seeded: c0 + the 64-byte PRNG seed of c1)
//   response decryption result ciphertexts (any level the server mod-switched to) -> decrypt -> decode -> the
//                       reference's packed `coarseDistanceScores` (exact integer squared L2 as floats), then
//                       `compute_nearest_coarse_vectors` (ref: client_lib.cpp:122-156) as it is today
//
// Conventions are SEAL 4.1's, restated from the published sources [EXT]: negacyclic NTT over the smallest
// primitive 2N-th root with bit-reversed output (util/ntt.cpp), BatchEncoder's matrix_reps_index_map
// (batchencoder.cpp), encrypt_zero_symmetric + multiply_add_plain_with_scaling_variant (util/rlwe.cpp,
// util/scalingvariant.cpp), the centred-binomial noise of util/clipnormal.h `cbd`, generate_one_kswitch_key
// (keygenerator.cpp), Ciphertext / KSwitchKeys::save_members (ciphertext.cpp, kswitchkeys.h), parms_id =
// BLAKE2b-256 of {scheme, N, primes, t} (encryptionparams.cpp).  Randomness comes from seal::Blake2xbPRNG
// (csrc/pf_seal_prng.h); the draws are not SEAL's draw for draw (its ternary sampler goes through
// std::uniform_int_distribution), which no party can observe.  The slot layout is the one the engine encodes the
// database in (csrc/pf_encode.cuh) — restated here from its definition, not shared code.
// Product code: independent of oracle/.  tests/test_client.py plays this client against the CPU oracle as the
// server (bit-exact streams in, exact distances out); tests/test_gpu_parity.py against the engine.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csrc/pf_blake2b.h"
#include "../csrc/pf_host_math.h"
#include "../csrc/pf_seal_prng.h"
#include "pf_json.hpp"

namespace prefhetch {

using pfh::u128;
using pfh::u64;

// ref: include/client/client_lib.h:9-12
struct DistanceIndexData {
    float distance;
    int64_t idx;
};

namespace detail {

inline u64 add_mod(u64 a, u64 b, u64 q) {
    const u64 s = a + b;
    return s >= q ? s - q : s;
}
inline u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
inline u64 signed_mod(int64_t v, u64 q) { return v >= 0 ? (u64)v % q : q - 1 - ((u64)(-(v + 1)) % q); }

// negacyclic NTT of length n modulo a prime q = 1 (mod 2n): forward takes coefficients in natural order to
// evaluations at psi^(2*bitrev(i)+1); inverse undoes it
struct NttPlan {
    u64 q = 0, n = 0, ninv = 0;
    int logn = 0;
    std::vector<u64> w, winv;     // psi^bitrev(i), psi^-bitrev(i)
    std::vector<u64> wsh, winvsh; // floor(w * 2^64 / q): Shoup quotients, one high multiply instead of a division
    u64 ninvsh = 0;

    // x * w mod q for x < q, with wq = floor(w * 2^64 / q); q < 2^63
    static u64 mul_shoup(u64 x, u64 w_, u64 wq, u64 q_) {
        const u64 hi = (u64)(((u128)x * wq) >> 64);
        const u64 r = x * w_ - hi * q_;
        return r >= q_ ? r - q_ : r;
    }

    NttPlan() = default;
    NttPlan(u64 n_, u64 q_) : q(q_), n(n_) {
        while (((u64)1 << logn) < n) logn++;
        const u64 psi = pfh::minimal_primitive_root(2 * n, q);
        if (!psi) throw std::invalid_argument("modulus " + std::to_string(q) + " has no primitive 2N-th root");
        const u64 ipsi = pfh::invmod(psi, q);
        w.assign(n, 0);
        winv.assign(n, 0);
        u64 p = 1, ip = 1;
        for (u64 i = 0; i < n; i++) {
            const uint32_t r = pfh::bitrev((uint32_t)i, logn);
            w[r] = p;
            winv[r] = ip;
            p = pfh::mulmod(p, psi, q);
            ip = pfh::mulmod(ip, ipsi, q);
        }
        ninv = pfh::invmod(n % q, q);
        wsh.resize(n);
        winvsh.resize(n);
        for (u64 i = 0; i < n; i++) {
            wsh[i] = pfh::shoup(w[i], q);
            winvsh[i] = pfh::shoup(winv[i], q);
        }
        ninvsh = pfh::shoup(ninv, q);
    }
    void forward(u64 *a) const {
        u64 t = n;
        for (u64 m = 1; m < n; m <<= 1) {
            t >>= 1;
            for (u64 i = 0; i < m; i++) {
                const u64 W = w[m + i], Wq = wsh[m + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (u64 j = 0; j < t; j++) {
                    const u64 u = x[j], v = mul_shoup(y[j], W, Wq, q);
                    x[j] = add_mod(u, v, q);
                    y[j] = sub_mod(u, v, q);
                }
            }
        }
    }
    void inverse(u64 *a) const {
        u64 t = 1;
        for (u64 m = n; m > 1; m >>= 1) {
            const u64 h = m >> 1;
            for (u64 i = 0; i < h; i++) {
                const u64 W = winv[h + i], Wq = winvsh[h + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (u64 j = 0; j < t; j++) {
                    const u64 u = x[j], v = y[j];
                    x[j] = add_mod(u, v, q);
                    y[j] = mul_shoup(sub_mod(u, v, q), W, Wq, q);
                }
            }
            t <<= 1;
        }
        for (u64 i = 0; i < n; i++) a[i] = mul_shoup(a[i], ninv, ninvsh, q);
    }
};

// little-endian multiword naturals, just what decryption needs
using Big = std::vector<u64>;
inline void big_trim(Big &a) {
    while (a.size() > 1 && a.back() == 0) a.pop_back();
}
inline void big_add_word(Big &a, u64 x) {
    for (size_t i = 0; i < a.size() && x; i++) {
        const u64 s = a[i] + x;
        x = s < x ? 1 : 0;
        a[i] = s;
    }
    if (x) a.push_back(x);
}
inline void big_add(Big &a, const Big &b) {
    if (a.size() < b.size()) a.resize(b.size(), 0);
    u64 carry = 0;
    for (size_t i = 0; i < a.size(); i++) {
        const u128 s = (u128)a[i] + (i < b.size() ? b[i] : 0) + carry;
        a[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    if (carry) a.push_back(carry);
}
inline int big_cmp(const Big &a, const Big &b) {
    const size_t n = std::max(a.size(), b.size());
    for (size_t i = n; i-- > 0;) {
        const u64 x = i < a.size() ? a[i] : 0, y = i < b.size() ? b[i] : 0;
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
inline Big big_sub(const Big &a, const Big &b) { // a >= b
    Big r(a);
    u64 borrow = 0;
    for (size_t i = 0; i < r.size(); i++) {
        const u64 y = i < b.size() ? b[i] : 0;
        const u64 d = r[i] - y - borrow;
        borrow = (r[i] < y || (borrow && r[i] == y)) ? 1 : 0;
        r[i] = d;
    }
    return r;
}
inline int big_bits(const Big &a) {
    for (size_t i = a.size(); i-- > 0;)
        if (a[i]) return (int)(64 * i) + 64 - __builtin_clzll(a[i]);
    return 0;
}

inline void put64(std::vector<uint8_t> &o, u64 v) {
    uint8_t b[8];
    memcpy(b, &v, 8);
    o.insert(o.end(), b, b + 8);
}
inline void put_seal_header(std::vector<uint8_t> &o, u64 total) {
    const uint8_t h[8] = {0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0}; // magic, header size, version 4.1, compr_mode none
    o.insert(o.end(), h, h + 8);
    put64(o, total);
}

} // namespace detail

constexpr size_t SEAL_CT_PREFIX = 113; // bytes before the words of a serialized ciphertext
constexpr size_t SEAL_SEED_INFO = 81;  // UniformRandomGeneratorInfo stream: header 16 + type 1 + seed 64

// The response of POST /coarsesearch-encrypted (pf_query_handlers.hpp) as the client holds it after parsing
struct EncryptedCoarseResponse {
    std::vector<uint8_t> ciphertexts;             // result streams in the server's aligned slots
    std::vector<uint64_t> result_offsets;         // [nresults + 1]: start of every stream
    uint64_t result_bytes = 0;                    // length of every stream
    std::vector<uint64_t> results_per_query;      // [nq]
    std::vector<int64_t> coarse_vector_indexes;   // labels, packed per query (as in Query.cc:53-61)
    std::vector<uint64_t> list_sizes_per_query;   // [nq]
    std::vector<uint64_t> probed_sizes;           // [nq][nprobe]
    size_t nq = 0, nprobe = 0;
};

class Client {
  public:
    // primes: the k = L + 1 coefficient primes, special prime last (what the server was created with);
    // dim, m, g: the layout parameters of the index (prefhetch::Server / pf_create)
    // (public signatures use uint64_t like prefhetch::Server; u64 inside is the same width)
    Client(uint32_t dim, uint64_t poly_degree, const std::vector<uint64_t> &primes, uint64_t plain_modulus, uint32_t m = 1, uint32_t g = 16)
        : d_(dim), N_(poly_degree), q_(primes.begin(), primes.end()), t_(plain_modulus), m_(m), g_(g) {
        if (q_.size() < 2) throw std::invalid_argument("need at least one data prime and the special prime");
        if (N_ < 8 || (N_ & (N_ - 1))) throw std::invalid_argument("poly_degree must be a power of two");
        k_ = (uint32_t)q_.size();
        L_ = k_ - 1;
        dpad_ = 1;
        while (dpad_ < d_) dpad_ <<= 1;
        if (!m_ || !g_ || (m_ & (m_ - 1)) || (g_ & (g_ - 1)) || dpad_ % m_) throw std::invalid_argument("bad layout (m, g)");
        dc_ = dpad_ / m_;
        if (dc_ % g_ || dc_ > N_ / 2) throw std::invalid_argument("bad layout (g does not divide the chunk, or chunk > N/2)");
        R_ = dc_ / g_;
        C_ = (uint32_t)(N_ / g_);
        for (u64 q : q_) {
            if (q >> 61) throw std::invalid_argument("coefficient primes above 61 bits are not supported (SEAL's own limit)");
            ntt_.emplace_back(N_, q);
        }
        ntt_t_ = detail::NttPlan(N_, t_);
        // BatchEncoder: slot i of row 0 is the evaluation at zeta^(3^i), of row 1 at zeta^(-3^i)
        slot_to_coeff_.resize(N_);
        const u64 two_n = 2 * N_, half = N_ / 2;
        u64 pos = 1;
        for (u64 i = 0; i < half; i++) {
            slot_to_coeff_[i] = pfh::bitrev((uint32_t)((pos - 1) >> 1), ntt_t_.logn);
            slot_to_coeff_[half + i] = pfh::bitrev((uint32_t)((two_n - pos - 1) >> 1), ntt_t_.logn);
            pos = pos * 3 % two_n;
        }
        for (uint32_t l = 1; l <= L_; l++) levels_.push_back(make_level(l));
        for (uint32_t j = 0; j < L_; j++) p_mod_q_.push_back(q_[k_ - 1] % q_[j]);
    }

    uint32_t candidatesPerResult() const { return C_; }
    uint32_t rotations() const { return R_; }
    uint32_t queryCiphertexts() const { return m_; }
    uint32_t dataLimbs() const { return L_; }
    u64 polyDegree() const { return N_; }

    // SEAL parms_id of the level with `limbs` data primes (limbs = L + 1: the key level)
    std::array<u64, 4> parmsId(uint32_t limbs) const {
        std::vector<u64> words = {1 /* scheme_type::bfv */, N_};
        for (uint32_t j = 0; j < limbs && j < k_; j++) words.push_back(q_[j]);
        words.push_back(t_);
        std::array<u64, 4> id;
        pfh::blake2b(words.data(), words.size() * 8, id.data(), 32);
        return id;
    }

    static std::array<uint8_t, 64> randomSeed() {
        std::array<uint8_t, 64> s{};
        FILE *f = fopen("/dev/urandom", "rb");
        const bool ok = f && fread(s.data(), 1, 64, f) == 64;
        if (f) fclose(f);
        if (!ok) throw std::runtime_error("cannot read /dev/urandom");
        return s;
    }

    // Secret key from a 64-byte seed; every later draw (noise, ciphertext seeds) continues the same PRNG, so a
    // seed fixes the whole transcript (tests) — use randomSeed() outside tests.
    void generateKeys(const std::array<uint8_t, 64> &seed) {
        prng_.reset(new pfh::SealBlake2xbPrng(seed.data()));
        sk_coeff_.assign(N_, 0);
        for (u64 i = 0; i < N_;) { // uniform over {-1, 0, 1}: two bits at a time, 3 rejected
            uint8_t b;
            prng_->generate(1, &b);
            for (int s = 0; s < 8 && i < N_; s += 2) {
                const int v = (b >> s) & 3;
                if (v != 3) sk_coeff_[i++] = (int8_t)(v - 1);
            }
        }
        sk_ntt_.assign((size_t)k_ * N_, 0);
        for (uint32_t j = 0; j < k_; j++) {
            u64 *s = sk_ntt_.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) s[i] = detail::signed_mod(sk_coeff_[i], q_[j]);
            ntt_[j].forward(s);
        }
    }
    const std::vector<int8_t> &secretKeyCoefficients() const { return sk_coeff_; }

    // seal::BatchEncoder::encode / decode of N slot values mod t
    std::vector<u64> encode(const std::vector<u64> &slots) const {
        std::vector<u64> plain(N_, 0);
        for (u64 i = 0; i < N_; i++) plain[slot_to_coeff_[i]] = slots[i] % t_;
        ntt_t_.inverse(plain.data());
        return plain;
    }
    std::vector<u64> decode(const std::vector<u64> &plain) const {
        std::vector<u64> tmp(plain), slots(N_);
        ntt_t_.forward(tmp.data());
        for (u64 i = 0; i < N_; i++) slots[i] = tmp[slot_to_coeff_[i]];
        return slots;
    }

    // slot vector of query chunk a: coordinate a*dc + (s mod dc) in slot s of both rows
    std::vector<u64> querySlots(const int64_t *q, uint32_t a) const {
        std::vector<u64> slots(N_, 0);
        const u64 half = N_ / 2;
        for (u64 s = 0; s < half; s++) {
            const uint32_t dim = a * dc_ + (uint32_t)(s % dc_);
            const u64 v = dim < d_ ? detail::signed_mod(q[dim], t_) : 0;
            slots[s] = v;
            slots[half + s] = v;
        }
        return slots;
    }
    // slot of partial sum j of candidate u of a result ciphertext
    uint32_t resultSlot(uint32_t u, uint32_t j) const {
        const uint32_t half = (uint32_t)(N_ / 2), per_row = half / g_;
        const uint32_t row = u / per_row, c = u % per_row;
        return row * half + (c % R_) + (c / R_) * dc_ + j * R_;
    }

    // Encryptor::encrypt_symmetric(plain) -> words [2][L][N], coefficient form.  c1 is the expansion of `ct_seed`
    // (sample_poly_uniform of a Blake2xbPRNG), which is all a seeded stream carries of it.
    std::vector<u64> encryptSymmetric(const std::vector<u64> &plain, const uint8_t ct_seed[64]) {
        need_keys();
        std::vector<u64> ct((size_t)2 * L_ * N_);
        u64 *c0 = ct.data(), *c1 = c0 + (size_t)L_ * N_;
        pfh::SealBlake2xbPrng a_prng(ct_seed);
        pfh::seal_sample_poly_uniform(a_prng, reinterpret_cast<const uint64_t *>(q_.data()), L_, N_, reinterpret_cast<uint64_t *>(c1));
        std::vector<int64_t> e(N_);
        sample_noise(e.data());
        const Level &lv = levels_[L_ - 1];
        const u64 half_t = (t_ + 1) >> 1;
        for (uint32_t j = 0; j < L_; j++) {
            const u64 q = q_[j];
            u64 *o = c0 + (size_t)j * N_;
            memcpy(o, c1 + (size_t)j * N_, N_ * 8);
            ntt_[j].forward(o);
            const u64 *s = sk_ntt_.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) o[i] = pfh::mulmod(o[i], s[i], q);
            ntt_[j].inverse(o);
            for (u64 i = 0; i < N_; i++) { // c0 = -(a*s + e) + round(Q*m/t)
                const u64 as_e = detail::add_mod(o[i], detail::signed_mod(e[i], q), q);
                const u128 prod = (u128)plain[i] * lv.q_mod_t + half_t;
                const u64 fix = (u64)(prod / t_);
                const u64 scaled = (u64)(((u128)plain[i] * lv.q_div_t_mod[j] + fix) % q);
                o[i] = detail::add_mod(as_e ? q - as_e : 0, scaled, q);
            }
        }
        return ct;
    }

    // Ciphertext::save, compr_mode none: the full stream, or the seeded one (c0 + seed of c1)
    std::vector<uint8_t> saveCiphertext(const std::vector<u64> &ct, const uint8_t *ct_seed /* NULL: full */) const {
        const u64 poly_words = (u64)L_ * N_, words = ct_seed ? poly_words : 2 * poly_words;
        const u64 total = SEAL_CT_PREFIX + words * 8 + (ct_seed ? SEAL_SEED_INFO : 0);
        std::vector<uint8_t> o;
        o.reserve(total);
        put_ct_prefix(o, total, parmsId(L_), false, L_, words);
        const uint8_t *w = reinterpret_cast<const uint8_t *>(ct.data());
        o.insert(o.end(), w, w + words * 8);
        if (ct_seed) {
            detail::put_seal_header(o, SEAL_SEED_INFO);
            o.push_back(1); // prng_type::blake2xb
            o.insert(o.end(), ct_seed, ct_seed + 64);
        }
        return o;
    }

    // The encrypted query of the additive endpoint (the prototype of client_lib.h:33-35): the m ciphertexts of one
    // query vector as SEAL streams, back to back; offsets gets m + 1 entries relative to the start of the blob.
    std::vector<uint8_t> compute_encrypted_coarse_query(const int64_t *query, std::vector<uint64_t> *offsets, bool seeded = true) {
        need_keys();
        std::vector<uint8_t> blob;
        if (offsets) offsets->assign(1, 0);
        for (uint32_t a = 0; a < m_; a++) {
            uint8_t seed[64];
            prng_->generate(64, seed);
            const std::vector<u64> ct = encryptSymmetric(encode(querySlots(query, a)), seed);
            const std::vector<uint8_t> s = saveCiphertext(ct, seeded ? seed : nullptr);
            blob.insert(blob.end(), s.begin(), s.end());
            if (offsets) offsets->push_back(blob.size());
        }
        return blob;
    }

    // KeyGenerator::create_galois_keys(steps 1..R-1).save(): the keys the server's hoisted rotations use
    // (pf_load_galois_keys / Server::setGaloisKeys).  Slot of element e is (e - 1) / 2; N slots in the stream.
    // seeded (the default, like Serializable<GaloisKeys>): every key ciphertext carries c0 and the seed of c1 — half
    // the bytes; the draws are the same either way, so both forms of one key set describe the same keys.
    std::vector<uint8_t> galoisKeys(bool seeded = true) {
        need_keys();
        const std::array<u64, 4> key_id = parmsId(k_);
        const u64 key_words = seeded ? (u64)k_ * N_ : (u64)2 * k_ * N_;
        const u64 key_stream = SEAL_CT_PREFIX + key_words * 8 + (seeded ? SEAL_SEED_INFO : 0);
        std::vector<std::vector<uint8_t>> slot(N_);
        for (uint32_t step = 1; step < R_; step++) {
            const u64 elt = pfh::powmod(3, step, 2 * N_);
            std::vector<uint8_t> &o = slot[(elt - 1) >> 1];
            const std::vector<u64> rot = rotated_secret_key(elt);
            for (uint32_t J = 0; J < L_; J++) {
                uint8_t seed[64];
                prng_->generate(64, seed);
                std::vector<u64> kw = encrypt_zero_key_level(seed);
                // c0 limb J += (P mod q_J) * sigma_elt(s)
                const u64 q = q_[J];
                u64 *dst = kw.data() + (size_t)J * N_;
                const u64 *r = rot.data() + (size_t)J * N_;
                for (u64 i = 0; i < N_; i++) dst[i] = detail::add_mod(dst[i], pfh::mulmod(r[i], p_mod_q_[J], q), q);
                put_ct_prefix(o, key_stream, key_id, true, k_, key_words);
                const uint8_t *w = reinterpret_cast<const uint8_t *>(kw.data());
                o.insert(o.end(), w, w + key_words * 8);
                if (seeded) {
                    detail::put_seal_header(o, SEAL_SEED_INFO);
                    o.push_back(1); // prng_type::blake2xb
                    o.insert(o.end(), seed, seed + 64);
                }
            }
        }
        u64 total = 16 + 32 + 8;
        for (const auto &s : slot) total += 8 + s.size();
        std::vector<uint8_t> out;
        out.reserve(total);
        detail::put_seal_header(out, total);
        for (u64 v : key_id) detail::put64(out, v);
        detail::put64(out, N_);
        for (const auto &s : slot) {
            detail::put64(out, s.empty() ? 0 : L_);
            out.insert(out.end(), s.begin(), s.end());
        }
        return out;
    }

    // Decryptor::decrypt + BatchEncoder::decode of one result stream (compr_mode none, coefficient form, any
    // level 1..L): the N slot values; *noise_budget = Decryptor::invariant_noise_budget
    std::vector<u64> decryptSlots(const uint8_t *p, size_t len, int *noise_budget = nullptr) const {
        need_keys();
        if (len < SEAL_CT_PREFIX || p[0] != 0x5E || p[1] != 0xA1 || p[5] != 0) throw std::invalid_argument("not an uncompressed SEAL stream");
        u64 total, size, n, limbs, words;
        memcpy(&total, p + 8, 8);
        memcpy(&size, p + 49, 8);
        memcpy(&n, p + 57, 8);
        memcpy(&limbs, p + 65, 8);
        memcpy(&words, p + 105, 8);
        if (p[48]) throw std::invalid_argument("result ciphertext in NTT form");
        if (size != 2 || n != N_ || limbs < 1 || limbs > L_ || words != 2 * n * limbs || total != SEAL_CT_PREFIX + words * 8 || total > len)
            throw std::invalid_argument("malformed result ciphertext");
        std::array<u64, 4> id;
        memcpy(id.data(), p + 16, 32);
        if (id != parmsId((uint32_t)limbs)) throw std::invalid_argument("result ciphertext of another parameter set (parms_id)");
        const uint32_t l = (uint32_t)limbs;
        std::vector<u64> ct(words);
        memcpy(ct.data(), p + SEAL_CT_PREFIX, words * 8);
        // x = c0 + c1*s per limb
        std::vector<u64> x((size_t)l * N_);
        for (uint32_t j = 0; j < l; j++) {
            const u64 q = q_[j];
            u64 *xj = x.data() + (size_t)j * N_;
            memcpy(xj, ct.data() + (size_t)(l + j) * N_, N_ * 8);
            for (u64 i = 0; i < N_; i++)
                if (xj[i] >= q) throw std::invalid_argument("result ciphertext word out of range");
            ntt_[j].forward(xj);
            const u64 *s = sk_ntt_.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) xj[i] = pfh::mulmod(xj[i], s[i], q);
            ntt_[j].inverse(xj);
            const u64 *c0 = ct.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) {
                if (c0[i] >= q) throw std::invalid_argument("result ciphertext word out of range");
                xj[i] = detail::add_mod(xj[i], c0[i], q);
            }
        }
        // m = round(t*x/Q) mod t with x composed by mixed radix (Garner); noise = |t*x - m'*Q|
        const Level &lv = levels_[l - 1];
        std::vector<u64> plain(N_), v(l);
        int max_noise_bits = 0;
        for (u64 i = 0; i < N_; i++) {
            for (uint32_t j = 0; j < l; j++) {
                u64 u = x[(size_t)j * N_ + i];
                for (uint32_t a = 0; a < j; a++) u = pfh::mulmod(detail::sub_mod(u, v[a] % q_[j], q_[j]), lv.garner[a][j], q_[j]);
                v[j] = u;
            }
            detail::Big X{v[l - 1]};
            for (uint32_t j = l - 1; j-- > 0;) {
                pfh::big_mul_word(X, q_[j]);
                detail::big_add_word(X, v[j]);
            }
            pfh::big_mul_word(X, t_); // t*x
            detail::Big Y(X);
            detail::big_add(Y, lv.q_half);
            for (uint32_t j = 0; j < l; j++) pfh::big_div_word(Y, q_[j]); // floor((t*x + Q/2) / Q), nested
            detail::big_trim(Y);
            const u64 mq = Y[0]; // <= t
            plain[i] = mq % t_;
            if (noise_budget) {
                detail::Big P(lv.q_big);
                pfh::big_mul_word(P, mq);
                const detail::Big diff = detail::big_cmp(X, P) >= 0 ? detail::big_sub(X, P) : detail::big_sub(P, X);
                max_noise_bits = std::max(max_noise_bits, detail::big_bits(diff));
            }
        }
        if (noise_budget) *noise_budget = std::max(0, detail::big_bits(lv.q_big) - max_noise_bits - 1);
        return decode(plain);
    }

    // One result ciphertext -> the squared L2 distances of its first `count` candidates to `query`
    // (sum of the g partial sums + ||q||^2, mod t; exact while 2 * max distance < t)
    std::vector<int64_t> decryptDistances(const uint8_t *ct, size_t len, const int64_t *query, uint32_t count, int *noise_budget = nullptr) const {
        if (count > C_) throw std::invalid_argument("a result ciphertext holds at most N/g candidates");
        const std::vector<u64> slots = decryptSlots(ct, len, noise_budget);
        u64 qq = 0;
        for (uint32_t i = 0; i < d_; i++) qq = (qq + (u64)((u128)detail::signed_mod(query[i], t_) * detail::signed_mod(query[i], t_) % t_)) % t_;
        std::vector<int64_t> out(count);
        for (uint32_t u = 0; u < count; u++) {
            u64 s = qq;
            for (uint32_t j = 0; j < g_; j++) s = (s + slots[resultSlot(u, j)]) % t_;
            out[u] = (int64_t)s;
        }
        return out;
    }

    // The response of the encrypted endpoint -> the reference's packed `coarseDistanceScores` (ref: Query.cc:53-61,
    // consumer client_lib.cpp:122-156): for query i, its probed lists in order, one float per stored vector —
    // aligned with the `coarseVectorIndexes` (labels) the server returns in the clear.  result(r) gives the r-th
    // result stream; probed_sizes [nq][nprobe] and results_per_query [nq] come from the response envelope.
    template <class ResultAt>
    std::vector<float> decrypt_coarse_scores(uint64_t nq, uint32_t nprobe, const int64_t *queries /*[nq][dim]*/, const uint64_t *probed_sizes,
                                             const uint64_t *results_per_query, ResultAt result, std::vector<uint64_t> *list_sizes_per_query = nullptr,
                                             int *min_noise_budget = nullptr) const {
        std::vector<float> scores;
        if (list_sizes_per_query) list_sizes_per_query->assign(nq, 0);
        if (min_noise_budget) *min_noise_budget = 1 << 30;
        uint64_t r = 0;
        for (uint64_t i = 0; i < nq; i++) {
            const uint64_t r_begin = r;
            for (uint32_t p = 0; p < nprobe; p++) {
                uint64_t left = probed_sizes[i * nprobe + p];
                while (left) {
                    const uint32_t count = (uint32_t)std::min<uint64_t>(left, C_);
                    const std::pair<const uint8_t *, size_t> s = result(r++);
                    int budget = 0;
                    const std::vector<int64_t> dist = decryptDistances(s.first, s.second, queries + i * d_, count, min_noise_budget ? &budget : nullptr);
                    if (min_noise_budget) *min_noise_budget = std::min(*min_noise_budget, budget);
                    for (int64_t v : dist) scores.push_back((float)v);
                    if (list_sizes_per_query) (*list_sizes_per_query)[i] += count;
                    left -= count;
                }
            }
            if (results_per_query && r - r_begin != results_per_query[i]) throw std::runtime_error("response envelope does not match its result count");
        }
        return scores;
    }

    // ref: src/client/client_lib.cpp:83-120 get_coarse_scores, for the encrypted endpoint: the JSON request body
    // (ciphertexts base64, as nlohmann would dump a string) ...
    static std::string coarse_search_encrypted_request(const std::vector<uint8_t> &query_ciphertexts, const std::vector<uint64_t> &ct_offsets,
                                                       const int64_t *nearest_centroids_id, size_t nq, size_t nprobe) {
        namespace json = handlers::json;
        std::string out = "{\"queryCiphertexts\":\"";
        out += json::base64_encode(query_ciphertexts.data(), query_ciphertexts.size());
        out += "\",\"ctOffsets\":";
        json::put_vector(out, ct_offsets.data(), ct_offsets.size());
        out += ",\"nearestCentroidIndexes\":";
        json::put_matrix(out, nearest_centroids_id, nq, nprobe);
        out.push_back('}');
        return out;
    }
    // ... and the response body
    static EncryptedCoarseResponse parse_coarse_search_encrypted_response(std::string_view body) {
        namespace json = handlers::json;
        const auto resp = json::object(body);
        EncryptedCoarseResponse r;
        r.ciphertexts = json::base64_decode(json::string(json::at(resp, "resultCiphertexts")));
        r.result_offsets = json::vector<uint64_t>(json::at(resp, "resultOffsets"));
        r.result_bytes = json::number<uint64_t>(json::at(resp, "resultBytes"));
        r.results_per_query = json::vector<uint64_t>(json::at(resp, "resultsPerQuery"));
        r.coarse_vector_indexes = json::vector<int64_t>(json::at(resp, "coarseVectorIndexes"));
        r.list_sizes_per_query = json::vector<uint64_t>(json::at(resp, "listSizesPerQuery"));
        r.probed_sizes = json::matrix<uint64_t>(json::at(resp, "probedListSizes"), r.nq, r.nprobe);
        if (r.results_per_query.size() != r.nq || r.list_sizes_per_query.size() != r.nq || r.result_offsets.empty())
            throw std::runtime_error("coarsesearch-encrypted response: array lengths disagree");
        for (size_t i = 0; i + 1 < r.result_offsets.size(); i++)
            if (r.result_offsets[i] > r.ciphertexts.size() || r.result_bytes > r.ciphertexts.size() - r.result_offsets[i])
                throw std::runtime_error("coarsesearch-encrypted response: result offsets beyond the ciphertext body");
        return r;
    }
    // get_coarse_scores after the round trip: the three outputs of the reference's function, decrypted
    void get_coarse_scores(const EncryptedCoarseResponse &r, const int64_t *queries /*[nq][dim]*/, std::vector<float> &coarse_scores,
                           std::vector<int64_t> &coarse_vectors_idx, std::vector<uint64_t> &list_sizes_per_query, int *min_noise_budget = nullptr) const {
        coarse_scores = decrypt_coarse_scores(
            r.nq, (uint32_t)r.nprobe, queries, r.probed_sizes.data(), r.results_per_query.data(),
            [&](uint64_t i) {
                if (i + 1 >= r.result_offsets.size()) throw std::runtime_error("response holds fewer results than its envelope describes");
                return std::pair<const uint8_t *, size_t>(r.ciphertexts.data() + r.result_offsets[i], (size_t)r.result_bytes);
            },
            &list_sizes_per_query, min_noise_budget);
        coarse_vectors_idx = r.coarse_vector_indexes;
        if (coarse_vectors_idx.size() != coarse_scores.size() || list_sizes_per_query != r.list_sizes_per_query)
            throw std::runtime_error("response envelope disagrees with the decrypted candidates");
    }

    // ref: src/client/client_lib.cpp:49-81 — stage 1 on the client: squared L2 of every query to every centroid in
    // the reference's arithmetic (float difference, squared in double by std::pow, accumulated into a float with a
    // rounding per step), sorted ascending.  The reference's std::ranges::sort leaves the order of equal distances
    // open; here ties keep centroid order, which is also what the engine's pf_coarse_quantize returns.
    static std::vector<std::vector<DistanceIndexData>> sort_nearest_centroids(const float *precise_query, uint64_t nq, const float *centroids,
                                                                             uint64_t nlist, uint32_t dim) {
        std::vector<std::vector<DistanceIndexData>> nearest(nq);
        for (uint64_t i = 0; i < nq; i++) {
            nearest[i].reserve(nlist);
            for (uint64_t j = 0; j < nlist; j++) {
                float distance = 0.0f;
                for (uint32_t k = 0; k < dim; k++) {
                    const float diff = precise_query[i * dim + k] - centroids[j * dim + k];
                    distance = (float)((double)distance + (double)diff * (double)diff);
                }
                nearest[i].push_back({distance, (int64_t)j});
            }
            std::stable_sort(nearest[i].begin(), nearest[i].end(), [](const DistanceIndexData &x, const DistanceIndexData &y) { return x.distance < y.distance; });
        }
        return nearest;
    }

    // ref: src/client/client_lib.cpp:122-156 — unpack per query, sort ascending by distance; a query with fewer than
    // coarse_probe candidates is an error, as there
    static std::vector<std::vector<DistanceIndexData>> compute_nearest_coarse_vectors(const std::vector<float> &coarse_distance_scores,
                                                                                     const std::vector<int64_t> &coarse_vector_indexes,
                                                                                     const std::vector<uint64_t> &list_sizes_per_query,
                                                                                     uint64_t coarse_probe) {
        std::vector<std::vector<DistanceIndexData>> nearest(list_sizes_per_query.size());
        size_t cur = 0;
        for (size_t i = 0; i < list_sizes_per_query.size(); i++) {
            if (list_sizes_per_query[i] < coarse_probe) throw std::runtime_error("Number of computed coarse scores is lesser than COARSE_PROBE");
            if (cur + list_sizes_per_query[i] > coarse_distance_scores.size() || cur + list_sizes_per_query[i] > coarse_vector_indexes.size())
                throw std::runtime_error("list sizes exceed the returned scores");
            nearest[i].reserve(list_sizes_per_query[i]);
            for (uint64_t j = 0; j < list_sizes_per_query[i]; j++) nearest[i].push_back({coarse_distance_scores[cur + j], coarse_vector_indexes[cur + j]});
            cur += list_sizes_per_query[i];
        }
        for (auto &q : nearest) std::stable_sort(q.begin(), q.end(), [](const DistanceIndexData &a, const DistanceIndexData &b) { return a.distance < b.distance; });
        return nearest;
    }

    // ref: src/client/client_lib.cpp:189-209 — pair the precise scores with the ids they were asked for (the first
    // coarse_probe of the sorted coarse vectors), sort ascending
    static std::vector<std::vector<DistanceIndexData>> compute_nearest_precise_vectors(const float *precise_scores /*[nq][coarse_probe]*/,
                                                                                      const std::vector<std::vector<DistanceIndexData>> &sorted_coarse_vectors,
                                                                                      uint64_t coarse_probe) {
        std::vector<std::vector<DistanceIndexData>> nearest(sorted_coarse_vectors.size());
        for (size_t i = 0; i < nearest.size(); i++) {
            if (sorted_coarse_vectors[i].size() < coarse_probe) throw std::runtime_error("fewer coarse vectors than COARSE_PROBE");
            for (uint64_t j = 0; j < coarse_probe; j++) nearest[i].push_back({precise_scores[i * coarse_probe + j], sorted_coarse_vectors[i][j].idx});
            std::stable_sort(nearest[i].begin(), nearest[i].end(), [](const DistanceIndexData &x, const DistanceIndexData &y) { return x.distance < y.distance; });
        }
        return nearest;
    }

    // ref: src/client/client_lib.cpp:246-330 benchmark_results — recall and MRR as the reference counts them: for each
    // of the first K ground-truth neighbours, its rank k among the K returned ids counts towards recall@1/10/100 when
    // k < 1/10/100 (divided by 1/10/100 * nq); MRR takes the first ground-truth neighbour only
    struct BenchmarkResults {
        float recall_1, recall_10, recall_100, mrr_1, mrr_10, mrr_100;
    };
    static BenchmarkResults benchmark_results(const int64_t *observed /*[nq][K]*/, uint64_t nq, uint64_t K, const int32_t *ground_truth /*[nq][gt_k]*/,
                                              uint64_t gt_k) {
        if (K > gt_k) throw std::runtime_error("K greater than nearest neigbours per query in ground truth dataset");
        float mrr_1 = 0, mrr_10 = 0, mrr_100 = 0;
        int r1 = 0, r10 = 0, r100 = 0;
        for (uint64_t i = 0; i < nq; i++)
            for (uint64_t j = 0; j < K; j++)
                for (uint64_t k = 0; k < K; k++)
                    if ((int64_t)ground_truth[i * gt_k + j] == observed[i * K + k]) {
                        r1 += k < 1;
                        r10 += k < 10;
                        r100 += k < 100;
                        if (j == 0) {
                            if (k < 1) mrr_1 += 1.0f / static_cast<float>(k + 1);
                            if (k < 10) mrr_10 += 1.0f / static_cast<float>(k + 1);
                            if (k < 100) mrr_100 += 1.0f / static_cast<float>(k + 1);
                        }
                        break;
                    }
        const float n = (float)nq;
        return {(float)r1 / (1 * n), (float)r10 / (10 * n), (float)r100 / (100 * n), mrr_1 / n, mrr_10 / n, mrr_100 / n};
    }

  private:
    struct Level {                     // the data level with l primes
        detail::Big q_big, q_half;     // Q_l, floor(Q_l / 2)
        u64 q_mod_t = 0;               // Q_l mod t
        std::vector<u64> q_div_t_mod;  // floor(Q_l / t) mod q_j
        std::vector<std::vector<u64>> garner; // garner[a][j] = q_a^-1 mod q_j, a < j
    };
    Level make_level(uint32_t l) const {
        Level lv;
        lv.q_big = {1};
        for (uint32_t j = 0; j < l; j++) pfh::big_mul_word(lv.q_big, q_[j]);
        lv.q_half = lv.q_big;
        pfh::big_div_word(lv.q_half, 2);
        detail::Big qt(lv.q_big);
        lv.q_mod_t = pfh::big_div_word(qt, t_);
        for (uint32_t j = 0; j < l; j++) lv.q_div_t_mod.push_back(pfh::big_mod_word(qt, q_[j]));
        lv.garner.assign(l, std::vector<u64>(l, 0));
        for (uint32_t a = 0; a < l; a++)
            for (uint32_t j = a + 1; j < l; j++) lv.garner[a][j] = pfh::invmod(q_[a] % q_[j], q_[j]);
        return lv;
    }
    void need_keys() const {
        if (!prng_) throw std::logic_error("generateKeys first");
    }
    // util/clipnormal.h cbd(): the difference of two 21-bit Hamming weights (standard deviation 3.24)
    void sample_noise(int64_t *e) {
        for (u64 i = 0; i < N_; i++) {
            uint8_t x[6];
            prng_->generate(6, x);
            x[2] &= 0x1F;
            x[5] &= 0x1F;
            e[i] = __builtin_popcount(x[0]) + __builtin_popcount(x[1]) + __builtin_popcount(x[2]) - __builtin_popcount(x[3]) -
                   __builtin_popcount(x[4]) - __builtin_popcount(x[5]);
        }
    }
    // encrypt_zero_symmetric at the key level: [2][k][N] NTT form, c1 = a = the expansion of `ct_seed` read as NTT
    // form (what Ciphertext::expand_seed re-creates on the server), c0 = -(a*s + e)
    std::vector<u64> encrypt_zero_key_level(const uint8_t ct_seed[64]) {
        std::vector<u64> out((size_t)2 * k_ * N_);
        u64 *c0 = out.data(), *c1 = c0 + (size_t)k_ * N_;
        pfh::SealBlake2xbPrng a_prng(ct_seed);
        pfh::seal_sample_poly_uniform(a_prng, reinterpret_cast<const uint64_t *>(q_.data()), k_, N_, reinterpret_cast<uint64_t *>(c1));
        std::vector<int64_t> e(N_);
        sample_noise(e.data());
        for (uint32_t j = 0; j < k_; j++) {
            const u64 q = q_[j];
            u64 *o = c0 + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) o[i] = detail::signed_mod(e[i], q);
            ntt_[j].forward(o);
            const u64 *a = c1 + (size_t)j * N_, *s = sk_ntt_.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) {
                const u64 v = detail::add_mod(pfh::mulmod(a[i], s[i], q), o[i], q);
                o[i] = v ? q - v : 0;
            }
        }
        return out;
    }
    // sigma_elt(s) over all k primes, NTT form: X -> X^elt on the coefficients, then transform
    std::vector<u64> rotated_secret_key(u64 elt) const {
        std::vector<int8_t> r(N_, 0);
        const u64 mask = 2 * N_ - 1;
        for (u64 i = 0; i < N_; i++) {
            const u64 e = (i * elt) & mask;
            r[e & (N_ - 1)] = (int8_t)(e >= N_ ? -sk_coeff_[i] : sk_coeff_[i]);
        }
        std::vector<u64> out((size_t)k_ * N_);
        for (uint32_t j = 0; j < k_; j++) {
            u64 *o = out.data() + (size_t)j * N_;
            for (u64 i = 0; i < N_; i++) o[i] = detail::signed_mod(r[i], q_[j]);
            ntt_[j].forward(o);
        }
        return out;
    }
    // Ciphertext::save_members up to the words (SEAL_CT_PREFIX bytes)
    void put_ct_prefix(std::vector<uint8_t> &o, u64 total, const std::array<u64, 4> &id, bool is_ntt, uint32_t limbs, u64 words) const {
        detail::put_seal_header(o, total);
        for (u64 v : id) detail::put64(o, v);
        o.push_back(is_ntt ? 1 : 0);
        detail::put64(o, 2);
        detail::put64(o, N_);
        detail::put64(o, limbs);
        const double scale = 1.0;
        u64 bits;
        memcpy(&bits, &scale, 8);
        detail::put64(o, bits);
        detail::put64(o, 1); // correction_factor
        detail::put_seal_header(o, 16 + 8 + words * 8 + 0);
        detail::put64(o, words);
    }

    uint32_t d_, dpad_ = 0, dc_ = 0, R_ = 0, C_ = 0, k_ = 0, L_ = 0;
    u64 N_;
    std::vector<u64> q_;
    u64 t_;
    uint32_t m_, g_;
    std::vector<detail::NttPlan> ntt_;
    detail::NttPlan ntt_t_;
    std::vector<uint32_t> slot_to_coeff_;
    std::vector<Level> levels_;
    std::vector<u64> p_mod_q_;
    std::vector<int8_t> sk_coeff_;
    std::vector<u64> sk_ntt_;
    std::unique_ptr<pfh::SealBlake2xbPrng> prng_;
};

} // namespace prefhetch
