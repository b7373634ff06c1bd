// pf_host_math.h — host-side number theory for building the engine's device tables (moduli,
// minimal primitive roots, twiddles with Shoup quotients, BatchEncoder index map, BFV scaling
// constants).  Product code: independent of oracle/.  Conventions are SEAL 4.1's
// (modulus.cpp, util/numth.cpp, util/ntt.cpp, batchencoder.cpp, context.cpp).
#pragma once
#include <stdint.h>

#include <vector>

namespace pfh {
typedef unsigned __int128 u128;
typedef unsigned long long u64;

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
inline u64 invmod(u64 a, u64 q) { return powmod(a, q - 2, q); }
inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
inline void ratio128(u64 q, u64 &r0, u64 &r1) { // floor(2^128/q)
    u128 top = (u128)1 << 64;
    r1 = (u64)(top / q);
    u128 rem = top % q;
    r0 = (u64)((rem << 64) / q);
}
inline uint32_t bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}
inline bool is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : bases) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) {
        d >>= 1;
        s++;
    }
    for (u64 a : bases) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < s; i++) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}
// numerically smallest primitive 2n-th root of unity mod q (0 if none)
inline u64 minimal_primitive_root(u64 two_n, u64 q) {
    if ((q - 1) % two_n) return 0;
    u64 root = 0;
    for (u64 g = 2; g < 2000 && !root; g++) {
        u64 r = powmod(g, (q - 1) / two_n, q);
        if (powmod(r, two_n / 2, q) == q - 1) root = r;
    }
    if (!root) return 0;
    u64 sq = mulmod(root, root, q), cur = root, best = root;
    for (u64 i = 0; i < two_n / 2; i++) {
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    return best;
}

// little-endian multiword helpers for Q = prod q_j
inline void big_mul_word(std::vector<u64> &a, u64 x) {
    u64 carry = 0;
    for (auto &w : a) {
        u128 p = (u128)w * x + carry;
        w = (u64)p;
        carry = (u64)(p >> 64);
    }
    if (carry) a.push_back(carry);
}
inline u64 big_div_word(std::vector<u64> &a, u64 d) { // a /= d, returns remainder
    u128 rem = 0;
    for (size_t i = a.size(); i-- > 0;) {
        u128 cur = (rem << 64) | a[i];
        a[i] = (u64)(cur / d);
        rem = cur % d;
    }
    return (u64)rem;
}
inline u64 big_mod_word(const std::vector<u64> &a, u64 q) {
    u128 rem = 0;
    for (size_t i = a.size(); i-- > 0;) rem = ((rem << 64) | a[i]) % q;
    return (u64)rem;
}
} // namespace pfh
