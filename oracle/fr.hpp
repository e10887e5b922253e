// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// bn256 scalar field `Fr` as halo2curves 0.6.1 stores it (third-party crate, pinned in
// /root/reference/Cargo.lock:779-781, NOT vendored): 4 little-endian u64 limbs holding
// x*R mod r (Montgomery form, R = 2^256), always fully reduced.  The reference reaches it as
// `halo2curves::bn256::Fr as Fp` (src/aes128.rs:11, src/utils.rs:3).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// use anything under oracle/.  The product (halo2_aes_b200/) never links or calls this.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t l[4];
    bool operator==(const Fr &o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const Fr &o) const { return !(*this == o); }
};

// r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
static const uint64_t FR_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                   0xb85045b68181585dull, 0x30644e72e131a029ull};
// R = 2^256 mod r  (Montgomery form of 1)
static const Fr FR_ONE = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull,
                           0x0e0a77c19a07df2full}};
// R^2 mod r
static const Fr FR_R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                          0x0216d0b17f4e44a5ull}};
// -r^{-1} mod 2^64
static const uint64_t FR_INV = 0xc2e1f593efffffffull;
static const Fr FR_ZERO = {{0, 0, 0, 0}};

static inline bool geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FR_MOD[i]) return true;
        if (a[i] < FR_MOD[i]) return false;
    }
    return true;
}

static inline void sub_mod_inplace(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - FR_MOD[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}

static inline Fr fr_add(const Fr &a, const Fr &b) {
    Fr o;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.l[i] + b.l[i];
        o.l[i] = (uint64_t)c;
        c >>= 64;
    }
    // r < 2^254 so a+b < 2^255: no carry out of the top limb.
    if (geq_mod(o.l)) sub_mod_inplace(o.l);
    return o;
}

static inline Fr fr_sub(const Fr &a, const Fr &b) {
    Fr o;
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        o.l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)o.l[i] + FR_MOD[i];
            o.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return o;
}

// Montgomery product a*b*R^{-1} mod r, CIOS over 4x64-bit limbs.
static inline Fr fr_mul(const Fr &a, const Fr &b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = (u128)m * FR_MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * FR_MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(o.l)) sub_mod_inplace(o.l);
    return o;
}

// `Fp::from(u64)`: lift a small integer into Montgomery form (one multiply by R^2), as
// halo2curves does.  Call sites: src/aes128.rs:188, src/utils.rs:23,32, src/table.rs:32...
static inline Fr fr_from_u64(uint64_t v) {
    Fr a = {{v, 0, 0, 0}};
    return fr_mul(a, FR_R2);
}

// `to_bytes()` / `to_repr()`: leave Montgomery form (multiply by 1), giving the canonical
// value as 32 little-endian bytes.  Call sites: src/utils.rs:12-13,23, gf_mul_chip.rs:81.
static inline void fr_to_canonical(const Fr &a, uint64_t out[4]) {
    Fr one = {{1, 0, 0, 0}};
    Fr c = fr_mul(a, one);
    memcpy(out, c.l, 32);
}

// `Fp::from_bytes(&[u8;32])`: canonical little-endian bytes -> Fr, None if >= r
// (src/utils.rs:18).  Returns false where halo2curves returns None.
static inline bool fr_from_canonical(const uint64_t in[4], Fr *out) {
    if (geq_mod(in)) return false;
    Fr a = {{in[0], in[1], in[2], in[3]}};
    *out = fr_mul(a, FR_R2);
    return true;
}

// `Ord for Fr` in halo2curves 0.6.1 compares the canonical representations numerically
// (most significant byte first).  Used by `permuted_input_expression.sort()` and the
// `BTreeMap` in halo2_proofs v0.3.0 plonk/lookup/prover.rs.
static inline int fr_cmp(const Fr &a, const Fr &b) {
    uint64_t ca[4], cb[4];
    fr_to_canonical(a, ca);
    fr_to_canonical(b, cb);
    for (int i = 3; i >= 0; --i) {
        if (ca[i] < cb[i]) return -1;
        if (ca[i] > cb[i]) return 1;
    }
    return 0;
}

static inline bool fr_is_zero(const Fr &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

static inline Fr fr_pow(const Fr &a, const uint64_t e[4]) {
    Fr acc = FR_ONE;
    for (int i = 255; i >= 0; --i) {
        acc = fr_mul(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = fr_mul(acc, a);
    }
    return acc;
}

static inline Fr fr_inv(const Fr &a) {  // a^(r-2); inv(0) = 0
    uint64_t e[4] = {FR_MOD[0] - 2, FR_MOD[1], FR_MOD[2], FR_MOD[3]};
    return fr_pow(a, e);
}

// ff::BatchInvert semantics (used by commit_product): zeros are skipped and stay zero.
static inline void fr_batch_invert(std::vector<Fr> &v) {
    std::vector<Fr> prefix(v.size());
    Fr acc = FR_ONE;
    for (size_t i = 0; i < v.size(); ++i) {
        prefix[i] = acc;
        if (!fr_is_zero(v[i])) acc = fr_mul(acc, v[i]);
    }
    acc = fr_inv(acc);
    for (size_t i = v.size(); i-- > 0;) {
        if (fr_is_zero(v[i])) continue;
        Fr inv_i = fr_mul(acc, prefix[i]);
        acc = fr_mul(acc, v[i]);
        v[i] = inv_i;
    }
}

}  // namespace orc
