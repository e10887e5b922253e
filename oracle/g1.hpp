// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// bn256 G1 as halo2curves 0.6.1 defines it [UPSTREAM, not vendored: Cargo.lock:779-781]:
// y^2 = x^3 + 3 over Fq, q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
// generator (1, 2), group order r (the Fr modulus).  `G1Affine { x, y }` holds Fq Montgomery limbs
// (R = 2^256), identity = (0, 0).  Used by the commitments of create_proof / keygen
// (`params.commit_lagrange(&poly, blind)` -> `best_multiexp(scalars, bases)`, arithmetic.rs):
// call sites main.rs:76-101, benches/aes128.rs:88-107, benches/key_schedule.rs:105-124.
//
// Here: Jacobian coordinates and a plain unsigned-window bucket method — deliberately not the
// formulas (XYZZ, signed digits) of the device code.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "fr.hpp"

namespace orc {

struct Fq {
    uint64_t l[4];
    bool operator==(const Fq &o) const { return !memcmp(l, o.l, 32); }
};
static const uint64_t FQ_MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull,
                                   0xb85045b68181585dull, 0x30644e72e131a029ull};
static const Fq FQ_ZERO = {{0, 0, 0, 0}};
// R mod q, R^2 mod q, -q^-1 mod 2^64
static const Fq FQ_ONE = {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull,
                           0x0e0a77c19a07df2full}};
static const Fq FQ_R2 = {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull,
                          0x06d89f71cab8351full}};
static const uint64_t FQ_INV = 0x87d20782e4866389ull;

static inline bool fq_geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FQ_MOD[i]) return true;
        if (a[i] < FQ_MOD[i]) return false;
    }
    return true;
}
static inline void fq_sub_mod(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - FQ_MOD[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}
static inline bool fq_is_zero(const Fq &a) { return !(a.l[0] | a.l[1] | a.l[2] | a.l[3]); }
static inline Fq fq_add(const Fq &a, const Fq &b) {
    Fq o;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.l[i] + b.l[i];
        o.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (fq_geq_mod(o.l)) fq_sub_mod(o.l);
    return o;
}
static inline Fq fq_sub(const Fq &a, const Fq &b) {
    Fq o;
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        o.l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)o.l[i] + FQ_MOD[i];
            o.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return o;
}
static inline Fq fq_mul(const Fq &a, const Fq &b) {
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
        uint64_t m = t[0] * FQ_INV;
        c = (u128)m * FQ_MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * FQ_MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fq o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fq_geq_mod(o.l)) fq_sub_mod(o.l);
    return o;
}
static inline Fq fq_from_u64(uint64_t v) {
    Fq a = {{v, 0, 0, 0}};
    return fq_mul(a, FQ_R2);
}
static inline Fq fq_neg(const Fq &a) { return fq_sub(FQ_ZERO, a); }
static inline Fq fq_inv(const Fq &a) {  // a^(q-2); inv(0) = 0
    uint64_t e[4] = {FQ_MOD[0] - 2, FQ_MOD[1], FQ_MOD[2], FQ_MOD[3]};
    Fq acc = FQ_ONE;
    for (int i = 255; i >= 0; --i) {
        acc = fq_mul(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = fq_mul(acc, a);
    }
    return acc;
}

struct G1Affine {
    Fq x, y;  // identity = (0, 0)
};
struct G1 {
    Fq x, y, z;  // Jacobian: (x/z^2, y/z^3); identity z = 0
};
static inline bool g1_is_identity(const G1 &p) { return fq_is_zero(p.z); }
static inline bool g1a_is_identity(const G1Affine &p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }
static inline G1 g1_identity() { return G1{FQ_ZERO, FQ_ONE, FQ_ZERO}; }
static inline G1 g1_from_affine(const G1Affine &p) {
    if (g1a_is_identity(p)) return g1_identity();
    return G1{p.x, p.y, FQ_ONE};
}
// dbl-2009-l (a = 0)
static inline G1 g1_double(const G1 &p) {
    if (g1_is_identity(p)) return p;
    const Fq A = fq_mul(p.x, p.x), B = fq_mul(p.y, p.y), C = fq_mul(B, B);
    Fq t = fq_add(p.x, B);
    Fq D = fq_sub(fq_sub(fq_mul(t, t), A), C);
    D = fq_add(D, D);
    const Fq E = fq_add(fq_add(A, A), A), F = fq_mul(E, E);
    G1 o;
    o.x = fq_sub(F, fq_add(D, D));
    Fq C8 = fq_add(C, C);
    C8 = fq_add(C8, C8);
    C8 = fq_add(C8, C8);
    o.y = fq_sub(fq_mul(E, fq_sub(D, o.x)), C8);
    const Fq yz = fq_mul(p.y, p.z);
    o.z = fq_add(yz, yz);
    return o;
}
// add-2007-bl, general addition with the doubling / inverse cases
static inline G1 g1_add(const G1 &p, const G1 &q) {
    if (g1_is_identity(p)) return q;
    if (g1_is_identity(q)) return p;
    const Fq z1z1 = fq_mul(p.z, p.z), z2z2 = fq_mul(q.z, q.z);
    const Fq u1 = fq_mul(p.x, z2z2), u2 = fq_mul(q.x, z1z1);
    const Fq s1 = fq_mul(fq_mul(p.y, q.z), z2z2), s2 = fq_mul(fq_mul(q.y, p.z), z1z1);
    if (u1 == u2) {
        if (s1 == s2) return g1_double(p);
        return g1_identity();
    }
    const Fq h = fq_sub(u2, u1);
    Fq i = fq_add(h, h);
    i = fq_mul(i, i);
    const Fq j = fq_mul(h, i);
    Fq r = fq_sub(s2, s1);
    r = fq_add(r, r);
    const Fq v = fq_mul(u1, i);
    G1 o;
    o.x = fq_sub(fq_sub(fq_mul(r, r), j), fq_add(v, v));
    Fq s1j = fq_mul(s1, j);
    s1j = fq_add(s1j, s1j);
    o.y = fq_sub(fq_mul(r, fq_sub(v, o.x)), s1j);
    Fq zz = fq_add(p.z, q.z);
    zz = fq_sub(fq_sub(fq_mul(zz, zz), z1z1), z2z2);
    o.z = fq_mul(zz, h);
    return o;
}
static inline G1Affine g1_to_affine(const G1 &p) {
    if (g1_is_identity(p)) return G1Affine{FQ_ZERO, FQ_ZERO};
    const Fq zi = fq_inv(p.z), zi2 = fq_mul(zi, zi);
    return G1Affine{fq_mul(p.x, zi2), fq_mul(fq_mul(p.y, zi2), zi)};
}
static inline bool g1a_on_curve(const G1Affine &p) {
    if (g1a_is_identity(p)) return true;
    const Fq lhs = fq_mul(p.y, p.y), rhs = fq_add(fq_mul(fq_mul(p.x, p.x), p.x), fq_from_u64(3));
    return lhs == rhs;
}
// scalar (canonical 4 x u64, little endian) times point, double-and-add from the top bit
static inline G1 g1_mul_canonical(const G1 &p, const uint64_t k[4]) {
    G1 acc = g1_identity();
    for (int i = 255; i >= 0; --i) {
        acc = g1_double(acc);
        if ((k[i / 64] >> (i % 64)) & 1) acc = g1_add(acc, p);
    }
    return acc;
}
static inline G1Affine g1_generator() { return G1Affine{fq_from_u64(1), fq_from_u64(2)}; }

// sum_i scalars[i] * bases[i]  (best_multiexp): scalars are Fr in Montgomery form.  Unsigned
// c-bit windows, one bucket array per window, running-sum bucket reduction, Horner over windows.
static inline G1 g1_msm(const Fr *scalars, const G1Affine *bases, size_t n, int c = 12) {
    std::vector<uint64_t> canon(4 * n);
    for (size_t i = 0; i < n; ++i) fr_to_canonical(scalars[i], &canon[4 * i]);
    const int windows = (254 + c - 1) / c;
    std::vector<G1> wsum(windows);
#pragma omp parallel for schedule(dynamic)
    for (int w = 0; w < windows; ++w) {
        std::vector<G1> buckets((size_t)1 << c, g1_identity());
        for (size_t i = 0; i < n; ++i) {
            const int bit = w * c;
            uint64_t d = canon[4 * i + bit / 64] >> (bit % 64);
            if (bit % 64 + c > 64 && bit / 64 + 1 < 4) d |= canon[4 * i + bit / 64 + 1] << (64 - bit % 64);
            d &= ((uint64_t)1 << c) - 1;
            if (d) buckets[d] = g1_add(buckets[d], g1_from_affine(bases[i]));
        }
        G1 run = g1_identity(), acc = g1_identity();
        for (size_t b = buckets.size() - 1; b >= 1; --b) {
            run = g1_add(run, buckets[b]);
            acc = g1_add(acc, run);
        }
        wsum[w] = acc;
    }
    G1 total = g1_identity();
    for (int w = windows - 1; w >= 0; --w) {
        for (int i = 0; i < c; ++i) total = g1_double(total);
        total = g1_add(total, wsum[w]);
    }
    return total;
}

}  // namespace orc
