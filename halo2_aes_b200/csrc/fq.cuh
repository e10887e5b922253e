// fq.cuh — bn256 base field Fq and the G1 group law on the device (commitments).
//
// Memory format == halo2curves 0.6.1 [UPSTREAM]: `Fq` = 4 little-endian u64 limbs of x*R mod q
// (R = 2^256), `G1Affine { x, y }` = 64 bytes, identity = (0, 0); curve y^2 = x^3 + 3.
// Arithmetic mirrors fr.cuh: 8 x 32-bit limbs, the same generated even/odd carry-chain Montgomery
// multiply (gen_fr_mul.py fq -> fq_mul_ptx.inc, self-tested against big integers), kept out of
// line.  Accumulators use extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed addition 8M + 2S with no inversion, identity ZZ = 0.
#pragma once
#include <stdint.h>

namespace h2 {

struct __align__(16) fq_t {
    uint32_t v[8];
};
// q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
#define FQ_P0 0xd87cfd47u
#define FQ_P1 0x3c208c16u
#define FQ_P2 0x6871ca8du
#define FQ_P3 0x97816a91u
#define FQ_P4 0x8181585du
#define FQ_P5 0xb85045b6u
#define FQ_P6 0xe131a029u
#define FQ_P7 0x30644e72u

__device__ __forceinline__ fq_t fq_zero() {
    fq_t z;
#pragma unroll
    for (int i = 0; i < 8; ++i) z.v[i] = 0;
    return z;
}
__device__ __forceinline__ fq_t fq_one() {  // R mod q
    fq_t o = {{0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu,
               0x9a07df2fu, 0x0e0a77c1u}};
    return o;
}
__device__ __forceinline__ bool fq_is_zero(const fq_t &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i];
    return o == 0;
}
__device__ __forceinline__ bool fq_eq(const fq_t &a, const fq_t &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
// conditional subtract of q (input < 2q)
__device__ __forceinline__ void fq_reduce_once_dev(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3,
                                                   uint32_t &r4, uint32_t &r5, uint32_t &r6, uint32_t &r7) {
    uint32_t d0, d1, d2, d3, d4, d5, d6, d7, keep;
    asm("sub.cc.u32 %0, %9, 0xd87cfd47;\n\t"
        "subc.cc.u32 %1, %10, 0x3c208c16;\n\t"
        "subc.cc.u32 %2, %11, 0x6871ca8d;\n\t"
        "subc.cc.u32 %3, %12, 0x97816a91;\n\t"
        "subc.cc.u32 %4, %13, 0x8181585d;\n\t"
        "subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
        "subc.cc.u32 %6, %15, 0xe131a029;\n\t"
        "subc.cc.u32 %7, %16, 0x30644e72;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(d4), "=r"(d5), "=r"(d6), "=r"(d7), "=r"(keep)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7));
    r0 = keep ? r0 : d0; r1 = keep ? r1 : d1; r2 = keep ? r2 : d2; r3 = keep ? r3 : d3;
    r4 = keep ? r4 : d4; r5 = keep ? r5 : d5; r6 = keep ? r6 : d6; r7 = keep ? r7 : d7;
}
__device__ __forceinline__ fq_t fq_add(const fq_t &a, const fq_t &b) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    fq_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);  // a + b < 2q < 2^255
    fq_t o = {{r0, r1, r2, r3, r4, r5, r6, r7}};
    return o;
}
__device__ __forceinline__ fq_t fq_sub(const fq_t &a, const fq_t &b) {
    const uint32_t P[8] = {FQ_P0, FQ_P1, FQ_P2, FQ_P3, FQ_P4, FQ_P5, FQ_P6, FQ_P7};
    fq_t o;
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow;
        o.v[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    if (borrow) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            c += (uint64_t)o.v[i] + P[i];
            o.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return o;
}
__device__ __forceinline__ fq_t fq_neg(const fq_t &a) { return fq_is_zero(a) ? a : fq_sub(fq_zero(), a); }
__device__ __forceinline__ fq_t fq_dbl(const fq_t &a) { return fq_add(a, a); }

#ifndef FQ_MUL_LINKAGE
#define FQ_MUL_LINKAGE __noinline__
#endif
static __device__ FQ_MUL_LINKAGE fq_t fq_mul(fq_t a, fq_t b) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fq_mul_ptx.inc"
    fq_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
    fq_t o = {{r0, r1, r2, r3, r4, r5, r6, r7}};
    return o;
}
__device__ __forceinline__ fq_t fq_sqr(const fq_t &a) { return fq_mul(a, a); }

// a^(q-2); inv(0) = 0
__device__ __forceinline__ fq_t fq_inv(const fq_t &a) {
    const uint32_t E[8] = {FQ_P0 - 2u, FQ_P1, FQ_P2, FQ_P3, FQ_P4, FQ_P5, FQ_P6, FQ_P7};
    fq_t acc = fq_one();
    for (int i = 253; i >= 0; --i) {  // q < 2^254
        acc = fq_sqr(acc);
        if ((E[i >> 5] >> (i & 31)) & 1) acc = fq_mul(acc, a);
    }
    return acc;
}

// ---- G1 ---------------------------------------------------------------------------------------
struct g1a_t {
    fq_t x, y;  // identity = (0, 0)
};
struct g1x_t {
    fq_t x, y, zz, zzz;  // identity: zz == 0
};
__device__ __forceinline__ g1x_t g1x_identity() {
    g1x_t p;
    p.x = fq_zero(); p.y = fq_zero(); p.zz = fq_zero(); p.zzz = fq_zero();
    return p;
}
__device__ __forceinline__ bool g1x_is_identity(const g1x_t &p) { return fq_is_zero(p.zz); }
__device__ __forceinline__ bool g1a_is_identity(const g1a_t &p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }

// dbl-2008-s-1 (a = 0)
__device__ __forceinline__ void g1x_double(g1x_t &p) {
    if (g1x_is_identity(p)) return;
    const fq_t u = fq_dbl(p.y), v = fq_sqr(u), w = fq_mul(u, v), s = fq_mul(p.x, v);
    const fq_t xx = fq_sqr(p.x), m = fq_add(fq_dbl(xx), xx);
    const fq_t x3 = fq_sub(fq_sqr(m), fq_dbl(s));
    const fq_t y3 = fq_sub(fq_mul(m, fq_sub(s, x3)), fq_mul(w, p.y));
    p.zz = fq_mul(v, p.zz);
    p.zzz = fq_mul(w, p.zzz);
    p.x = x3;
    p.y = y3;
}
// p += (q.x, +-q.y): madd-2008-s, with the doubling / cancellation cases
__device__ __forceinline__ void g1x_madd(g1x_t &p, const g1a_t &q, bool negate) {
    if (g1a_is_identity(q)) return;
    const fq_t qy = negate ? fq_neg(q.y) : q.y;
    if (g1x_is_identity(p)) {
        p.x = q.x; p.y = qy; p.zz = fq_one(); p.zzz = fq_one();
        return;
    }
    const fq_t u2 = fq_mul(q.x, p.zz), s2 = fq_mul(qy, p.zzz);
    const fq_t pp_ = fq_sub(u2, p.x), r = fq_sub(s2, p.y);
    if (fq_is_zero(pp_)) {
        if (fq_is_zero(r)) {  // same point: double the affine one
            p.x = q.x; p.y = qy; p.zz = fq_one(); p.zzz = fq_one();
            g1x_double(p);
        } else {
            p = g1x_identity();
        }
        return;
    }
    const fq_t pp = fq_sqr(pp_), ppp = fq_mul(pp_, pp), qq = fq_mul(p.x, pp);
    const fq_t x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_dbl(qq));
    const fq_t y3 = fq_sub(fq_mul(r, fq_sub(qq, x3)), fq_mul(p.y, ppp));
    p.zz = fq_mul(p.zz, pp);
    p.zzz = fq_mul(p.zzz, ppp);
    p.x = x3;
    p.y = y3;
}
// p += q: add-2008-s
__device__ __forceinline__ void g1x_add(g1x_t &p, const g1x_t &q) {
    if (g1x_is_identity(q)) return;
    if (g1x_is_identity(p)) {
        p = q;
        return;
    }
    const fq_t u1 = fq_mul(p.x, q.zz), u2 = fq_mul(q.x, p.zz);
    const fq_t s1 = fq_mul(p.y, q.zzz), s2 = fq_mul(q.y, p.zzz);
    const fq_t pp_ = fq_sub(u2, u1), r = fq_sub(s2, s1);
    if (fq_is_zero(pp_)) {
        if (fq_is_zero(r)) g1x_double(p);
        else p = g1x_identity();
        return;
    }
    const fq_t pp = fq_sqr(pp_), ppp = fq_mul(pp_, pp), qq = fq_mul(u1, pp);
    const fq_t x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_dbl(qq));
    const fq_t y3 = fq_sub(fq_mul(r, fq_sub(qq, x3)), fq_mul(s1, ppp));
    p.zz = fq_mul(fq_mul(p.zz, q.zz), pp);
    p.zzz = fq_mul(fq_mul(p.zzz, q.zzz), ppp);
    p.x = x3;
    p.y = y3;
}
__device__ __forceinline__ g1a_t g1x_to_affine(const g1x_t &p) {
    g1a_t o;
    if (g1x_is_identity(p)) {
        o.x = fq_zero(); o.y = fq_zero();
        return o;
    }
    // 1/zzz = zz^3 ... one inversion: i = 1/(zz*zzz); 1/zz = i*zzz; 1/zzz = i*zz
    const fq_t i = fq_inv(fq_mul(p.zz, p.zzz));
    o.x = fq_mul(p.x, fq_mul(i, p.zzz));
    o.y = fq_mul(p.y, fq_mul(i, p.zz));
    return o;
}

}  // namespace h2
