// fr.cuh — bn256 scalar field on 8x32-bit limbs, Montgomery form (R = 2^256).
//
// Memory format == halo2curves 0.6.1 `bn256::Fr` (4 little-endian u64 limbs of x*R mod r), so a
// column of these is bit-compatible with the Vec<Fr> the reference's prover consumes
// (src/aes128.rs:11 `halo2curves::bn256::Fr as Fp`).
//
// Two implementations of the multiply:
//   * a portable host+device CIOS (also what the CPU-only tests run through h2aes_dbg_fr_*);
//   * on the device, a PTX mad.lo/madc.hi carry-chain version (IMAD pipe), used by the kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#define FR_D __device__ __forceinline__
#else
#define FR_HD inline
#endif

namespace h2 {

struct __align__(16) fr_t {
    uint32_t v[8];
};

// r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
#define FR_P0 0xf0000001u
#define FR_P1 0x43e1f593u
#define FR_P2 0x79b97091u
#define FR_P3 0x2833e848u
#define FR_P4 0x8181585du
#define FR_P5 0xb85045b6u
#define FR_P6 0xe131a029u
#define FR_P7 0x30644e72u
#define FR_INV32 0xefffffffu  // -r^{-1} mod 2^32

FR_HD uint32_t fr_p(int i) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    return P[i];
}

FR_HD fr_t fr_zero() {
    fr_t z;
#pragma unroll
    for (int i = 0; i < 8; ++i) z.v[i] = 0;
    return z;
}
// R mod r = Montgomery form of 1
FR_HD fr_t fr_one() {
    fr_t o = {{0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu,
               0x9a07df2fu, 0x0e0a77c1u}};
    return o;
}
// R^2 mod r
FR_HD fr_t fr_r2() {
    fr_t o = {{0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du,
               0x7f4e44a5u, 0x0216d0b1u}};
    return o;
}

FR_HD bool fr_is_zero(const fr_t &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i];
    return o == 0;
}
FR_HD bool fr_eq(const fr_t &a, const fr_t &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i];
    return o == 0;
}

// a >= r ?
FR_HD bool fr_geq_p(const uint32_t *a) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a[i] > P[i]) return true;
        if (a[i] < P[i]) return false;
    }
    return true;
}

// conditional subtract of r (input < 2r, output < r)
FR_HD void fr_reduce_once(uint32_t *a) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t d[8];
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] - P[i] - borrow;
        d[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    if (!borrow) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = d[i];
    }
}

FR_HD fr_t fr_add(const fr_t &a, const fr_t &b) {
    fr_t o;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a.v[i] + b.v[i];
        o.v[i] = (uint32_t)c;
        c >>= 32;
    }
    fr_reduce_once(o.v);  // a+b < 2r < 2^255: no carry out
    return o;
}

FR_HD fr_t fr_sub(const fr_t &a, const fr_t &b) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    fr_t o;
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow;
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

// Portable CIOS Montgomery product a*b*R^-1 mod r.
FR_HD fr_t fr_mul_portable(const fr_t &a, const fr_t &b) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint64_t x = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
            t[j] = (uint32_t)x;
            c = x >> 32;
        }
        uint64_t x = (uint64_t)t[8] + c;
        t[8] = (uint32_t)x;
        t[9] = (uint32_t)(x >> 32);
        uint32_t m = t[0] * FR_INV32;
        x = (uint64_t)m * P[0] + t[0];
        c = x >> 32;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            x = (uint64_t)m * P[j] + t[j] + c;
            t[j - 1] = (uint32_t)x;
            c = x >> 32;
        }
        x = (uint64_t)t[8] + c;
        t[7] = (uint32_t)x;
        t[8] = t[9] + (uint32_t)(x >> 32);
    }
    fr_t o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = t[i];
    fr_reduce_once(o.v);  // result < 2r, t[8] == 0
    return o;
}

#if defined(__CUDACC__)
// conditional subtract of r with one borrow chain (device)
FR_D void fr_reduce_once_dev(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4,
                             uint32_t &r5, uint32_t &r6, uint32_t &r7) {
    uint32_t d0, d1, d2, d3, d4, d5, d6, d7, keep;
    asm("sub.cc.u32 %0, %9, 0xf0000001;\n\t"
        "subc.cc.u32 %1, %10, 0x43e1f593;\n\t"
        "subc.cc.u32 %2, %11, 0x79b97091;\n\t"
        "subc.cc.u32 %3, %12, 0x2833e848;\n\t"
        "subc.cc.u32 %4, %13, 0x8181585d;\n\t"
        "subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
        "subc.cc.u32 %6, %15, 0xe131a029;\n\t"
        "subc.cc.u32 %7, %16, 0x30644e72;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(d4), "=r"(d5), "=r"(d6), "=r"(d7), "=r"(keep)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7));
    // keep == 0xffffffff when the subtraction borrowed (value was already < r)
    r0 = keep ? r0 : d0; r1 = keep ? r1 : d1; r2 = keep ? r2 : d2; r3 = keep ? r3 : d3;
    r4 = keep ? r4 : d4; r5 = keep ? r5 : d5; r6 = keep ? r6 : d6; r7 = keep ? r7 : d7;
}
#endif

// Montgomery product.  Device: the generated even/odd carry-chain PTX (gen_fr_mul.py, ~150
// IMAD.WIDE-class instructions); host: the portable CIOS above (same results, both < r).
#if defined(__CUDACC__)
// ONE out-of-line copy per translation unit: a kernel with a dozen inlined multiplies is
// hundreds of KB of SASS and stalls on instruction fetch (ncu: stall_no_inst 36% in
// k_grand_product); as a call the whole multiply stays resident in the instruction cache.
// Operands and result travel in registers (structs of 8 x u32 by value).
#ifndef FR_MUL_LINKAGE
#define FR_MUL_LINKAGE __noinline__
#endif
static __device__ FR_MUL_LINKAGE fr_t fr_mul_call(fr_t a, fr_t b) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fr_mul_ptx.inc"
    fr_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
    fr_t o = {{r0, r1, r2, r3, r4, r5, r6, r7}};
    return o;
}
// The same product inlined at the call site: for kernels whose hot loop has a single multiply
// (saves the call's operand shuffling; measured +13 % in the NTT stages).
static __device__ __forceinline__ fr_t fr_mul_inline(fr_t a, fr_t b) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fr_mul_ptx.inc"
    fr_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
    fr_t o = {{r0, r1, r2, r3, r4, r5, r6, r7}};
    return o;
}
// Two independent products in one call: the carry chains of the two multiplies interleave, so a
// warp keeps the multiplier busy on its own (used where occupancy is bounded by shared memory).
struct fr_pair_t {
    fr_t x, y;
};
static __device__ FR_MUL_LINKAGE fr_pair_t fr_mul2_call(fr_t a, fr_t b, fr_t c, fr_t d) {
    fr_pair_t o;
    {
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fr_mul_ptx.inc"
        fr_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
        o.x = fr_t{{r0, r1, r2, r3, r4, r5, r6, r7}};
    }
    {
        const fr_t a = c, b = d;
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#include "fr_mul_ptx.inc"
        fr_reduce_once_dev(r0, r1, r2, r3, r4, r5, r6, r7);
        o.y = fr_t{{r0, r1, r2, r3, r4, r5, r6, r7}};
    }
    return o;
}
#endif

#if defined(__CUDACC__)
// x = x*b, y = y*d (device: one interleaved call; host: two portable products)
FR_HD void fr_mul2(fr_t &x, const fr_t &b, fr_t &y, const fr_t &d) {
#if defined(__CUDA_ARCH__) && !defined(H2AES_PORTABLE_MUL)
    const fr_pair_t o = fr_mul2_call(x, b, y, d);
    x = o.x;
    y = o.y;
#else
    x = fr_mul_portable(x, b);
    y = fr_mul_portable(y, d);
#endif
}
#endif

FR_HD fr_t fr_mul(const fr_t &a, const fr_t &b) {
#if defined(__CUDA_ARCH__) && !defined(H2AES_PORTABLE_MUL)
    return fr_mul_call(a, b);
#else
    return fr_mul_portable(a, b);
#endif
}

FR_HD fr_t fr_sqr(const fr_t &a) { return fr_mul(a, a); }

FR_HD fr_t fr_from_u32(uint32_t x) {
    fr_t a = fr_zero();
    a.v[0] = x;
    return fr_mul(a, fr_r2());
}

// leave Montgomery form: canonical value (what `to_repr`/`Ord` look at)
FR_HD fr_t fr_to_canonical(const fr_t &a) {
    fr_t one = fr_zero();
    one.v[0] = 1;
    return fr_mul(a, one);
}

// numeric compare of canonical values: -1, 0, 1
FR_HD int fr_cmp_canonical(const fr_t &ca, const fr_t &cb) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (ca.v[i] < cb.v[i]) return -1;
        if (ca.v[i] > cb.v[i]) return 1;
    }
    return 0;
}

// a^(r-2) (Fermat): 254 dependent squarings; kept as the cross-check of fr_inv
FR_HD fr_t fr_inv_fermat(const fr_t &a) {
    const uint32_t E[8] = {FR_P0 - 2u, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    fr_t acc = fr_one();
    for (int i = 253; i >= 0; --i) {  // r < 2^254
        acc = fr_sqr(acc);
        if ((E[i >> 5] >> (i & 31)) & 1) acc = fr_mul(acc, a);
    }
    return acc;
}

// ---- inverse by the binary extended Euclidean algorithm --------------------------------------
// Only shifts, adds and compares on 8 limbs (~25 k cheap ALU instructions) instead of a chain of
// 380 dependent Montgomery multiplies (~70 k, mostly quarter-rate IMAD.WIDE): the batch inversion
// of the per-bin tables waits on ONE inversion per CTA, so its latency is the kernel's runtime.
FR_HD void fr_shr1(uint32_t *x) {  // x >>= 1
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[7] >>= 1;
}
FR_HD void fr_half_mod(uint32_t *x) {  // x = x/2 mod r  (x < r)
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    if (x[0] & 1u) {  // x + r is even and < 2^255
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            c += (uint64_t)x[i] + P[i];
            x[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    fr_shr1(x);
}
FR_HD bool fr_raw_geq(const uint32_t *a, const uint32_t *b) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
FR_HD void fr_raw_sub(uint32_t *a, const uint32_t *b) {  // a -= b (a >= b)
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        a[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
}
FR_HD bool fr_raw_is_one(const uint32_t *a) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; ++i) o |= a[i];
    return o == 0;
}

// inverse of 0 is 0 (same convention as ff::Field::invert().unwrap_or(0)).  Input and output in
// Montgomery form: the loop inverts the raw value aR, giving (aR)^-1; one multiply by R^3 turns
// that into a^-1 R.
FR_HD fr_t fr_inv(const fr_t &a) {
    if (fr_is_zero(a)) return a;
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t u[8], v[8];
    fr_t x1 = fr_zero(), x2 = fr_zero();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u[i] = a.v[i];
        v[i] = P[i];
    }
    x1.v[0] = 1;
    // invariants: x1 * a == u, x2 * a == v (mod r); gcd(u, v) == 1; u, v odd after the shifts
    while (!fr_raw_is_one(u) && !fr_raw_is_one(v)) {
        while (!(u[0] & 1u)) {
            fr_shr1(u);
            fr_half_mod(x1.v);
        }
        while (!(v[0] & 1u)) {
            fr_shr1(v);
            fr_half_mod(x2.v);
        }
        if (fr_raw_geq(u, v)) {
            fr_raw_sub(u, v);
            x1 = fr_sub(x1, x2);
        } else {
            fr_raw_sub(v, u);
            x2 = fr_sub(x2, x1);
        }
    }
    const fr_t r3 = {{0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u,
                      0x7fcc657cu, 0x0cf8594bu}};  // 2^768 mod r
    return fr_mul(fr_raw_is_one(u) ? x1 : x2, r3);
}

#if defined(__CUDACC__)
// 32-byte vector access helpers (column elements are 32-byte aligned)
FR_D fr_t fr_load(const uint64_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
FR_D fr_t fr_load_s(const uint4 *q) {  // shared / generic
    uint4 lo = q[0], hi = q[1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
// one 256-bit load per element (LDG.E.ENL2.256 on sm_100a): half the LSU wavefronts of two
// 128-bit loads when the lanes of a warp gather from 32 different lines.  _nc: read-only data.
FR_D fr_t fr_ldg_256(const void *p) {
    fr_t r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p)
                 : "memory");
    return r;
}
FR_D fr_t fr_ldg_256_nc(const void *p) {
    fr_t r;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
          "=r"(r.v[6]), "=r"(r.v[7])
        : "l"(p));
    return r;
}
// one 256-bit store per element (STG.E.ENL2.256 on sm_100a)
FR_D void fr_store_256(uint64_t *p, const fr_t &a) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]),
                 "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
                 "r"(a.v[7])
                 : "memory");
}
// the same for the bulk output columns, which nothing on the GPU reads back soon: evict-first in
// L2 (STG.E.EF), so that the streams of A', S', Z and advice cells do not flush the per-bin tables
// and the factor codes the next kernels gather from
#ifndef H2_STREAM_HINT
#define H2_STREAM_HINT ".cs"
#endif
FR_D void fr_store_256_cs(uint64_t *p, const fr_t &a) {
    asm volatile("st.global" H2_STREAM_HINT ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                 "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
                 "r"(a.v[6]), "r"(a.v[7])
                 : "memory");
}
FR_D void store_256_cs(uint64_t *p, const uint4 &lo, const uint4 &hi) {
    asm volatile("st.global" H2_STREAM_HINT ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                 "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z),
                 "r"(hi.w)
                 : "memory");
}
FR_D void store_256(uint64_t *p, const uint4 &lo, const uint4 &hi) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x),
                 "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
#endif

}  // namespace h2
