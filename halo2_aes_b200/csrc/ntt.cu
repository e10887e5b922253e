// ntt.cu — the evaluation domain on the device: number-theoretic transforms over bn256 Fr.
//
// Replaces halo2_proofs v0.3.0 [UPSTREAM] poly/domain.rs
// EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff} and
// arithmetic.rs best_fft — what keygen_pk applies to the fixed / sigma columns and create_proof to
// the advice, A', S', Z columns right after the witness path (call sites main.rs:78-101,
// benches/aes128.rs:91-107, benches/key_schedule.rs:108-124).
//
// best_fft(a, omega, K) is the plain DFT a'[j] = sum_i a[i] omega^(ij) in natural order, so any
// exact algorithm gives the same bits.  Here: a four-step transform, N = n1 * n2,
//   k_ntt_cols  for every column i2 of the n1 x n2 (row-major) view: n1-point transform over the
//               strided index i1 in shared memory, times omega^(i2*j1), stored back in place;
//   k_ntt_rows  for every row j1: n2-point transform over the contiguous index i2 in shared
//               memory, stored transposed: X[j1 + n1*j2].
// A CTA holds a 64 KiB tile (2048 elements: G columns / rows at a time, so global accesses are
// G*32 contiguous bytes), does log2(tile transform) radix-2 stages on it, two at a time in
// registers, with one Montgomery multiply per butterfly, and touches HBM once per pass: 4 x 32 B per element per transform
// against (K/2 + 1) multiplies — the kernels are bound by integer-multiply issue.
// Twiddles come from one table of extended_omega^e (e < 2^extended_k) per domain; the inverse
// transform reads it backwards.  The coset shift (zeta^i, period 3) and the 1/N of the inverse
// transforms ride on the load of the first pass / the store of the second.
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

// few multiply sites per kernel here: inlined, they save the call's register shuffling (+13 %)
#ifndef NTT_CALL_MUL
#define FR_MUL_LINKAGE __forceinline__
#endif
#include "internal.h"

namespace h2 {

#ifndef NTT_THREADS
#define NTT_THREADS 256
#endif
#ifndef NTT_OCC
#define NTT_OCC 3
#endif
constexpr int NTT_T = NTT_THREADS;  // threads per CTA
constexpr int NTT_TILE_LOG = 11;    // elements per shared-memory tile (64 KiB)
constexpr int NTT_W_LO_BITS = 10;

__device__ __forceinline__ fr_t nfr_ld(const fr_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
__device__ __forceinline__ void nfr_st(fr_t *p, const fr_t &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

struct NttArgs {
    const uint64_t *in;
    uint64_t *out;
    uint64_t in_stride, out_stride;  // elements between consecutive columns of the batch
    uint64_t in_len;                 // input elements >= in_len read as zero (padding)
    uint64_t out_len;                // output elements >= out_len are not stored (truncation)
    int K, ka, kb;                   // N = 2^K = n1 * n2, n1 = 2^ka (strided), n2 = 2^kb (contiguous)
    int lg;                          // log2 of the columns / rows a CTA holds
    int nest;                        // k_ntt_rows as the last pass of a three-pass transform: the batch
                                     // index is (column, outer row j1o < 2^nest) and output element e
                                     // of the inner transform lands at j1o + (e << nest)
    int tw_shift;                    // omega_N^e = tw[e << tw_shift]
    uint64_t tw_mask;                // table entries - 1
    int inverse;                     // read the table backwards: omega_N^-e
    int mul3;                        // multiply element x by c[x % 3] (pass A: on load; B: on store)
    const fr_t *tw;
    const fr_t *ctw;                 // compact tables: omega_m^e, e < m, at ctw + m, m = 2..2^TILE_LOG
    const fr_t *c;                   // 3 constants (device)
};

__device__ __forceinline__ fr_t ntt_twiddle(const NttArgs &a, uint64_t e) {
    uint64_t idx = e << a.tw_shift;
    if (a.inverse) idx = (a.tw_mask + 1 - idx) & a.tw_mask;
    return nfr_ld(a.tw + idx);
}

// log2(m) decimation-in-time stages on 2^lg independent m-point arrays in shared memory (array g
// at s + g*m, input in bit-reversed order, output in natural order).  Two stages at a time: a
// thread takes the 4 elements p, p+h, p+2h, p+3h through stages h and 2h in registers (4
// multiplies, 3 twiddles, one barrier per two stages); an odd stage count starts with the
// multiply-free stage h = 1 on its own.
__device__ __forceinline__ void ntt_tile_stages(fr_t *s, int lg, int ks, const NttArgs &a) {
    // omega_m^(+-e) from the m-entry table of this transform size (32 KiB at m = 1024: stays in L1,
    // unlike the strided walk through the domain-wide table)
    const fr_t *__restrict__ ctw = a.ctw + (1u << ks);
    const uint32_t mmask = (1u << ks) - 1;
    auto stage_tw = [&](uint32_t e) { return nfr_ld(ctw + (a.inverse ? ((mmask + 1 - e) & mmask) : e)); };
    int lh = 0;
    if (ks & 1) {
        __syncthreads();
        const uint32_t total = 1u << (lg + ks - 1);
        for (uint32_t b = threadIdx.x; b < total; b += NTT_T) {
            const fr_t u = nfr_ld(s + 2 * b), t = nfr_ld(s + 2 * b + 1);
            nfr_st(s + 2 * b, fr_add(u, t));
            nfr_st(s + 2 * b + 1, fr_sub(u, t));
        }
        lh = 1;
    }
    if (ks < 2) {
        __syncthreads();
        return;
    }
    const uint32_t quarter_log = (uint32_t)(ks - 2);
    const uint32_t total = 1u << (lg + ks - 2);  // groups of 4 per round
    for (; lh < ks; lh += 2) {
        __syncthreads();
        const uint32_t h = 1u << lh;
        for (uint32_t b = threadIdx.x; b < total; b += NTT_T) {
            const uint32_t g = b >> quarter_log, bb = b & ((1u << quarter_log) - 1);
            const uint32_t j = bb & (h - 1), blk = bb >> lh;
            fr_t *p = s + (g << ks) + (blk << (lh + 2)) + j;
            fr_t x0 = nfr_ld(p), x1 = nfr_ld(p + h), x2 = nfr_ld(p + 2 * h), x3 = nfr_ld(p + 3 * h);
            // stage h: (x0, x1) and (x2, x3), both with omega_N^(j N / 2h)
            if (lh > 0) {
                const fr_t w = stage_tw(j << (ks - 1 - lh));
                fr_mul2(x1, w, x3, w);
            }
            fr_t t = x1;
            x1 = fr_sub(x0, t);
            x0 = fr_add(x0, t);
            t = x3;
            x3 = fr_sub(x2, t);
            x2 = fr_add(x2, t);
            // stage 2h: (x0, x2) with omega_N^(j N / 4h), (x1, x3) with omega_N^((j + h) N / 4h)
            const fr_t w3 = stage_tw((j + h) << (ks - 2 - lh));
            if (lh > 0) fr_mul2(x2, stage_tw(j << (ks - 2 - lh)), x3, w3);
            else x3 = fr_mul(x3, w3);
            nfr_st(p, fr_add(x0, x2));
            nfr_st(p + 2 * h, fr_sub(x0, x2));
            nfr_st(p + h, fr_add(x1, x3));
            nfr_st(p + 3 * h, fr_sub(x1, x3));
        }
    }
    __syncthreads();
}

// pass A: columns i2 in [c0, c0 + 2^lg) of the n1 x n2 view; transform over i1, twiddle, in place
__global__ void __launch_bounds__(NTT_T, NTT_OCC) k_ntt_cols(const NttArgs a) {
    extern __shared__ uint4 ntt_smem[];
    fr_t *s = reinterpret_cast<fr_t *>(ntt_smem);
    const int ka = a.ka, kb = a.kb, lg = a.lg;
    const uint64_t c0 = (uint64_t)blockIdx.x << lg;
    const uint64_t *__restrict__ in = a.in + (uint64_t)blockIdx.y * a.in_stride * 4;
    uint64_t *__restrict__ out = a.out + (uint64_t)blockIdx.y * a.out_stride * 4;
    const uint32_t elems = 1u << (lg + ka);
    for (uint32_t idx = threadIdx.x; idx < elems; idx += NTT_T) {
        const uint32_t g = idx & ((1u << lg) - 1), i1 = idx >> lg;
        const uint64_t i = ((uint64_t)i1 << kb) + c0 + g;
        fr_t v = fr_zero();
        if (i < a.in_len) {
            v = fr_load(in + 4 * i);
            if (a.mul3) {
                const uint32_t r = (uint32_t)(i % 3);
                if (r) v = fr_mul(v, nfr_ld(a.c + r));
            }
        }
        nfr_st(s + (g << ka) + (__brev(i1) >> (32 - ka)), v);
    }
    ntt_tile_stages(s, lg, ka, a);
    for (uint32_t idx = threadIdx.x; idx < elems; idx += NTT_T) {
        const uint32_t g = idx & ((1u << lg) - 1), j1 = idx >> lg;
        const uint64_t i2 = c0 + g;
        fr_t v = nfr_ld(s + (g << ka) + j1);
        v = fr_mul(v, ntt_twiddle(a, i2 * j1));  // i2 * j1 < n2 * n1 = N
        fr_store_256(out + 4 * (((uint64_t)j1 << kb) + i2), v);
    }
}

// pass B: rows j1 in [r0, r0 + 2^lg); transform over i2; X[j1 + n1*j2]
__global__ void __launch_bounds__(NTT_T, NTT_OCC) k_ntt_rows(const NttArgs a) {
    extern __shared__ uint4 ntt_smem[];
    fr_t *s = reinterpret_cast<fr_t *>(ntt_smem);
    const int ka = a.ka, kb = a.kb, lg = a.lg;
    const uint64_t r0 = (uint64_t)blockIdx.x << lg;
    const uint64_t *__restrict__ in = a.in + (uint64_t)blockIdx.y * a.in_stride * 4;
    const uint64_t j1o = blockIdx.y & ((1u << a.nest) - 1);
    uint64_t *__restrict__ out = a.out + (uint64_t)(blockIdx.y >> a.nest) * a.out_stride * 4;
    const uint32_t elems = 1u << (lg + kb);
    for (uint32_t idx = threadIdx.x; idx < elems; idx += NTT_T) {
        const uint32_t g = idx >> kb, i2 = idx & ((1u << kb) - 1);
        const fr_t v = fr_load(in + 4 * (((r0 + g) << kb) + i2));
        nfr_st(s + (g << kb) + (__brev(i2) >> (32 - kb)), v);
    }
    ntt_tile_stages(s, lg, kb, a);
    for (uint32_t idx = threadIdx.x; idx < elems; idx += NTT_T) {
        const uint32_t g = idx & ((1u << lg) - 1), j2 = idx >> lg;
        const uint64_t o = j1o + ((r0 + g + ((uint64_t)j2 << ka)) << a.nest);
        fr_t v = nfr_ld(s + (g << kb) + j2);
        if (a.mul3) v = fr_mul(v, nfr_ld(a.c + (uint32_t)(o % 3)));
        if (o < a.out_len) fr_store_256(out + 4 * o, v);
    }
}

// tw[e] = w^e from w_lo[t] = w^t (t < 1024) and w_hi[b] = w^(1024 b)
__global__ void __launch_bounds__(256)
k_ntt_twiddles(const fr_t *__restrict__ w_lo, const fr_t *__restrict__ w_hi, uint64_t n,
               fr_t *__restrict__ tw) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
        nfr_st(tw + e, fr_mul(nfr_ld(w_hi + (e >> NTT_W_LO_BITS)),
                              nfr_ld(w_lo + (e & ((1u << NTT_W_LO_BITS) - 1)))));
}

// compact per-size tables: ctw[m + e] = omega_m^e = tw[e * (table entries / m)], m = 2..2^max_log
__global__ void __launch_bounds__(256)
k_ntt_compact_twiddles(const fr_t *__restrict__ tw, int table_log, int max_log, fr_t *__restrict__ ctw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // heap index m + e
    if (i < 2 || i >= (2u << max_log)) return;
    const int ks = 31 - __clz(i);
    const uint32_t e = i - (1u << ks);
    nfr_st(ctw + i, nfr_ld(tw + ((uint64_t)e << (table_log - ks))));
}

}  // namespace h2

using namespace h2;

namespace h2 {
fr_t domain_zeta();
fr_t domain_extended_omega(int extended_k);
}

namespace {

int32_t dfail(h2aes_ctx *ctx, int32_t code, const char *what) {
    if (ctx) ctx->last_error = what;
    return code;
}
int32_t dcuda(h2aes_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? H2AES_ERR_NOMEM : H2AES_ERR_CUDA;
}
#define DCU(ctx, call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return dcuda(ctx, e__, #call);  \
    } while (0)

fr_t canon_to_mont(const uint32_t c[8]) {
    fr_t a;
    for (int i = 0; i < 8; ++i) a.v[i] = c[i];
    return fr_mul(a, fr_r2());
}
// halo2curves 0.6.1 bn256::Fr [UPSTREAM]: ROOT_OF_UNITY = 7^((r-1)/2^28), ZETA = 7^((r-1)/3)
fr_t root_of_unity_2_28() {
    const uint32_t root[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du,
                              0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    return canon_to_mont(root);
}
fr_t zeta() {
    const uint32_t z[8] = {0xb99c90ddu, 0x8b17ea66u, 0x8d8daaa7u, 0x5bfc4108u,
                           0x41a91758u, 0xb3c4d79du, 0u, 0u};
    return canon_to_mont(z);
}
fr_t fr_from_u64_host(uint64_t x) {
    fr_t a = fr_zero();
    a.v[0] = (uint32_t)x;
    a.v[1] = (uint32_t)(x >> 32);
    return fr_mul(a, fr_r2());
}

}  // namespace
namespace h2 {
fr_t domain_zeta() { return zeta(); }
// extended_omega = ROOT_OF_UNITY squared (S - extended_k) times
fr_t domain_extended_omega(int extended_k) {
    fr_t w = root_of_unity_2_28();
    for (int i = extended_k; i < 28; ++i) w = fr_mul(w, w);
    return w;
}
}  // namespace h2
namespace {

// EvaluationDomain::new: extended_k = smallest with 2^extended_k >= n * (j - 1)
bool domain_shape(int j, int k, int *extended_k, uint64_t *quotient_len) {
    if (j < 2 || j > 17 || k < 1 || k > 26) return false;
    const uint64_t n = 1ull << k, q = (uint64_t)(j - 1);
    int ek = k;
    while ((1ull << ek) < n * q) ++ek;
    *extended_k = ek;
    *quotient_len = n * q;
    return true;
}

// transform of 2^K-element columns: `in` (in_len elements, zero padded) -> out, through scratch
struct Plan {
    int K;
    int inverse;
    const fr_t *c_in;   // null or 3 constants applied on load
    const fr_t *c_out;  // null or 3 constants applied on store
    uint64_t in_stride, out_stride, in_len, out_len;
};

int32_t run_plan(h2aes_domain *d, const Plan &p, const uint64_t *in, uint64_t *out, int ncols) {
    h2aes_ctx *ctx = d->ctx;
    const int K = p.K;
    const uint64_t N = 1ull << K;
    // two passes while both factors fit a shared-memory tile; beyond that (K = 23, 24) a third:
    // N = n1o * M, the M-point row transforms are themselves two passes (over all rows at once)
    const int ko = K > 2 * NTT_TILE_LOG ? K - 2 * ((K + 2) / 3) : 0;  // outer strided factor 2^ko
    const int Ki = K - ko;                                           // inner transforms: 2^Ki points
    const int ka = (Ki + 1) / 2, kb = Ki - ka;
    NttArgs a;
    a.inverse = p.inverse;
    a.tw = d->d_tw;
    a.tw_mask = d->en - 1;
    a.ctw = d->d_ctw;
    a.nest = 0;
    auto smem_of = [](int lg, int k) { return (size_t)32 << (lg + k); };
    for (int c0 = 0; c0 < ncols; c0 += d->batch) {
        const int nb = std::min(d->batch, ncols - c0);
        const uint64_t *src = in + (uint64_t)c0 * p.in_stride * 4;
        uint64_t src_stride = p.in_stride, src_len = p.in_len;
        const fr_t *c_in = p.c_in;
        uint64_t *scratch = reinterpret_cast<uint64_t *>(d->d_scratch);
        unsigned rows_batch = (unsigned)nb;  // independent inner transforms
        if (ko) {
            // pass 0: the 2^ko-point transforms over the outermost (strided) index, twiddled by
            // omega_N^(i_M * j1o); rows j1o of M contiguous elements remain
            a.K = K; a.ka = ko; a.kb = Ki;
            a.tw_shift = d->extended_k - K;
            a.lg = std::min(NTT_TILE_LOG - ko, Ki);
            a.in = src; a.out = scratch;
            a.in_stride = src_stride; a.out_stride = N;
            a.in_len = src_len; a.out_len = N;
            a.mul3 = c_in != nullptr; a.c = c_in;
            prof_begin(ctx, "k_ntt_cols");
            k_ntt_cols<<<dim3((unsigned)(1u << (Ki - a.lg)), nb), NTT_T, smem_of(a.lg, ko), ctx->stream>>>(a);
            prof_end(ctx);
            ctx->launches += 1;
            src = scratch; src_stride = 1ull << Ki; src_len = 1ull << Ki; c_in = nullptr;
            rows_batch = (unsigned)nb << ko;
        }
        const int lga = std::min(NTT_TILE_LOG - ka, kb), lgb = std::min(NTT_TILE_LOG - kb, ka);
        a.K = Ki; a.ka = ka; a.kb = kb;
        a.tw_shift = d->extended_k - Ki;
        a.in = src; a.out = scratch;
        a.in_stride = src_stride; a.out_stride = 1ull << Ki;
        a.in_len = src_len; a.out_len = 1ull << Ki;
        a.lg = lga;
        a.mul3 = c_in != nullptr; a.c = c_in;
        prof_begin(ctx, "k_ntt_cols");
        k_ntt_cols<<<dim3((unsigned)(1u << (kb - lga)), rows_batch), NTT_T, smem_of(lga, ka), ctx->stream>>>(a);
        prof_end(ctx);
        a.in = scratch;
        a.out = out + (uint64_t)c0 * p.out_stride * 4;
        a.in_stride = 1ull << Ki; a.out_stride = p.out_stride;
        a.in_len = 1ull << Ki; a.out_len = p.out_len;
        a.lg = lgb;
        a.nest = ko;
        a.mul3 = p.c_out != nullptr; a.c = p.c_out;
        prof_begin(ctx, "k_ntt_rows");
        k_ntt_rows<<<dim3((unsigned)(1u << (ka - lgb)), rows_batch), NTT_T, smem_of(lgb, kb), ctx->stream>>>(a);
        prof_end(ctx);
        a.nest = 0;
        ctx->launches += 2;
    }
    DCU(ctx, cudaGetLastError());
    return H2AES_OK;
}

int32_t check(h2aes_domain *d, const void *p, int32_t ncols) {
    if (!d) return H2AES_ERR_ARG;
    if (!p || ncols < 0) return dfail(d->ctx, H2AES_ERR_ARG, "domain: null buffer or negative count");
    DCU(d->ctx, cudaSetDevice(d->ctx->device));
    return H2AES_OK;
}

}  // namespace

extern "C" {

int32_t h2aes_domain_query(int32_t j, int32_t k, int32_t *extended_k, uint64_t *quotient_len) {
    int ek;
    uint64_t q;
    if (!domain_shape(j, k, &ek, &q)) return H2AES_ERR_ARG;
    if (extended_k) *extended_k = ek;
    if (quotient_len) *quotient_len = q;
    return H2AES_OK;
}

int32_t h2aes_domain_create(h2aes_ctx *ctx, int32_t j, int32_t k, h2aes_domain **out) {
    if (!ctx || !out) return H2AES_ERR_ARG;
    *out = nullptr;
    int ek;
    uint64_t q;
    if (!domain_shape(j, k, &ek, &q)) return dfail(ctx, H2AES_ERR_ARG, "domain: bad j or k");
    // both sub-transforms must fit the shared-memory tile
    if (k < 2 || ek > 24) return dfail(ctx, H2AES_ERR_ARG, "domain: 2 <= k, extended_k <= 24");
    DCU(ctx, cudaSetDevice(ctx->device));
    h2aes_domain *d = new (std::nothrow) h2aes_domain;
    if (!d) return H2AES_ERR_NOMEM;
    d->ctx = ctx; d->j = j; d->k = k; d->extended_k = ek;
    d->n = 1ull << k; d->en = 1ull << ek; d->quotient_len = q;
    d->batch = ek >= 22 ? 4 : 8;
    cudaError_t e;
    fr_t *d_lo = nullptr, *d_hi = nullptr;
    const uint64_t n_hi = (d->en >> NTT_W_LO_BITS) + 1;
    if ((e = cudaMalloc((void **)&d->d_tw, d->en * sizeof(fr_t))) ||
        (e = cudaMalloc((void **)&d->d_consts, 12 * sizeof(fr_t))) ||
        (e = cudaMalloc((void **)&d->d_ctw, sizeof(fr_t) * (2u << NTT_TILE_LOG))) ||
        (e = cudaMalloc((void **)&d->d_scratch, (size_t)d->batch * d->en * sizeof(fr_t))) ||
        (e = cudaMalloc((void **)&d_lo, sizeof(fr_t) << NTT_W_LO_BITS)) ||
        (e = cudaMalloc((void **)&d_hi, n_hi * sizeof(fr_t)))) {
        cudaFree(d_lo); cudaFree(d_hi);
        h2aes_domain_destroy(d);
        return dcuda(ctx, e, "domain: cudaMalloc");
    }
    const fr_t w = domain_extended_omega(ek);
    std::vector<fr_t> lo(1u << NTT_W_LO_BITS), hi(n_hi);
    lo[0] = fr_one();
    for (size_t t = 1; t < lo.size(); ++t) lo[t] = fr_mul(lo[t - 1], w);
    const fr_t step = fr_mul(lo.back(), w);
    hi[0] = fr_one();
    for (size_t b = 1; b < hi.size(); ++b) hi[b] = fr_mul(hi[b - 1], step);
    // constants: [0..3) into the coset (1, zeta, zeta^2); [3..6) 1/n; [6..9) 1/2^extended_k times
    // (1, zeta^2, zeta) out of the coset
    const fr_t z = zeta(), z2 = fr_mul(z, z);
    const fr_t ninv = fr_inv(fr_from_u64_host(d->n)), eninv = fr_inv(fr_from_u64_host(d->en));
    fr_t c[12] = {fr_one(), z, z2, ninv, ninv, ninv, eninv, fr_mul(eninv, z2), fr_mul(eninv, z),
                  fr_one(), fr_one(), fr_one()};
    if ((e = cudaMemcpyAsync(d_lo, lo.data(), lo.size() * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream)) ||
        (e = cudaMemcpyAsync(d_hi, hi.data(), hi.size() * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream)) ||
        (e = cudaMemcpyAsync(d->d_consts, c, sizeof(c), cudaMemcpyHostToDevice, ctx->stream))) {
        cudaFree(d_lo); cudaFree(d_hi);
        h2aes_domain_destroy(d);
        return dcuda(ctx, e, "domain: upload");
    }
    prof_begin(ctx, "k_ntt_twiddles");
    k_ntt_twiddles<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_lo, d_hi, d->en, d->d_tw);
    prof_end(ctx);
    const int max_log = std::min(NTT_TILE_LOG, ek);
    k_ntt_compact_twiddles<<<((2u << max_log) + 255) / 256, 256, 0, ctx->stream>>>(d->d_tw, ek, max_log, d->d_ctw);
    ctx->launches += 2;
    e = cudaStreamSynchronize(ctx->stream);  // lo/hi/c are host temporaries
    cudaFree(d_lo); cudaFree(d_hi);
    if (!e) e = cudaFuncSetAttribute(k_ntt_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
    if (!e) e = cudaFuncSetAttribute(k_ntt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
    if (e) {
        h2aes_domain_destroy(d);
        return dcuda(ctx, e, "domain: twiddles");
    }
    *out = d;
    return H2AES_OK;
}

int32_t h2aes_domain_destroy(h2aes_domain *d) {
    if (!d) return H2AES_ERR_ARG;
    cudaSetDevice(d->ctx->device);
    cudaFree(d->d_tw);
    cudaFree(d->d_ctw);
    cudaFree(d->d_consts);
    cudaFree(d->d_scratch);
    cudaFree(d->d_lagrange);
    cudaFree(d->d_tinv);
    delete d;
    return H2AES_OK;
}

int32_t h2aes_fft_dev(h2aes_domain *d, uint64_t *cols_dev, int32_t ncols, int32_t inverse) {
    if (int32_t rc = check(d, cols_dev, ncols)) return rc;
    const Plan p{d->k, inverse != 0, nullptr, nullptr, d->n, d->n, d->n, d->n};
    return run_plan(d, p, cols_dev, cols_dev, ncols);
}

int32_t h2aes_lagrange_to_coeff_dev(h2aes_domain *d, uint64_t *cols_dev, int32_t ncols) {
    if (int32_t rc = check(d, cols_dev, ncols)) return rc;
    const Plan p{d->k, 1, nullptr, d->d_consts + 3, d->n, d->n, d->n, d->n};
    return run_plan(d, p, cols_dev, cols_dev, ncols);
}

int32_t h2aes_coeff_to_extended_dev(h2aes_domain *d, const uint64_t *coeff_dev, int32_t ncols,
                                    uint64_t *ext_dev) {
    if (int32_t rc = check(d, coeff_dev, ncols)) return rc;
    if (!ext_dev) return dfail(d->ctx, H2AES_ERR_ARG, "domain: null output");
    const Plan p{d->extended_k, 0, d->d_consts, nullptr, d->n, d->en, d->n, d->en};
    return run_plan(d, p, coeff_dev, ext_dev, ncols);
}

int32_t h2aes_extended_to_coeff_dev(h2aes_domain *d, uint64_t *ext_dev, int32_t ncols) {
    if (int32_t rc = check(d, ext_dev, ncols)) return rc;
    const Plan p{d->extended_k, 1, nullptr, d->d_consts + 6, d->en, d->en, d->en, d->en};
    int32_t rc = run_plan(d, p, ext_dev, ext_dev, ncols);
    if (rc) return rc;
    // halo2 truncates the vector to n * (j - 1) coefficients; the tail of the buffer is zeroed
    if (d->quotient_len < d->en)
        for (int c = 0; c < ncols; ++c)
            DCU(d->ctx, cudaMemsetAsync(ext_dev + ((uint64_t)c * d->en + d->quotient_len) * 4, 0,
                                        (d->en - d->quotient_len) * 32, d->ctx->stream));
    return H2AES_OK;
}

int32_t h2aes_domain_transform_host(h2aes_domain *d, int32_t op, const uint64_t *in, int32_t ncols,
                                    uint64_t *out) {
    if (int32_t rc = check(d, in, ncols)) return rc;
    if (!out || op < 0 || op > 3) return dfail(d->ctx, H2AES_ERR_ARG, "domain: bad op or null output");
    h2aes_ctx *ctx = d->ctx;
    const uint64_t n_in = op == H2AES_DOMAIN_EXTENDED_TO_COEFF ? d->en : d->n;
    const uint64_t n_buf = op >= H2AES_DOMAIN_COEFF_TO_EXTENDED ? d->en : d->n;
    const uint64_t n_out = op == H2AES_DOMAIN_COEFF_TO_EXTENDED ? d->en
                           : op == H2AES_DOMAIN_EXTENDED_TO_COEFF ? d->quotient_len : d->n;
    uint64_t *buf = nullptr, *src = nullptr;
    DCU(ctx, cudaMalloc((void **)&buf, (size_t)ncols * n_buf * 32 + 32));
    cudaError_t e = cudaSuccess;
    int32_t rc = H2AES_OK;
    if (op == H2AES_DOMAIN_COEFF_TO_EXTENDED) {
        e = cudaMalloc((void **)&src, (size_t)ncols * d->n * 32 + 32);
        if (!e) e = cudaMemcpyAsync(src, in, (size_t)ncols * d->n * 32, cudaMemcpyHostToDevice, ctx->stream);
        if (!e) rc = h2aes_coeff_to_extended_dev(d, src, ncols, buf);
    } else {
        e = cudaMemcpyAsync(buf, in, (size_t)ncols * n_in * 32, cudaMemcpyHostToDevice, ctx->stream);
        if (!e) rc = op == H2AES_DOMAIN_FFT ? h2aes_fft_dev(d, buf, ncols, 0)
                     : op == H2AES_DOMAIN_LAGRANGE_TO_COEFF ? h2aes_lagrange_to_coeff_dev(d, buf, ncols)
                                                            : h2aes_extended_to_coeff_dev(d, buf, ncols);
    }
    if (!e && !rc) {
        if (n_out == n_buf)
            e = cudaMemcpyAsync(out, buf, (size_t)ncols * n_out * 32, cudaMemcpyDeviceToHost, ctx->stream);
        else
            e = cudaMemcpy2DAsync(out, n_out * 32, buf, n_buf * 32, n_out * 32, ncols,
                                  cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (!e) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    cudaFree(src);
    if (e) return dcuda(ctx, e, "domain transform (host)");
    return rc;
}

}  // extern "C"
