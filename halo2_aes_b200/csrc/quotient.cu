// quotient.cu — the quotient polynomial's numerator on the extended coset: the gate and the
// permutation expressions (k_gate_perm_quotient), then the lookup expressions (k_lookup_quotient).
//
// Replaces halo2_proofs v0.3.0 [UPSTREAM] plonk/evaluation.rs
// Evaluator::evaluate_h, lookup part, and poly/domain.rs divide_by_vanishing_poly — the first
// consumer of A', S', Z in create_proof (call sites main.rs:94-101, benches/aes128.rs:100-107,
// benches/key_schedule.rs:117-124).  For every lookup, in meta.lookup order, with idx an extended
// row and r_next / r_prev = idx +- 2^(extended_k - k):
//   values = values*y + (1 - z) * l0
//   values = values*y + (z^2 - z) * l_last
//   values = values*y + (z[r_next] (a'+beta)(s'+gamma) - z (A+beta)(S+gamma)) * l_active_row
//   values = values*y + (a' - s') * l0
//   values = values*y + (a' - s')(a' - a'[r_prev]) * l_active_row
// A = theta-fold of (q*tag, q*x[, q*y[, q*z]]) and S = theta-fold of the first m table columns,
// all evaluated on the coset (src/chips/*.rs configure: the meta.lookup closures).
//
// One lane per extended row, lookups in sequence inside the thread (the y-fold is ordered);
// everything a row needs is read exactly once with 32-byte coalesced loads.  ~20 Montgomery
// multiplies against 6 x 32 B per (row, lookup): bound by integer-multiply issue.
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "internal.h"

namespace h2 {
fr_t domain_zeta();
fr_t domain_extended_omega(int extended_k);

__device__ __forceinline__ fr_t qfr_ld(const fr_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}

struct QArgs {
    Shape sh;
    int first, count;
    uint64_t en, rot;          // extended rows; 2^(extended_k - k)
    const fr_t *advice;        // [n_advice][en]
    const fr_t *table;         // [4][en]
    const fr_t *sel;           // [count][en]
    const fr_t *Ap, *Sp, *Z;   // [count][en]
    const fr_t *lagrange;      // [3][en] l0, l_last, l_active_row
    fr_t theta, beta, gamma, y;
    fr_t *values;              // [en]
};

__global__ void __launch_bounds__(256, 2) k_lookup_quotient(const QArgs a) {
    const uint64_t en = a.en;
    const int ops = a.sh.ops_per_set();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const fr_t one = fr_one();
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < en; idx += stride) {
        const uint64_t r_next = (idx + a.rot) & (en - 1), r_prev = (idx - a.rot) & (en - 1);
        const fr_t l0 = qfr_ld(a.lagrange + idx), l_last = qfr_ld(a.lagrange + en + idx),
                   l_act = qfr_ld(a.lagrange + 2 * en + idx);
        // table expressions folded with theta, for 2, 3 and 4 columns, plus gamma
        fr_t s2 = fr_add(fr_mul(qfr_ld(a.table + idx), a.theta), qfr_ld(a.table + en + idx));
        fr_t s3 = fr_add(fr_mul(s2, a.theta), qfr_ld(a.table + 2 * en + idx));
        fr_t s4 = fr_add(fr_mul(s3, a.theta), qfr_ld(a.table + 3 * en + idx));
        s2 = fr_add(s2, a.gamma);
        s3 = fr_add(s3, a.gamma);
        s4 = fr_add(s4, a.gamma);
        fr_t v = qfr_ld(a.values + idx);
#pragma unroll 1
        for (int li = 0; li < a.count; ++li) {
            const int l = a.first + li;
            const int set = l / ops;
            const uint32_t op = (uint32_t)(l % ops) + 1;  // == table.rs Tag of the lookup
            const int m = view_of_op(op) + 2;
            // input expressions: q*tag, q*x, ... folded with theta = q * fold(tag, x, ...)
            const fr_t *adv = a.advice + (uint64_t)(3 * set) * en + idx;
            fr_t in = fr_from_u32(op);
#pragma unroll 1
            for (int jx = 0; jx + 1 < m; ++jx) in = fr_add(fr_mul(in, a.theta), qfr_ld(adv + (uint64_t)jx * en));
            in = fr_mul(in, qfr_ld(a.sel + (uint64_t)li * en + idx));
            const fr_t tv = fr_mul(fr_add(in, a.beta), m == 2 ? s2 : (m == 3 ? s3 : s4));
            const fr_t *zc = a.Z + (uint64_t)li * en, *apc = a.Ap + (uint64_t)li * en;
            const fr_t z = qfr_ld(zc + idx), ap = qfr_ld(apc + idx),
                       sp = qfr_ld(a.Sp + (uint64_t)li * en + idx);
            const fr_t ams = fr_sub(ap, sp);
            v = fr_add(fr_mul(v, a.y), fr_mul(fr_sub(one, z), l0));
            v = fr_add(fr_mul(v, a.y), fr_mul(fr_sub(fr_mul(z, z), z), l_last));
            fr_t t = fr_mul(fr_mul(qfr_ld(zc + r_next), fr_add(ap, a.beta)), fr_add(sp, a.gamma));
            t = fr_sub(t, fr_mul(z, tv));
            v = fr_add(fr_mul(v, a.y), fr_mul(t, l_act));
            v = fr_add(fr_mul(v, a.y), fr_mul(ams, l0));
            t = fr_mul(ams, fr_sub(ap, qfr_ld(apc + r_prev)));
            v = fr_add(fr_mul(v, a.y), fr_mul(t, l_act));
        }
        fr_store_256(reinterpret_cast<uint64_t *>(a.values + idx), v);
    }
}

struct GPArgs {
    Shape sh;
    uint64_t en, rot;
    const fr_t *advice;    // [n_advice][en]
    const fr_t *rc;        // [en] round_constants
    const fr_t *qrc;       // [en] q_eq_rcon as a fixed column
    const fr_t *sigma;     // [n_columns][en]
    const fr_t *Zp;        // [n_sets][en]
    const fr_t *lagrange;  // [3][en]
    const fr_t *tw;        // [en] extended_omega^idx
    fr_t beta, gamma, y, beta_zeta, delta;
    fr_t *values;
};

// evaluate_h before the lookups: the "Equality RC" gate (key_schedule.rs:59-64), then the
// permutation argument's expressions (see the header comment of h2aes_gate_perm_quotient_dev)
__global__ void __launch_bounds__(256, 2) k_gate_perm_quotient(const GPArgs a) {
    const uint64_t en = a.en;
    const int n_cols = perm_n_columns(a.sh), n_sets = perm_n_sets(a.sh);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const fr_t one = fr_one();
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < en; idx += stride) {
        const uint64_t r_next = (idx + a.rot) & (en - 1);
        const uint64_t r_last = (idx - (uint64_t)UNUSABLE_ROWS * a.rot) & (en - 1);
        const fr_t l0 = qfr_ld(a.lagrange + idx), l_last = qfr_ld(a.lagrange + en + idx),
                   l_act = qfr_ld(a.lagrange + 2 * en + idx);
        fr_t v = qfr_ld(a.values + idx);
        // gate: q_eq_rcon * (words_column - round_constants)
        const fr_t words = qfr_ld(a.advice + (uint64_t)a.sh.words_column() * en + idx);
        v = fr_add(fr_mul(v, a.y), fr_mul(qfr_ld(a.qrc + idx), fr_sub(words, qfr_ld(a.rc + idx))));
        // l0 (1 - z_0);  l_last (z_last^2 - z_last);  l0 (z_i - z_{i-1}[r_last])
        v = fr_add(fr_mul(v, a.y), fr_mul(fr_sub(one, qfr_ld(a.Zp + idx)), l0));
        const fr_t zl = qfr_ld(a.Zp + (uint64_t)(n_sets - 1) * en + idx);
        v = fr_add(fr_mul(v, a.y), fr_mul(fr_sub(fr_mul(zl, zl), zl), l_last));
#pragma unroll 1
        for (int si = 1; si < n_sets; ++si) {
            const fr_t t = fr_sub(qfr_ld(a.Zp + (uint64_t)si * en + idx),
                                  qfr_ld(a.Zp + (uint64_t)(si - 1) * en + r_last));
            v = fr_add(fr_mul(v, a.y), fr_mul(t, l0));
        }
        // l_active_row (z_i[r_next] prod (v + beta sigma + gamma) - z_i prod (v + delta^p beta X + gamma))
        fr_t cur = fr_mul(a.beta_zeta, qfr_ld(a.tw + idx));  // beta * X, X = zeta * extended_omega^idx
#pragma unroll 1
        for (int si = 0; si < n_sets; ++si) {
            fr_t left = qfr_ld(a.Zp + (uint64_t)si * en + r_next), right = qfr_ld(a.Zp + (uint64_t)si * en + idx);
#pragma unroll 1
            for (int p = PERM_CHUNK * si; p < PERM_CHUNK * si + PERM_CHUNK && p < n_cols; ++p) {
                int is_fixed, index;
                perm_column(a.sh, p, &is_fixed, &index);
                const fr_t val = qfr_ld(is_fixed ? a.rc + idx : a.advice + (uint64_t)index * en + idx);
                const fr_t vg = fr_add(val, a.gamma);
                left = fr_mul(left, fr_add(vg, fr_mul(a.beta, qfr_ld(a.sigma + (uint64_t)p * en + idx))));
                right = fr_mul(right, fr_add(vg, cur));
                cur = fr_mul(cur, a.delta);
            }
            v = fr_add(fr_mul(v, a.y), fr_mul(fr_sub(left, right), l_act));
        }
        fr_store_256(reinterpret_cast<uint64_t *>(a.values + idx), v);
    }
}

// one-hot / indicator Lagrange columns: col 0 = l0, 1 = l_last (row n-6), 2 = l_blind (last 5)
__global__ void k_lagrange_seed(uint64_t n, fr_t *__restrict__ cols) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const fr_t one = fr_one(), zero = fr_zero();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += stride) {
        const uint64_t c = i / n, r = i % n;
        const bool on = c == 0 ? r == 0 : (c == 1 ? r == n - UNUSABLE_ROWS : r > n - UNUSABLE_ROWS);
        fr_store_256(reinterpret_cast<uint64_t *>(cols + i), on ? one : zero);
    }
}
// l_active_row = 1 - (l_last + l_blind), in place over the l_blind column
__global__ void k_lagrange_active(uint64_t en, fr_t *__restrict__ ext) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const fr_t one = fr_one();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < en; i += stride) {
        const fr_t v = fr_sub(one, fr_add(qfr_ld(ext + en + i), qfr_ld(ext + 2 * en + i)));
        fr_store_256(reinterpret_cast<uint64_t *>(ext + 2 * en + i), v);
    }
}

__global__ void __launch_bounds__(256)
k_divide_vanishing(uint64_t total, uint64_t period_mask, const fr_t *__restrict__ tinv,
                   fr_t *__restrict__ a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        fr_store_256(reinterpret_cast<uint64_t *>(a + i), fr_mul(qfr_ld(a + i), qfr_ld(tinv + (i & period_mask))));
}

}  // namespace h2

using namespace h2;

namespace {
int32_t qfail(h2aes_ctx *ctx, int32_t code, const char *what) {
    if (ctx) ctx->last_error = what;
    return code;
}
int32_t qcuda(h2aes_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? H2AES_ERR_NOMEM : H2AES_ERR_CUDA;
}
#define QCU(ctx, call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return qcuda(ctx, e__, #call);  \
    } while (0)

fr_t pow_u64(fr_t base, uint64_t e) {
    fr_t acc = fr_one();
    for (; e; e >>= 1) {
        if (e & 1) acc = fr_mul(acc, base);
        base = fr_mul(base, base);
    }
    return acc;
}

// keygen.rs: l0, l_last, l_blind as Lagrange columns -> coefficient form -> extended coset
int32_t ensure_lagrange(h2aes_domain *d) {
    if (d->d_lagrange) return H2AES_OK;
    h2aes_ctx *ctx = d->ctx;
    if (d->n <= (uint64_t)UNUSABLE_ROWS) return qfail(ctx, H2AES_ERR_ARG, "domain too small for the blinding rows");
    fr_t *seed = nullptr, *ext = nullptr;
    QCU(ctx, cudaMalloc((void **)&seed, 3 * d->n * sizeof(fr_t)));
    cudaError_t e = cudaMalloc((void **)&ext, 3 * d->en * sizeof(fr_t));
    if (e) {
        cudaFree(seed);
        return qcuda(ctx, e, "cudaMalloc(lagrange)");
    }
    k_lagrange_seed<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d->n, seed);
    int32_t rc = h2aes_lagrange_to_coeff_dev(d, reinterpret_cast<uint64_t *>(seed), 3);
    if (!rc) rc = h2aes_coeff_to_extended_dev(d, reinterpret_cast<uint64_t *>(seed), 3,
                                              reinterpret_cast<uint64_t *>(ext));
    if (!rc) {
        k_lagrange_active<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d->en, ext);
        ctx->launches += 2;
        e = cudaStreamSynchronize(ctx->stream);
        if (e) rc = qcuda(ctx, e, "lagrange columns");
    }
    cudaFree(seed);
    if (rc) {
        cudaFree(ext);
        return rc;
    }
    d->d_lagrange = ext;
    return H2AES_OK;
}

int32_t ensure_tinv(h2aes_domain *d) {
    if (d->d_tinv) return H2AES_OK;
    h2aes_ctx *ctx = d->ctx;
    const uint64_t period = 1ull << (d->extended_k - d->k);
    // t_evaluations[i] = (zeta * extended_omega^i)^n - 1, i < period; inverted
    std::vector<fr_t> t(period);
    const fr_t zn = pow_u64(domain_zeta(), d->n), wn = pow_u64(domain_extended_omega(d->extended_k), d->n);
    fr_t cur = zn;
    for (uint64_t i = 0; i < period; ++i) {
        t[i] = fr_inv(fr_sub(cur, fr_one()));
        cur = fr_mul(cur, wn);
    }
    fr_t *dev = nullptr;
    QCU(ctx, cudaMalloc((void **)&dev, period * sizeof(fr_t)));
    cudaError_t e = cudaMemcpyAsync(dev, t.data(), period * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream);
    if (!e) e = cudaStreamSynchronize(ctx->stream);  // t is a host temporary
    if (e) {
        cudaFree(dev);
        return qcuda(ctx, e, "vanishing inverses");
    }
    d->d_tinv = dev;
    return H2AES_OK;
}

fr_t load_host(const uint64_t *p) {
    fr_t a;
    for (int i = 0; i < 4; ++i) {
        a.v[2 * i] = (uint32_t)p[i];
        a.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
    }
    return a;
}
}  // namespace

extern "C" {

int32_t h2aes_lookup_quotient_dev(h2aes_circuit *c, h2aes_domain *d, const uint64_t *theta,
                                  const uint64_t *beta, const uint64_t *gamma, const uint64_t *y,
                                  int32_t first, int32_t count, const uint64_t *advice_ext,
                                  const uint64_t *table_ext, const uint64_t *sel_ext,
                                  const uint64_t *Ap_ext, const uint64_t *Sp_ext,
                                  const uint64_t *Z_ext, uint64_t *values_ext) {
    if (!c || !d) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (d->ctx != ctx) return qfail(ctx, H2AES_ERR_ARG, "quotient: domain and circuit of different contexts");
    if (!theta || !beta || !gamma || !y || !advice_ext || !table_ext || !sel_ext || !Ap_ext ||
        !Sp_ext || !Z_ext || !values_ext)
        return qfail(ctx, H2AES_ERR_ARG, "quotient: null pointer");
    if (d->k != c->sh.k || d->j != 5) return qfail(ctx, H2AES_ERR_ARG, "quotient: domain must be (j=5, k of the circuit)");
    if (first < 0 || count < 0 || first + count > c->sh.n_lookups())
        return qfail(ctx, H2AES_ERR_ARG, "quotient: lookup range");
    if (c->n_shards != 1) return qfail(ctx, H2AES_ERR_STATE, "quotient: unsharded circuits only");
    QCU(ctx, cudaSetDevice(ctx->device));
    if (int32_t rc = ensure_lagrange(d)) return rc;
    if (count == 0) return H2AES_OK;
    QArgs a;
    a.sh = c->sh;
    a.first = first; a.count = count;
    a.en = d->en; a.rot = 1ull << (d->extended_k - d->k);
    a.advice = reinterpret_cast<const fr_t *>(advice_ext);
    a.table = reinterpret_cast<const fr_t *>(table_ext);
    a.sel = reinterpret_cast<const fr_t *>(sel_ext);
    a.Ap = reinterpret_cast<const fr_t *>(Ap_ext);
    a.Sp = reinterpret_cast<const fr_t *>(Sp_ext);
    a.Z = reinterpret_cast<const fr_t *>(Z_ext);
    a.lagrange = d->d_lagrange;
    a.theta = load_host(theta); a.beta = load_host(beta); a.gamma = load_host(gamma); a.y = load_host(y);
    a.values = reinterpret_cast<fr_t *>(values_ext);
    const unsigned grid = (unsigned)std::min<uint64_t>((a.en + 255) / 256, (uint64_t)ctx->sm_count * 16);
    prof_begin(ctx, "k_lookup_quotient");
    k_lookup_quotient<<<grid, 256, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    QCU(ctx, cudaGetLastError());
    return H2AES_OK;
}

int32_t h2aes_gate_perm_quotient_dev(h2aes_circuit *c, h2aes_domain *d, const uint64_t *beta,
                                     const uint64_t *gamma, const uint64_t *y,
                                     const uint64_t *advice_ext, const uint64_t *rc_ext,
                                     const uint64_t *qrc_ext, const uint64_t *sigma_ext,
                                     const uint64_t *Zp_ext, uint64_t *values_ext) {
    if (!c || !d) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (d->ctx != ctx) return qfail(ctx, H2AES_ERR_ARG, "quotient: domain and circuit of different contexts");
    if (!beta || !gamma || !y || !advice_ext || !rc_ext || !qrc_ext || !sigma_ext || !Zp_ext || !values_ext)
        return qfail(ctx, H2AES_ERR_ARG, "quotient: null pointer");
    if (d->k != c->sh.k || d->j != 5) return qfail(ctx, H2AES_ERR_ARG, "quotient: domain must be (j=5, k of the circuit)");
    if (c->n_shards != 1) return qfail(ctx, H2AES_ERR_STATE, "quotient: unsharded circuits only");
    QCU(ctx, cudaSetDevice(ctx->device));
    if (int32_t rc = ensure_lagrange(d)) return rc;
    GPArgs a;
    a.sh = c->sh;
    a.en = d->en; a.rot = 1ull << (d->extended_k - d->k);
    a.advice = reinterpret_cast<const fr_t *>(advice_ext);
    a.rc = reinterpret_cast<const fr_t *>(rc_ext);
    a.qrc = reinterpret_cast<const fr_t *>(qrc_ext);
    a.sigma = reinterpret_cast<const fr_t *>(sigma_ext);
    a.Zp = reinterpret_cast<const fr_t *>(Zp_ext);
    a.lagrange = d->d_lagrange;
    a.tw = d->d_tw;
    a.beta = load_host(beta); a.gamma = load_host(gamma); a.y = load_host(y);
    a.beta_zeta = fr_mul(a.beta, domain_zeta());
    // halo2curves 0.6.1 bn256::Fr::DELTA = 7^(2^28) [UPSTREAM]
    const uint32_t dl[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu,
                            0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
    fr_t dc;
    for (int i = 0; i < 8; ++i) dc.v[i] = dl[i];
    a.delta = fr_mul(dc, fr_r2());
    a.values = reinterpret_cast<fr_t *>(values_ext);
    const unsigned grid = (unsigned)std::min<uint64_t>((a.en + 255) / 256, (uint64_t)ctx->sm_count * 16);
    prof_begin(ctx, "k_gate_perm_quotient");
    k_gate_perm_quotient<<<grid, 256, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    QCU(ctx, cudaGetLastError());
    return H2AES_OK;
}

int32_t h2aes_divide_by_vanishing_dev(h2aes_domain *d, uint64_t *ext_dev, int32_t ncols) {
    if (!d) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = d->ctx;
    if (!ext_dev || ncols < 0) return qfail(ctx, H2AES_ERR_ARG, "divide_by_vanishing: null buffer");
    QCU(ctx, cudaSetDevice(ctx->device));
    if (int32_t rc = ensure_tinv(d)) return rc;
    if (ncols == 0) return H2AES_OK;
    prof_begin(ctx, "k_divide_vanishing");
    k_divide_vanishing<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
        (uint64_t)ncols * d->en, (1ull << (d->extended_k - d->k)) - 1, d->d_tinv,
        reinterpret_cast<fr_t *>(ext_dev));
    prof_end(ctx);
    ctx->launches += 1;
    QCU(ctx, cudaGetLastError());
    return H2AES_OK;
}

}  // extern "C"
