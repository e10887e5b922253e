// api.cu — the extern "C" boundary declared in include/h2aes.h.
//
// Host-side glue only: argument checking with the reference's error behaviour (capacity panic of
// src/aes128.rs:160-162 -> H2AES_ERR_CAPACITY, encrypt-before-schedule_key of :170 ->
// H2AES_ERR_NOT_SCHEDULED, halo2's NotEnoughRowsAvailable -> H2AES_ERR_ROWS), device workspace
// management, and the launch order of the kernels in witness.cu / lookup.cu.  There is no CPU
// implementation of any column here: without a CUDA device every compute entry point fails.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "internal.h"
#include "permmap.h"

namespace h2 {
cudaError_t launch_build_frtab(h2aes_ctx *ctx);
cudaError_t launch_witness(h2aes_circuit *c, const uint8_t *key, const uint8_t *pts_dev,
                           uint64_t n_blocks, uint64_t *advice_dev, uint8_t *cts_dev, const RowMap &rm);
cudaError_t launch_chip_rows(h2aes_ctx *ctx, int tag, const uint8_t *x, const uint8_t *y, uint64_t n,
                             uint64_t *cells, uint8_t *out);
cudaError_t launch_witness_append(h2aes_circuit *c, const uint8_t *key, const uint8_t *pts_dev,
                                  uint64_t first_block, uint64_t n, uint64_t *advice_dev,
                                  uint8_t *cts_dev);
cudaError_t launch_fixed(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed_dev,
                         uint64_t *selectors_dev, const RowMap &rm);
cudaError_t launch_copies(h2aes_circuit *c, uint64_t n_copies, uint64_t *out);
cudaError_t launch_view_bins(h2aes_circuit *c);
cudaError_t launch_lookup_hist(h2aes_circuit *c, const RowMap &rm);
cudaError_t launch_lookup_bins(h2aes_circuit *c);
cudaError_t launch_permuted(h2aes_circuit *c, int first, int count, uint64_t *A, uint64_t *S,
                            uint64_t *Ap, uint64_t *Sp, const RowMap &rm);
cudaError_t launch_inv_tables(h2aes_circuit *c);
cudaError_t launch_z_scan(h2aes_circuit *c, int first, int count, const fr_t *carry_in,
                          uint64_t *Z, const RowMap &rm);
cudaError_t launch_z_partial(h2aes_circuit *c, uint64_t *Z, const RowMap &rm);
cudaError_t launch_z_finish(h2aes_circuit *c, const fr_t *carry_in, uint64_t *Z, const RowMap &rm);
cudaError_t launch_shard_carry(h2aes_circuit *c, const uint64_t *gathered_dev, uint32_t segs);
uint32_t lookup_bin_tiles();
cudaError_t launch_perm_tables(h2aes_circuit *c, const fr_t &omega, const fr_t &delta);
cudaError_t launch_perm_sigma(h2aes_circuit *c, uint64_t *sigma_dev);
cudaError_t launch_perm_mapping(h2aes_circuit *c, uint64_t n_blocks, const PermTemplate *tmpl_dev);
cudaError_t launch_perm_product(h2aes_circuit *c, const uint64_t *advice_dev, uint64_t *Zp);
uint64_t run_index_stride(uint64_t n);
cudaError_t launch_dbg_fr_mul(h2aes_ctx *ctx, const fr_t *a, const fr_t *b, uint64_t n, fr_t *out);

// S-box from first principles (GF(2^8) inverse + affine map); H2AES_SBOX_REFERENCE then applies
// the one deviation src/constant.rs:14 has from FIPS-197 (entry 255 is 23, not 0x16).
static void make_sbox(int mode, uint8_t sbox[256]) {
    for (int x = 0; x < 256; ++x) {
        uint8_t inv = 0;
        if (x) {
            for (int y = 1; y < 256; ++y) {
                uint8_t a = (uint8_t)x, b = (uint8_t)y, p = 0;
                for (int i = 0; i < 8; ++i) {
                    if (b & 1) p ^= a;
                    a = xtime(a);
                    b >>= 1;
                }
                if (p == 1) { inv = (uint8_t)y; break; }
            }
        }
        uint8_t s = inv;
        for (int i = 1; i <= 4; ++i) s ^= (uint8_t)((inv << i) | (inv >> (8 - i)));
        sbox[x] = (uint8_t)(s ^ 0x63);
    }
    if (mode == H2AES_SBOX_REFERENCE) sbox[255] = 23;
}

static bool shape_ok(int kind, int k, int N) {
    if (kind != KIND_AES && kind != KIND_KEYSCHED) return false;
    if (k < 1 || k > 28) return false;
    if (kind == KIND_AES && (N < 1 || N > MAX_SETS)) return false;
    return true;
}
static Shape make_shape(int kind, int k, int N) { return Shape{kind, k, kind == KIND_AES ? N : 1}; }
// halo2: the 66561-row table (and the 6 unusable rows) must fit  [UPSTREAM NotEnoughRowsAvailable]
static bool rows_ok(const Shape &sh) { return sh.k >= 17; }

static fr_t fr_from_limbs(const uint64_t *p) {
    fr_t a;
    for (int i = 0; i < 4; ++i) {
        a.v[2 * i] = (uint32_t)p[i];
        a.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
    }
    return a;
}
static void fr_to_limbs(const fr_t &a, uint64_t *p) {
    for (int i = 0; i < 4; ++i) p[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
}  // namespace h2

using namespace h2;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static int32_t cuda_fail(h2aes_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        ctx->last_error = buf;
    }
    return e == cudaErrorMemoryAllocation ? H2AES_ERR_NOMEM : H2AES_ERR_CUDA;
}
static int32_t fail(h2aes_ctx *ctx, int32_t code, const char *what) {
    if (ctx) ctx->last_error = what;
    return code;
}
#define CU(ctx, call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// allocate once: a lazy workspace builder that failed half-way is simply called again
#define DA(ptr, count)                              \
    do {                                            \
        if (!(ptr)) CU(ctx, dalloc(&(ptr), count)); \
    } while (0)

template <class T>
static cudaError_t dalloc(T **p, size_t count) {
    return cudaMalloc((void **)p, count * sizeof(T) ? count * sizeof(T) : 1);
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// host-only queries
// ---------------------------------------------------------------------------------------------
int32_t h2aes_abi_version(void) { return H2AES_ABI_VERSION; }

const char *h2aes_status_string(int32_t s) {
    switch (s) {
    case H2AES_OK: return "ok";
    case H2AES_ERR_ARG: return "bad argument";
    case H2AES_ERR_CAPACITY: return "AES calls too many. doesn't fit in the rows";
    case H2AES_ERR_ROWS: return "not enough rows available";
    case H2AES_ERR_NOT_SCHEDULED: return "Keys should be scheduled";
    case H2AES_ERR_LOOKUP: return "lookup input not in table";
    case H2AES_ERR_CUDA: return "CUDA error";
    case H2AES_ERR_NCCL: return "NCCL error";
    case H2AES_ERR_NOMEM: return "out of device memory";
    case H2AES_ERR_STATE: return "call order violated";
    case H2AES_ERR_ZERO_DENOM: return "zero denominator in the lookup grand product";
    default: return "unknown status";
    }
}

int32_t h2aes_layout_query(int32_t kind, int32_t k, int32_t N, h2aes_layout *out) {
    if (!out || !shape_ok(kind, k, N)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    memset(out, 0, sizeof *out);
    out->kind = kind;
    out->k = k;
    out->n_sets = sh.n_sets();
    out->n_advice = sh.n_advice();
    out->n_fixed = sh.n_fixed();
    out->n_selectors = sh.n_selectors();
    out->n_lookups = sh.n_lookups();
    out->n_rows = sh.n();
    out->usable_rows = sh.usable();
    out->table_rows = TABLE_ROWS;
    out->key_schedule_rows = KEY_SCHEDULE_ROWS;
    out->rows_per_block = AES_ROWS;
    out->capacity_set0 = kind == KIND_AES ? sh.cap0() : 0;
    out->capacity_set = kind == KIND_AES ? sh.cap() : 0;
    out->capacity_blocks = sh.capacity();
    out->copies_key_schedule = COPIES_KEY_SCHEDULE;
    out->copies_per_block = COPIES_PER_BLOCK;
    return H2AES_OK;
}

int32_t h2aes_block_position(int32_t k, int32_t N, uint64_t block, int32_t *set,
                             uint64_t *start_row) {
    if (!set || !start_row || !shape_ok(KIND_AES, k, N)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(KIND_AES, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    int s;
    uint64_t r;
    if (!block_position(sh, block, &s, &r)) return H2AES_ERR_CAPACITY;
    *set = s;
    *start_row = r;
    return H2AES_OK;
}

int32_t h2aes_row_op(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks, int32_t set,
                     uint64_t row) {
    if (!shape_ok(kind, k, N)) return -H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return -H2AES_ERR_ROWS;
    if (set < 0 || set >= sh.n_sets() || row >= sh.n()) return -H2AES_ERR_ARG;
    if (n_blocks > sh.capacity()) return -H2AES_ERR_CAPACITY;
    return (int32_t)row_op(sh, n_blocks, set, row);
}

int32_t h2aes_active_rows(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks, int32_t lookup,
                          uint64_t *out) {
    if (!out || !shape_ok(kind, k, N)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    if (lookup < 0 || lookup >= sh.n_lookups()) return H2AES_ERR_ARG;
    if (n_blocks > sh.capacity()) return H2AES_ERR_CAPACITY;
    const int ops = sh.ops_per_set();
    *out = active_rows(sh, n_blocks, lookup / ops, (uint32_t)(lookup % ops) + 1);
    return H2AES_OK;
}

int32_t h2aes_copy_count(int32_t kind, uint64_t n_blocks, uint64_t *out) {
    if (!out || (kind != KIND_AES && kind != KIND_KEYSCHED)) return H2AES_ERR_ARG;
    if (kind == KIND_KEYSCHED && n_blocks) return H2AES_ERR_ARG;
    *out = COPIES_KEY_SCHEDULE + COPIES_PER_BLOCK * n_blocks;
    return H2AES_OK;
}

int32_t h2aes_copy_constraints_host(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks,
                                    uint64_t first, uint64_t count, uint64_t *out) {
    if (!shape_ok(kind, k, N) || (!out && count)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    if (kind == KIND_KEYSCHED && n_blocks) return H2AES_ERR_ARG;
    if (n_blocks > sh.capacity()) return H2AES_ERR_CAPACITY;
    const uint64_t total = COPIES_KEY_SCHEDULE + COPIES_PER_BLOCK * n_blocks;
    if (first > total || count > total - first) return H2AES_ERR_ARG;
    for (uint64_t i = 0; i < count; ++i) {
        CopyPair cp;
        if (!copy_pair(sh, first + i, &cp)) return H2AES_ERR_CAPACITY;
        out[4 * i] = cp.dst.col;
        out[4 * i + 1] = cp.dst.row;
        out[4 * i + 2] = cp.src.col;
        out[4 * i + 3] = cp.src.row;
    }
    return H2AES_OK;
}

// ---------------------------------------------------------------------------------------------
// permutation argument: host-side, input-independent bookkeeping (keygen side)
// ---------------------------------------------------------------------------------------------
int32_t h2aes_perm_query(int32_t kind, int32_t k, int32_t N, int32_t *n_columns, int32_t *chunk_len,
                         int32_t *n_sets, int32_t *is_fixed, int32_t *column_index) {
    if (!shape_ok(kind, k, N)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    const int nc = perm_n_columns(sh);
    if (n_columns) *n_columns = nc;
    if (chunk_len) *chunk_len = PERM_CHUNK;
    if (n_sets) *n_sets = perm_n_sets(sh);
    for (int p = 0; p < nc; ++p) {
        int f, idx;
        perm_column(sh, p, &f, &idx);
        if (is_fixed) is_fixed[p] = f;
        if (column_index) column_index[p] = idx;
    }
    return H2AES_OK;
}

int32_t h2aes_perm_mapping_host(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks,
                                uint32_t *mapping_out) {
    if (!mapping_out || !shape_ok(kind, k, N)) return H2AES_ERR_ARG;
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return H2AES_ERR_ROWS;
    if (kind == KIND_KEYSCHED && n_blocks) return H2AES_ERR_ARG;
    if (n_blocks > sh.capacity()) return H2AES_ERR_CAPACITY;
    // closed form (permmap.h): (position, row) pairs like halo2's Vec<Vec<(usize,usize)>>
    PermTemplate t;
    perm_template_build(sh, &t);
    const uint64_t n = sh.n();
    const uint32_t nc = (uint32_t)perm_n_columns(sh);
    for (uint32_t p = 0; p < nc; ++p)
        for (uint64_t row = 0; row < n; ++row) {
            uint32_t op;
            uint64_t orow;
            perm_successor(sh, n_blocks, t, p, row, &op, &orow);
            mapping_out[2 * (p * n + row)] = op;
            mapping_out[2 * (p * n + row) + 1] = (uint32_t)orow;
        }
    return H2AES_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
int32_t h2aes_ctx_create(int32_t device, int32_t sbox_mode, h2aes_ctx **out) {
    if (!out || (sbox_mode != H2AES_SBOX_REFERENCE && sbox_mode != H2AES_SBOX_FIPS197))
        return H2AES_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return H2AES_ERR_CUDA;
    if (device < 0 || device >= count) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = new (std::nothrow) h2aes_ctx();
    if (!ctx) return H2AES_ERR_NOMEM;
    ctx->device = device;
    ctx->sbox_mode = sbox_mode;
    make_sbox(sbox_mode, ctx->h_sbox);
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) !=
            cudaSuccess ||
        (e = dalloc(&ctx->d_sbox, 256)) != cudaSuccess ||
        (e = dalloc(&ctx->d_frtab, 512)) != cudaSuccess ||
        (e = cudaMemcpy(ctx->d_sbox, ctx->h_sbox, 256, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = launch_build_frtab(ctx)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        cudaFree(ctx->d_sbox);
        cudaFree(ctx->d_frtab);
        delete ctx;
        return e == cudaErrorMemoryAllocation ? H2AES_ERR_NOMEM : H2AES_ERR_CUDA;
    }
    *out = ctx;
    return H2AES_OK;
}

int32_t h2aes_ctx_destroy(h2aes_ctx *ctx) {
    if (!ctx) return H2AES_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->comm) h2aes_comm_destroy(ctx);
    if (ctx->side) {
        cudaStreamDestroy(ctx->side);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_join);
    }
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    cudaFree(ctx->d_sbox);
    cudaFree(ctx->d_frtab);
    delete ctx;
    return H2AES_OK;
}

int32_t h2aes_ctx_set_stream(h2aes_ctx *ctx, void *stream) {
    if (!ctx) return H2AES_ERR_ARG;
    ctx->stream = (cudaStream_t)stream;
    return H2AES_OK;
}

int32_t h2aes_ctx_synchronize(h2aes_ctx *ctx) {
    if (!ctx) return H2AES_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return H2AES_OK;
}

const char *h2aes_ctx_last_error(h2aes_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t h2aes_ctx_kernel_launches(h2aes_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t h2aes_ctx_profile(h2aes_ctx *ctx, int32_t enable) {
    if (!ctx) return H2AES_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    ctx->profiling = enable != 0;
    return H2AES_OK;
}

int32_t h2aes_ctx_profile_read(h2aes_ctx *ctx, char *buf, uint64_t cap) {
    if (!ctx || !buf || !cap) return H2AES_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    struct Agg { const char *name; uint64_t n; double ms; };
    std::vector<Agg> agg;
    for (auto &r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        size_t i = 0;
        while (i < agg.size() && strcmp(agg[i].name, r.name)) ++i;
        if (i == agg.size()) agg.push_back(Agg{r.name, 0, 0.0});
        agg[i].n += 1;
        agg[i].ms += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    std::string s;
    char line[160];
    for (auto &a : agg) {
        snprintf(line, sizeof line, "%s %llu %.6f\n", a.name, (unsigned long long)a.n, a.ms);
        s += line;
    }
    if (s.size() + 1 > cap) return fail(ctx, H2AES_ERR_ARG, "profile buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
    return H2AES_OK;
}

int32_t h2aes_host_alloc(uint64_t bytes, void **out) {
    if (!out) return H2AES_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? H2AES_ERR_NOMEM : H2AES_ERR_CUDA;
    return H2AES_OK;
}
int32_t h2aes_host_free(void *p) {
    if (!p) return H2AES_ERR_ARG;
    return cudaFreeHost(p) == cudaSuccess ? H2AES_OK : H2AES_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// circuit
// ---------------------------------------------------------------------------------------------
static void circuit_free(h2aes_circuit *c) {
    void *ptrs[] = {c->d_rk,       c->d_key,      c->d_states,   c->d_trace,    c->d_chal,
                    c->d_val,      c->d_canon,    c->d_dval,     c->d_rank,     c->d_ts,
                    c->d_sorted,   c->d_slot,     c->d_head,     c->d_bhist,    c->d_bcur,     c->d_bkt,
                    c->d_bkt_b,    c->d_splval,   c->d_splidx,
                    c->d_viewmeta, c->d_ib,       c->d_ig,       c->d_nac,      c->d_cnt,
                    c->d_ibg,      c->d_ibp,      c->d_tracek,   c->d_codek,    c->d_codep,    c->d_glist,    c->d_gbase,    c->d_gcount,
                    c->d_canon_b,  c->d_active,   c->d_lbsum,    c->d_jidx,     c->d_lidx,
                    c->d_ftab,     c->d_pmap,     c->d_ptmpl,    c->d_wlo,
                    c->d_whi,      c->d_dpow,     c->d_bdp,      c->d_pnum,     c->d_ppre,
                    c->d_omega_tab,     c->d_pmask,     c->d_zmask,    c->d_znonconst,
                    c->d_offz,     c->d_nzk,      c->d_loffz,    c->d_lk,       c->d_meta,
                    c->d_ztot,    c->d_zagg,     c->d_zinc,     c->d_zflag,    c->d_zctl,
                    c->d_err,
                    c->d_partial,  c->d_carry,    c->d_gathered,    c->d_stage,    c->d_pt,       c->d_ct};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->h_err) cudaFreeHost(c->h_err);
    delete[] c->code_valid;
}

int32_t h2aes_circuit_create(h2aes_ctx *ctx, int32_t kind, int32_t k, int32_t N,
                             h2aes_circuit **out) {
    if (!ctx || !out) return H2AES_ERR_ARG;
    *out = nullptr;
    if (!shape_ok(kind, k, N)) return fail(ctx, H2AES_ERR_ARG, "bad circuit shape");
    const Shape sh = make_shape(kind, k, N);
    if (!rows_ok(sh)) return fail(ctx, H2AES_ERR_ROWS, "k < 17: the lookup table does not fit");
    if (k > 26) return fail(ctx, H2AES_ERR_ARG, "k > 26 is not supported");
    CU(ctx, cudaSetDevice(ctx->device));
    h2aes_circuit *c = new (std::nothrow) h2aes_circuit();
    if (!c) return H2AES_ERR_NOMEM;
    c->ctx = ctx;
    c->sh = sh;
    c->states_cap = sh.capacity() ? sh.capacity() : 1;
    cudaError_t e;
    if ((e = dalloc(&c->d_rk, 176)) != cudaSuccess || (e = dalloc(&c->d_key, 16)) != cudaSuccess ||
        (e = dalloc(&c->d_states, c->states_cap * ST_BYTES)) != cudaSuccess ||
        (e = dalloc(&c->d_trace, (size_t)sh.n_sets() * sh.n())) != cudaSuccess ||
        (e = dalloc(&c->d_err, 8)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&c->h_err, 8 * sizeof(uint32_t), cudaHostAllocDefault)) !=
            cudaSuccess ||
        (e = cudaMemsetAsync(c->d_err, 0, 8 * sizeof(uint32_t), ctx->stream)) != cudaSuccess) {
        circuit_free(c);
        delete c;
        return cuda_fail(ctx, e, "circuit workspace");
    }
    *out = c;
    return H2AES_OK;
}

int32_t h2aes_circuit_destroy(h2aes_circuit *c) {
    if (!c) return H2AES_ERR_ARG;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    circuit_free(c);
    delete c;
    return H2AES_OK;
}

// the rows this (shard of the) circuit holds: contiguous n/G rows, or ownership tiles of
// H2AES_SHARD_TILE_ROWS rows dealt round-robin (h2aes_circuit_set_shard_tiled)
static RowMap shard_map(const h2aes_circuit *c) {
    return row_map_of(c->sh.n(), c->shard_ts, (uint32_t)c->n_shards, (uint32_t)c->shard);
}

int32_t h2aes_circuit_set_shard_tiled(h2aes_circuit *c, int32_t shard, int32_t n_shards) {
    int32_t rc = h2aes_circuit_set_shard(c, shard, n_shards);
    if (rc != H2AES_OK || n_shards == 1) return rc;
    if (zscan_tiles(c->sh.n()) < (uint64_t)n_shards)
        return fail(c->ctx, H2AES_ERR_ARG, "tiled shards: fewer ownership tiles than shards");
    c->shard_ts = H2AES_SHARD_TILE_ROWS;
    return rc;
}

int32_t h2aes_circuit_shard_query(h2aes_circuit *c, uint64_t *rows, uint64_t *partial_words) {
    if (!c) return H2AES_ERR_ARG;
    const RowMap rm = shard_map(c);
    if (rows) *rows = rm.rows;
    if (partial_words) *partial_words = (uint64_t)c->sh.n_lookups() * rm.segments() * 4;
    return H2AES_OK;
}

int32_t h2aes_circuit_set_shard(h2aes_circuit *c, int32_t shard, int32_t n_shards) {
    if (!c) return H2AES_ERR_ARG;
    if (n_shards < 1 || n_shards > 64 || (n_shards & (n_shards - 1)) || shard < 0 ||
        shard >= n_shards)
        return fail(c->ctx, H2AES_ERR_ARG, "n_shards must be a power of two <= 64");
    c->shard = shard;
    c->n_shards = n_shards;
    c->shard_ts = 0;
    c->have_witness = c->have_bins = c->have_hist = c->have_beta_gamma = c->have_partial = false;
    return H2AES_OK;
}

int32_t h2aes_witness_dev(h2aes_circuit *c, const uint8_t *key, const uint8_t *plaintexts_dev,
                          uint64_t n_blocks, uint64_t *advice_dev, uint8_t *ciphertexts_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!key || !advice_dev || (n_blocks && !plaintexts_dev))
        return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (c->sh.kind == KIND_KEYSCHED && n_blocks)
        return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (n_blocks > c->sh.capacity())
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    CU(ctx, cudaSetDevice(ctx->device));
    const RowMap rm = shard_map(c);
    CU(ctx, launch_witness(c, key, plaintexts_dev, n_blocks, advice_dev, ciphertexts_dev, rm));
    memcpy(c->key, key, 16);
    c->n_blocks = n_blocks;
    c->have_witness = true;
    c->have_bins = c->have_hist = c->have_beta_gamma = c->have_partial = false;
    return H2AES_OK;
}

int32_t h2aes_encrypt_append_dev(h2aes_circuit *c, const uint8_t *plaintexts_dev, uint64_t n,
                                 uint64_t *advice_dev, uint8_t *ciphertexts_dev,
                                 uint64_t *first_block) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!advice_dev || (n && !plaintexts_dev)) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (c->sh.kind != KIND_AES) return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (c->n_shards != 1) return fail(ctx, H2AES_ERR_STATE, "encrypt_append: unsharded circuits only");
    if (!c->have_witness) return fail(ctx, H2AES_ERR_NOT_SCHEDULED, "Keys should be scheduled");
    if (n > c->sh.capacity() - c->n_blocks)
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    if (first_block) *first_block = c->n_blocks;
    if (!n) return H2AES_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_witness_append(c, c->key, plaintexts_dev, c->n_blocks, n, advice_dev, ciphertexts_dev));
    c->n_blocks += n;
    c->have_bins = c->have_hist = c->have_beta_gamma = c->have_partial = false;
    return H2AES_OK;
}

int32_t h2aes_chip_rows_dev(h2aes_ctx *ctx, int32_t tag, const uint8_t *x_dev, const uint8_t *y_dev,
                            uint64_t n, uint64_t *cells_dev, uint8_t *out_dev) {
    if (!ctx) return H2AES_ERR_ARG;
    if (tag < 1 || tag > 5) return fail(ctx, H2AES_ERR_ARG, "chip tag: 1 range, 2 xor, 3 sbox, 4 mul2, 5 mul3");
    if (n && (!x_dev || (tag == 2 && !y_dev))) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (!n) return H2AES_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_chip_rows(ctx, tag, x_dev, y_dev, n, cells_dev, out_dev));
    return H2AES_OK;
}

static int32_t ensure_stage(h2aes_circuit *c, uint64_t bytes) {
    if (c->stage_bytes >= bytes) return H2AES_OK;
    if (c->d_stage) {
        CU(c->ctx, cudaStreamSynchronize(c->ctx->stream));
        cudaFree(c->d_stage);
        c->d_stage = nullptr;
        c->stage_bytes = 0;
    }
    CU(c->ctx, cudaMalloc(&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return H2AES_OK;
}

int32_t h2aes_witness_host(h2aes_circuit *c, const uint8_t *key, const uint8_t *plaintexts,
                           uint64_t n_blocks, uint64_t *advice, uint8_t *ciphertexts) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!key || !advice || (n_blocks && !plaintexts)) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (c->sh.kind == KIND_KEYSCHED && n_blocks)
        return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (n_blocks > c->sh.capacity())
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    CU(ctx, cudaSetDevice(ctx->device));
    if (c->pt_cap < n_blocks || !c->d_pt) {
        if (c->d_pt) cudaFree(c->d_pt);
        if (c->d_ct) cudaFree(c->d_ct);
        c->d_pt = c->d_ct = nullptr;
        const uint64_t cap = n_blocks ? n_blocks : 1;
        CU(ctx, dalloc(&c->d_pt, cap * 16));
        CU(ctx, dalloc(&c->d_ct, cap * 16));
        c->pt_cap = cap;
    }
    const RowMap rm = shard_map(c);
    const uint64_t bytes = (uint64_t)c->sh.n_advice() * rm.rows * 32;
    int32_t rc = ensure_stage(c, bytes);
    if (rc) return rc;
    if (n_blocks)
        CU(ctx, cudaMemcpyAsync(c->d_pt, plaintexts, n_blocks * 16, cudaMemcpyHostToDevice,
                                ctx->stream));
    rc = h2aes_witness_dev(c, key, c->d_pt, n_blocks, (uint64_t *)c->d_stage, c->d_ct);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(advice, c->d_stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (ciphertexts && n_blocks)
        CU(ctx, cudaMemcpyAsync(ciphertexts, c->d_ct, n_blocks * 16, cudaMemcpyDeviceToHost,
                                ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return H2AES_OK;
}

int32_t h2aes_round_keys(h2aes_circuit *c, uint8_t *out176) {
    if (!c || !out176) return H2AES_ERR_ARG;
    if (!c->have_witness) return fail(c->ctx, H2AES_ERR_NOT_SCHEDULED, "Keys should be scheduled");
    CU(c->ctx, cudaSetDevice(c->ctx->device));
    CU(c->ctx, cudaMemcpyAsync(out176, c->d_rk, 176, cudaMemcpyDeviceToHost, c->ctx->stream));
    CU(c->ctx, cudaStreamSynchronize(c->ctx->stream));
    return H2AES_OK;
}

int32_t h2aes_fixed_columns_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed_dev,
                                uint64_t *selectors_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (c->sh.kind == KIND_KEYSCHED && n_blocks)
        return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (n_blocks > c->sh.capacity())
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    if (!fixed_dev && !selectors_dev) return H2AES_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const RowMap rm = shard_map(c);
    CU(ctx, launch_fixed(c, n_blocks, fixed_dev, selectors_dev, rm));
    return H2AES_OK;
}

int32_t h2aes_fixed_columns_host(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed,
                                 uint64_t *selectors) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    const RowMap rm = shard_map(c);
    const uint64_t fbytes = fixed ? (uint64_t)c->sh.n_fixed() * rm.rows * 32 : 0;
    const uint64_t sbytes = selectors ? (uint64_t)c->sh.n_selectors() * rm.rows * 32 : 0;
    int32_t rc = ensure_stage(c, fbytes + sbytes + 32);
    if (rc) return rc;
    uint64_t *df = fixed ? (uint64_t *)c->d_stage : nullptr;
    uint64_t *ds = selectors ? (uint64_t *)((char *)c->d_stage + fbytes) : nullptr;
    rc = h2aes_fixed_columns_dev(c, n_blocks, df, ds);
    if (rc) return rc;
    if (fixed) CU(ctx, cudaMemcpyAsync(fixed, df, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (selectors)
        CU(ctx, cudaMemcpyAsync(selectors, ds, sbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return H2AES_OK;
}

int32_t h2aes_copy_constraints_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *copies_dev) {
    if (!c || !copies_dev) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (c->sh.kind == KIND_KEYSCHED && n_blocks)
        return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (n_blocks > c->sh.capacity())
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_copies(c, COPIES_KEY_SCHEDULE + COPIES_PER_BLOCK * n_blocks, copies_dev));
    return H2AES_OK;
}

// ---------------------------------------------------------------------------------------------
// lookup helper columns
// ---------------------------------------------------------------------------------------------
static int32_t ensure_lookup_ws(h2aes_circuit *c) {
    if (c->lk_alloc) return H2AES_OK;
    h2aes_ctx *ctx = c->ctx;
    const size_t L = (size_t)c->sh.n_lookups();
    const size_t VB = (size_t)N_VIEWS * BIN_STRIDE, LB = L * BIN_STRIDE;
    c->zm = (c->sh.n() + 31) / 32 + 4;
    // block totals of L columns of n rows (>= the permutation's n_sets * n rows as one column);
    // k_z_prefix scans them in tiles of ZP_TILE_TOTALS
    c->zblocks = L * (zscan_blocks(c->sh.n()) + 1);
    c->ztiles = L * (zscan_blocks(c->sh.n()) / ZP_TILE_TOTALS + 2 * zscan_tiles(c->sh.n()) + 4);  // (tiled shards: 2 per segment)
    static_assert(H2AES_SHARD_TILE_ROWS == ZT_ROWS, "an ownership tile is a row tile of the scan");
    DA(c->d_chal, 4);
    DA(c->d_val, VB);
    DA(c->d_canon, VB);
    DA(c->d_canon_b, VB);
    DA(c->d_active, L);
    DA(c->d_lbsum, L * lookup_bin_tiles());
    c->idx_stride = run_index_stride(c->sh.n());
    DA(c->d_jidx, L * c->idx_stride);
    DA(c->d_lidx, L * c->idx_stride);
    DA(c->d_dval, VB);
    DA(c->d_rank, VB);
    DA(c->d_ts, VB);
    DA(c->d_sorted, VB);
    DA(c->d_slot, VB);
    DA(c->d_head, VB);
    DA(c->d_bkt, VB);
    DA(c->d_bkt_b, VB);
    DA(c->d_splval, (size_t)N_VIEWS * N_BUCKETS);
    DA(c->d_splidx, (size_t)N_VIEWS * N_BUCKETS);
    DA(c->d_bhist, (size_t)N_VIEWS * (N_BUCKETS + 1));
    DA(c->d_bcur, (size_t)N_VIEWS * N_BUCKETS);
    DA(c->d_viewmeta, (size_t)N_VIEWS * 8);
    DA(c->d_ib, VB);
    DA(c->d_ig, VB);
    DA(c->d_nac, VB);
    DA(c->d_ibg, VB);
    DA(c->d_ibp, VB);
    DA(c->d_ftab, 4 * VB);  // the four merged tables share one allocation (gathered by selector)
    c->d_nacp = c->d_ftab + FSEL_NACP * VB;
    c->d_ibgz = c->d_ftab + FSEL_IBGZ * VB;
    c->d_ibpz = c->d_ftab + FSEL_IBPZ * VB;
    c->d_sgz = c->d_ftab + FSEL_SGZ * VB;
    DA(c->d_tracek, (size_t)c->sh.n_sets() * c->sh.n());
    DA(c->d_codek, L * c->sh.n());
    DA(c->d_codep, L * c->sh.n());
    if (!c->d_gbase) {
        // list of the rows whose grand-product factor is not a table entry (k_permuted ->
        // k_z_general).  Such a row is a table row (< 66561 of them), or lies outside the zero
        // run of A' and is active or received a leftover table value other than the padding
        // value: at most the lookup's active rows plus one row per table bin, twice (the zero
        // run's own rows with such a leftover value).
        uint32_t base[5 * MAX_SETS + 1];
        const int ops = c->sh.ops_per_set();
        base[0] = 0;
        for (size_t l = 0; l < L; ++l) {
            uint64_t cap = active_rows(c->sh, c->sh.capacity(), (int)l / ops, (uint32_t)(l % ops) + 1) +
                           2 * (TABLE_ROWS + 1) + 64;
            if (cap > c->sh.n()) cap = c->sh.n();
            base[l + 1] = base[l] + (uint32_t)cap;
        }
        DA(c->d_glist, (size_t)base[L]);
        DA(c->d_gcount, L);
        DA(c->d_gbase, L + 1);
        CU(ctx, cudaMemcpyAsync(c->d_gbase, base, (L + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));  // `base` is on the stack
    }
    if (!c->code_valid) c->code_valid = new bool[L]();
    DA(c->d_cnt, LB);
    DA(c->d_offz, LB);
    DA(c->d_nzk, LB);
    DA(c->d_loffz, LB);
    DA(c->d_lk, LB);
    DA(c->d_meta, L);
    DA(c->d_zmask, 2 * L * c->zm);  // non-constant rows, then the rows of k_z_general
    DA(c->d_ztot, c->zblocks);
    DA(c->d_zagg, c->ztiles);
    DA(c->d_zinc, c->ztiles);
    if (!c->d_zflag) {
        DA(c->d_zflag, c->ztiles);
        CU(ctx, cudaMemsetAsync(c->d_zflag, 0, c->ztiles * 4, ctx->stream));
    }
    if (!c->d_zctl) {
        DA(c->d_zctl, 4);
        CU(ctx, cudaMemsetAsync(c->d_zctl, 0, 16, ctx->stream));
    }
    DA(c->d_znonconst, L);
    CU(ctx, cudaMemsetAsync(c->d_znonconst, 0, L * 8, ctx->stream));
    // one partial product / carry per (lookup, segment): a tiled shard holds up to
    // ceil(tiles / G) ownership tiles of every column
    DA(c->d_partial, L * (zscan_tiles(c->sh.n()) + 1));
    DA(c->d_carry, L * (zscan_tiles(c->sh.n()) + 1));
    c->lk_alloc = true;
    return H2AES_OK;
}

// Device-side error flags raised by the lookup kernels are folded into the status here.
static int32_t check_device_flags(h2aes_circuit *c) {
    h2aes_ctx *ctx = c->ctx;
    CU(ctx, cudaMemcpyAsync(c->h_err, c->d_err, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (c->h_err[0] || c->h_err[1] || c->h_err[2])
        CU(ctx, cudaMemsetAsync(c->d_err, 0, 8 * sizeof(uint32_t), ctx->stream));
    if (c->h_err[0]) return fail(ctx, H2AES_ERR_ZERO_DENOM, "zero denominator in the grand product");
    if (c->h_err[1]) return fail(ctx, H2AES_ERR_LOOKUP, "lookup input not in table");
    if (c->h_err[2]) return fail(ctx, H2AES_ERR_STATE, "internal: general-row list overflow");
    return H2AES_OK;
}

// rows with each lookup's selector enabled (closed form, layout.h) -> device, for k_lookup_bins
static int32_t upload_active(h2aes_circuit *c) {
    const int L = c->sh.n_lookups(), ops = c->sh.ops_per_set();
    uint32_t host[5 * MAX_SETS];
    for (int l = 0; l < L; ++l)
        host[l] = (uint32_t)active_rows(c->sh, c->n_blocks, l / ops, (uint32_t)(l % ops) + 1);
    CU(c->ctx, cudaMemcpyAsync(c->d_active, host, sizeof(uint32_t) * L, cudaMemcpyHostToDevice,
                               c->ctx->stream));
    return H2AES_OK;
}

static int32_t set_theta(h2aes_circuit *c, const uint64_t *theta, bool *changed) {
    *changed = !c->have_theta || memcmp(c->theta, theta, 32) != 0;
    if (*changed) {
        memcpy(c->theta, theta, 32);
        CU(c->ctx, cudaMemcpyAsync(c->d_chal, c->theta, 32, cudaMemcpyHostToDevice, c->ctx->stream));
        c->have_theta = true;
    }
    return H2AES_OK;
}

static int32_t set_beta_gamma(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma) {
    h2aes_ctx *ctx = c->ctx;
    if (c->have_beta_gamma && !memcmp(c->beta, beta, 32) && !memcmp(c->gamma, gamma, 32))
        return H2AES_OK;
    memcpy(c->beta, beta, 32);
    memcpy(c->gamma, gamma, 32);
    CU(ctx, cudaMemcpyAsync(c->d_chal + 1, c->beta, 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(c->d_chal + 2, c->gamma, 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, launch_inv_tables(c));
    c->have_beta_gamma = true;
    return H2AES_OK;
}

// (comm.cu: the per-(beta,gamma) tables on a side stream while the permuted columns are expanded)
int32_t h2aes_internal_set_beta_gamma(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma) {
    return set_beta_gamma(c, beta, gamma);
}

int32_t h2aes_lookup_permuted_dev(h2aes_circuit *c, const uint64_t *theta, int32_t first,
                                  int32_t count, uint64_t *A_dev, uint64_t *S_dev,
                                  uint64_t *Ap_dev, uint64_t *Sp_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!theta || !Ap_dev || !Sp_dev) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->sh.n_lookups())
        return fail(ctx, H2AES_ERR_ARG, "lookup range");
    if (!c->have_witness) return fail(ctx, H2AES_ERR_STATE, "lookup columns need a witness first");
    if (c->n_shards != 1)
        return fail(ctx, H2AES_ERR_STATE, "sharded circuit: use the *_shard_dev entry points");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = ensure_lookup_ws(c);
    if (rc) return rc;
    bool changed;
    if ((rc = set_theta(c, theta, &changed))) return rc;
    if (changed || !c->have_bins) {
        if ((rc = upload_active(c))) return rc;
        CU(ctx, launch_view_bins(c));
        CU(ctx, launch_lookup_hist(c, shard_map(c)));
        CU(ctx, launch_lookup_bins(c));
        c->have_bins = true;
        c->have_beta_gamma = false;
        for (int l = 0; l < c->sh.n_lookups(); ++l) c->code_valid[l] = false;
    }
    if (count) {
        CU(ctx, launch_permuted(c, first, count, A_dev, S_dev, Ap_dev, Sp_dev, shard_map(c)));
        for (int l = first; l < first + count; ++l) c->code_valid[l] = true;
    }
    return H2AES_OK;
}

int32_t h2aes_lookup_product_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                 int32_t first, int32_t count, uint64_t *Z_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!beta || !gamma || !Z_dev) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->sh.n_lookups())
        return fail(ctx, H2AES_ERR_ARG, "lookup range");
    if (c->n_shards != 1)
        return fail(ctx, H2AES_ERR_STATE, "sharded circuit: use the *_shard_dev entry points");
    if (!c->have_witness || !c->have_bins)
        return fail(ctx, H2AES_ERR_STATE, "the grand product needs the permuted columns first");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = set_beta_gamma(c, beta, gamma);
    if (rc) return rc;
    // per-row codes come from the permuted-column kernel; make them for lookups it has not seen
    for (int l = first; l < first + count; ++l) {
        if (c->code_valid[l]) continue;
        int e = l;
        while (e < first + count && !c->code_valid[e]) ++e;
        CU(ctx, launch_permuted(c, l, e - l, nullptr, nullptr, nullptr, nullptr, shard_map(c)));
        for (int x = l; x < e; ++x) c->code_valid[x] = true;
        l = e - 1;
    }
    if (count) {
        CU(ctx, launch_z_scan(c, first, count, nullptr, Z_dev, shard_map(c)));
    }
    return H2AES_OK;
}

// _host variants: stage through device memory a bounded number of lookups at a time.
static int lookups_per_stage(const h2aes_circuit *c, int ncols) {
    const uint64_t per = (uint64_t)ncols * c->sh.n() * 32;
    const uint64_t budget = 1ull << 30;
    uint64_t m = budget / per;
    return (int)(m < 1 ? 1 : m);
}

int32_t h2aes_lookup_permuted_host(h2aes_circuit *c, const uint64_t *theta, int32_t first,
                                   int32_t count, uint64_t *A, uint64_t *S, uint64_t *Ap,
                                   uint64_t *Sp) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!theta || !Ap || !Sp) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->sh.n_lookups())
        return fail(ctx, H2AES_ERR_ARG, "lookup range");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t col = c->sh.n() * 32;
    const int ncols = 2 + (A ? 1 : 0) + (S ? 1 : 0);
    const int step = lookups_per_stage(c, ncols);
    for (int done = 0; done < count || done == 0; done += step) {
        const int m = count - done < step ? count - done : step;
        int32_t rc = ensure_stage(c, (uint64_t)ncols * (m > 0 ? m : 1) * col);
        if (rc) return rc;
        char *p = (char *)c->d_stage;
        uint64_t *dA = nullptr, *dS = nullptr, *dAp, *dSp;
        if (A) { dA = (uint64_t *)p; p += (uint64_t)m * col; }
        if (S) { dS = (uint64_t *)p; p += (uint64_t)m * col; }
        dAp = (uint64_t *)p; p += (uint64_t)m * col;
        dSp = (uint64_t *)p;
        rc = h2aes_lookup_permuted_dev(c, theta, first + done, m, dA, dS, dAp, dSp);
        if (rc) return rc;
        const uint64_t off = (uint64_t)done * c->sh.n() * 4, bytes = (uint64_t)m * col;
        if (m > 0) {
            if (A) CU(ctx, cudaMemcpyAsync(A + off, dA, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            if (S) CU(ctx, cudaMemcpyAsync(S + off, dS, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaMemcpyAsync(Ap + off, dAp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaMemcpyAsync(Sp + off, dSp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (count == 0) break;
    }
    return check_device_flags(c);
}

int32_t h2aes_lookup_product_host(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                  int32_t first, int32_t count, uint64_t *Z) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!beta || !gamma || !Z) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (first < 0 || count < 0 || first + count > c->sh.n_lookups())
        return fail(ctx, H2AES_ERR_ARG, "lookup range");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t col = c->sh.n() * 32;
    const int step = lookups_per_stage(c, 1);
    for (int done = 0; done < count; done += step) {
        const int m = count - done < step ? count - done : step;
        int32_t rc = ensure_stage(c, (uint64_t)m * col);
        if (rc) return rc;
        rc = h2aes_lookup_product_dev(c, beta, gamma, first + done, m, (uint64_t *)c->d_stage);
        if (rc) return rc;
        CU(ctx, cudaMemcpyAsync(Z + (uint64_t)done * c->sh.n() * 4, c->d_stage, (uint64_t)m * col,
                                cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return check_device_flags(c);
}

int32_t h2aes_lookup_stats(h2aes_circuit *c, uint64_t *nonconst_rows) {
    if (!c || !nonconst_rows) return H2AES_ERR_ARG;
    if (!c->lk_alloc) return fail(c->ctx, H2AES_ERR_STATE, "no grand product has run yet");
    CU(c->ctx, cudaSetDevice(c->ctx->device));
    CU(c->ctx, cudaMemcpyAsync(nonconst_rows, c->d_znonconst, sizeof(uint64_t) * c->sh.n_lookups(),
                               cudaMemcpyDeviceToHost, c->ctx->stream));
    CU(c->ctx, cudaStreamSynchronize(c->ctx->stream));
    return H2AES_OK;
}

int32_t h2aes_lookup_check(h2aes_circuit *c) {
    if (!c) return H2AES_ERR_ARG;
    if (!c->lk_alloc) return H2AES_OK;
    CU(c->ctx, cudaSetDevice(c->ctx->device));
    return check_device_flags(c);
}

// ---------------------------------------------------------------------------------------------
// permutation argument on the device
// ---------------------------------------------------------------------------------------------
// halo2curves 0.6.1 bn256::Fr [UPSTREAM]: ROOT_OF_UNITY = 7^((r-1)/2^28), DELTA = 7^(2^28)
static fr_t fr_from_canonical_limbs(const uint32_t c[8]) {
    fr_t a;
    for (int i = 0; i < 8; ++i) a.v[i] = c[i];
    return fr_mul(a, fr_r2());
}
static fr_t domain_omega(int k) {  // EvaluationDomain::new: the 2^28-th root squared 28-k times
    const uint32_t root[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du,
                              0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    fr_t w = fr_from_canonical_limbs(root);
    for (int i = k; i < 28; ++i) w = fr_mul(w, w);
    return w;
}
static fr_t domain_delta() {
    const uint32_t d[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu,
                           0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
    return fr_from_canonical_limbs(d);
}

static int32_t ensure_lookup_ws(h2aes_circuit *c);

static int32_t ensure_perm_ws(h2aes_circuit *c) {
    h2aes_ctx *ctx = c->ctx;
    if (c->sh.k > 24) return fail(ctx, H2AES_ERR_ARG, "permutation argument: k <= 24");
    int32_t rc = ensure_lookup_ws(c);  // the grand-product scan workspace is shared
    if (rc) return rc;
    if (!c->perm_alloc) {
        const size_t n = c->sh.n(), nc = (size_t)perm_n_columns(c->sh), ns = (size_t)perm_n_sets(c->sh);
        DA(c->d_pmap, nc * n);
        DA(c->d_wlo, 1024);
        DA(c->d_whi, (n >> 10) + 1);
        DA(c->d_dpow, nc);
        DA(c->d_omega_tab, n);
        DA(c->d_bdp, nc);
        DA(c->d_pnum, ns * n);
        DA(c->d_pmask, ns * (n / 32));
        DA(c->d_ppre, ns * n);
        c->perm_alloc = true;
    }
    if (!c->perm_tables) {
        CU(ctx, launch_perm_tables(c, domain_omega(c->sh.k), domain_delta()));
        c->perm_tables = true;
    }
    return H2AES_OK;
}

// cycle structure of (shape, n_blocks) on the device, packed position << 24 | row: the two small
// successor templates are built on the host (a few thousand cells) and every cell's successor is
// evaluated in closed form by k_p_mapping (permmap.h)
static int32_t ensure_perm_mapping(h2aes_circuit *c, uint64_t n_blocks) {
    h2aes_ctx *ctx = c->ctx;
    if (c->perm_blocks == n_blocks) return H2AES_OK;
    if (!c->d_ptmpl) {
        PermTemplate t;
        perm_template_build(c->sh, &t);
        DA(c->d_ptmpl, 1);
        CU(ctx, cudaMemcpyAsync(c->d_ptmpl, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));  // `t` is on the stack
    }
    CU(ctx, launch_perm_mapping(c, n_blocks, c->d_ptmpl));
    c->perm_blocks = n_blocks;
    return H2AES_OK;
}

int32_t h2aes_perm_sigma_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *sigma_dev) {
    if (!c || !sigma_dev) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (c->sh.kind == KIND_KEYSCHED && n_blocks)
        return fail(ctx, H2AES_ERR_ARG, "the key-schedule circuit has no AES blocks");
    if (n_blocks > c->sh.capacity())
        return fail(ctx, H2AES_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
    if (c->n_shards != 1) return fail(ctx, H2AES_ERR_STATE, "permutation argument: unsharded only");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = ensure_perm_ws(c);
    if (rc) return rc;
    if ((rc = ensure_perm_mapping(c, n_blocks))) return rc;
    CU(ctx, launch_perm_sigma(c, sigma_dev));
    return H2AES_OK;
}

int32_t h2aes_perm_product_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                               const uint64_t *advice_dev, uint64_t *Zp_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!beta || !gamma || !advice_dev || !Zp_dev) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (!c->have_witness) return fail(ctx, H2AES_ERR_STATE, "the permutation product needs a witness");
    if (c->n_shards != 1) return fail(ctx, H2AES_ERR_STATE, "permutation argument: unsharded only");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = ensure_perm_ws(c);
    if (rc) return rc;
    if ((rc = ensure_perm_mapping(c, c->n_blocks))) return rc;
    // beta, gamma share the challenge slots of the lookup product; its cached tables are dropped
    memcpy(c->beta, beta, 32);
    memcpy(c->gamma, gamma, 32);
    c->have_beta_gamma = false;
    CU(ctx, cudaMemcpyAsync(c->d_chal + 1, c->beta, 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(c->d_chal + 2, c->gamma, 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, launch_perm_product(c, advice_dev, Zp_dev));
    return H2AES_OK;
}

// ---------------------------------------------------------------------------------------------
// row-sharded single circuit
// ---------------------------------------------------------------------------------------------
int32_t h2aes_lookup_hist_dev(h2aes_circuit *c, const uint64_t *theta, uint32_t **hist_dev,
                              uint64_t *hist_words) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!theta || !hist_dev || !hist_words) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (!c->have_witness) return fail(ctx, H2AES_ERR_STATE, "lookup columns need a witness first");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = ensure_lookup_ws(c);
    if (rc) return rc;
    bool changed;
    if ((rc = set_theta(c, theta, &changed))) return rc;
    const RowMap rm = shard_map(c);
    CU(ctx, launch_view_bins(c));
    CU(ctx, launch_lookup_hist(c, rm));
    c->have_hist = true;
    c->have_bins = c->have_beta_gamma = c->have_partial = false;
    *hist_dev = c->d_cnt;
    *hist_words = (uint64_t)c->sh.n_lookups() * BIN_STRIDE;
    return H2AES_OK;
}

int32_t h2aes_lookup_permuted_shard_dev(h2aes_circuit *c, uint64_t *A_dev, uint64_t *S_dev,
                                        uint64_t *Ap_dev, uint64_t *Sp_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!c->have_hist) return fail(ctx, H2AES_ERR_STATE, "h2aes_lookup_hist_dev must run first");
    CU(ctx, cudaSetDevice(ctx->device));
    if ((Ap_dev == nullptr) != (Sp_dev == nullptr) || (!Ap_dev && (A_dev || S_dev)))
        return fail(ctx, H2AES_ERR_ARG, "A' and S' must both be given (or no output at all)");
    int32_t rc = upload_active(c);
    if (rc) return rc;
    CU(ctx, launch_lookup_bins(c));
    c->have_bins = true;
    const RowMap rm = shard_map(c);
    // with no outputs this still records the per-row codes the grand product reads
    CU(ctx, launch_permuted(c, 0, c->sh.n_lookups(), A_dev, S_dev, Ap_dev, Sp_dev, rm));
    return H2AES_OK;
}

int32_t h2aes_lookup_partial_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                 uint64_t *Z_dev, uint64_t **partial_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!beta || !gamma || !Z_dev || !partial_dev) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (!c->have_bins) return fail(ctx, H2AES_ERR_STATE, "permuted columns must be built first");
    CU(ctx, cudaSetDevice(ctx->device));
    int32_t rc = set_beta_gamma(c, beta, gamma);
    if (rc) return rc;
    const RowMap rm = shard_map(c);
    // pass 1 of the scan: thread prefixes stay in the workspace, the product of the shard's
    // factors per lookup (its partial product) goes to d_partial; Z_dev is written by pass 2
    CU(ctx, launch_z_partial(c, Z_dev, rm));
    c->have_partial = true;
    c->partial_Z = Z_dev;
    *partial_dev = (uint64_t *)c->d_partial;
    return H2AES_OK;
}

int32_t h2aes_lookup_product_shard_dev(h2aes_circuit *c, const uint64_t *gathered_dev,
                                       uint64_t *Z_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    if (!gathered_dev || !Z_dev) return fail(ctx, H2AES_ERR_ARG, "null pointer");
    if (!c->have_partial) return fail(ctx, H2AES_ERR_STATE, "h2aes_lookup_partial_dev must run first");
    if (Z_dev != c->partial_Z)
        return fail(ctx, H2AES_ERR_ARG, "Z_dev must be the buffer given to h2aes_lookup_partial_dev");
    CU(ctx, cudaSetDevice(ctx->device));
    const RowMap rm = shard_map(c);
    CU(ctx, launch_shard_carry(c, gathered_dev, rm.segments()));
    CU(ctx, launch_z_finish(c, c->d_carry, Z_dev, rm));
    c->have_partial = false;
    return H2AES_OK;
}

// ---------------------------------------------------------------------------------------------
// debug / test hooks (host evaluation of the shared host+device source)
// ---------------------------------------------------------------------------------------------
int32_t h2aes_dbg_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!a || !b || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_mul(fr_from_limbs(a), fr_from_limbs(b)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_add(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!a || !b || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_add(fr_from_limbs(a), fr_from_limbs(b)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_sub(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!a || !b || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_sub(fr_from_limbs(a), fr_from_limbs(b)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_inv(const uint64_t *a, uint64_t *out) {
    if (!a || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_inv(fr_from_limbs(a)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_inv_fermat(const uint64_t *a, uint64_t *out) {
    if (!a || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_inv_fermat(fr_from_limbs(a)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_from_u64(uint64_t v, uint64_t *out) {
    if (!out) return H2AES_ERR_ARG;
    fr_t a = fr_zero();
    a.v[0] = (uint32_t)v;
    a.v[1] = (uint32_t)(v >> 32);
    fr_to_limbs(fr_mul(a, fr_r2()), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_to_canonical(const uint64_t *a, uint64_t *out) {
    if (!a || !out) return H2AES_ERR_ARG;
    fr_to_limbs(fr_to_canonical(fr_from_limbs(a)), out);
    return H2AES_OK;
}
int32_t h2aes_dbg_fr_mul_dev(h2aes_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t n,
                             uint64_t *out) {
    if (!ctx || !a || !b || !out || !n) return H2AES_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    fr_t *d = nullptr;
    CU(ctx, dalloc(&d, 3 * n));
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(d + n, b, n * 32, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = launch_dbg_fr_mul(ctx, d, d + n, n, d + 2 * n)) == cudaSuccess &&
        (e = cudaMemcpyAsync(out, d + 2 * n, n * 32, cudaMemcpyDeviceToHost, ctx->stream)) ==
            cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "h2aes_dbg_fr_mul_dev");
    return H2AES_OK;
}
int32_t h2aes_dbg_tables(int32_t sbox_mode, uint8_t *sbox256, uint8_t *mul2_256,
                         uint8_t *mul3_256) {
    if (sbox_mode != H2AES_SBOX_REFERENCE && sbox_mode != H2AES_SBOX_FIPS197) return H2AES_ERR_ARG;
    if (sbox256) make_sbox(sbox_mode, sbox256);
    for (int x = 0; x < 256; ++x) {
        if (mul2_256) mul2_256[x] = gf_mul2((uint8_t)x);
        if (mul3_256) mul3_256[x] = gf_mul3((uint8_t)x);
    }
    return H2AES_OK;
}
int32_t h2aes_dbg_block_trace(int32_t sbox_mode, const uint8_t *round_keys176,
                              const uint8_t *plaintext16, uint32_t *trace1360) {
    if (!round_keys176 || !plaintext16 || !trace1360) return H2AES_ERR_ARG;
    if (sbox_mode != H2AES_SBOX_REFERENCE && sbox_mode != H2AES_SBOX_FIPS197) return H2AES_ERR_ARG;
    uint8_t sbox[256], st[ST_BYTES];
    make_sbox(sbox_mode, sbox);
    aes_block_states(sbox, round_keys176, plaintext16, st);
    for (uint32_t q = 0; q < AES_ROWS; ++q) trace1360[q] = block_trace_word(st, round_keys176, q);
    return H2AES_OK;
}
int32_t h2aes_dbg_key_schedule_trace(int32_t sbox_mode, const uint8_t *key16,
                                     uint8_t *round_keys176, uint32_t *trace400,
                                     uint8_t *words96) {
    if (!key16 || !round_keys176) return H2AES_ERR_ARG;
    if (sbox_mode != H2AES_SBOX_REFERENCE && sbox_mode != H2AES_SBOX_FIPS197) return H2AES_ERR_ARG;
    uint8_t sbox[256];
    make_sbox(sbox_mode, sbox);
    expand_key(sbox, key16, round_keys176);
    if (trace400)
        for (uint32_t r = 0; r < KEY_SCHEDULE_ROWS; ++r)
            trace400[r] = ks_trace_word(sbox, round_keys176, r);
    if (words96)
        for (uint32_t r = 0; r < (uint32_t)WORDS_ROWS; ++r)
            words96[r] = (uint8_t)ks_words_byte(round_keys176, r);
    return H2AES_OK;
}

}  // extern "C"
