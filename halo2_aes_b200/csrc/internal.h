// internal.h — state behind the opaque handles of include/h2aes.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/h2aes.h"
#include "fr.cuh"
#include "layout.h"
#include "permmap.h"

namespace h2 {

// number of distinct table tuples seen through an m-column lookup (m = view + 2)
constexpr int N_VIEWS = 3;
constexpr uint32_t VIEW0_ENTRIES = 5 * 256 + 1;  // (tag,x) pairs + the zero row
constexpr uint32_t VIEW_MAX_ENTRIES = (uint32_t)TABLE_ROWS;
constexpr uint32_t BUCKET_BITS = 12;  // room for the N_SPLIT + 1 buckets of lookup.cu
constexpr uint32_t N_BUCKETS = 1u << BUCKET_BITS;
constexpr uint32_t BIN_STRIDE = 66568;  // >= VIEW_MAX_ENTRIES + 2, multiple of 8

H2_HD uint32_t view_entries(int v) { return v == 0 ? VIEW0_ENTRIES : VIEW_MAX_ENTRIES; }
// view (0: 2 columns, 1: 3 columns, 2: 4 columns) used by the lookup of op `op`
H2_HD int view_of_op(uint32_t op) { return op == OP_RANGE ? 0 : (op == OP_XOR ? 2 : 1); }
// table row -> entry index inside a view
H2_HD uint32_t entry_of_table_row(int v, uint32_t trow) {
    if (v != 0) return trow;
    if (trow < TROW_XOR) return trow;  // U8 rows 0..255, SBOX rows 256..511
    if (trow < TROW_MUL2) return 512 + ((trow - TROW_XOR) >> 8);
    if (trow < TROW_ZERO) return 768 + (trow - TROW_MUL2);  // MUL2 768.., MUL3 1024..
    return 1280;
}
// any table row that realises entry e of view 0 (used to evaluate the tuple)
H2_HD uint32_t table_row_of_entry(int v, uint32_t e) {
    if (v != 0) return e;
    if (e < 512) return e;
    if (e < 768) return TROW_XOR + ((e - 512) << 8);
    if (e < 1280) return TROW_MUL2 + (e - 768);
    return TROW_ZERO;
}
// how many table rows (of the 66561 assigned ones) collapse onto entry e of view v
H2_HD uint32_t entry_multiplicity(int v, uint32_t e) {
    return (v == 0 && e >= 512 && e < 768) ? 256u : 1u;
}

struct LookupMeta {  // per lookup, produced by the bin-scan kernel
    uint32_t n_nonempty;  // NZ: bins with at least one input row
    uint32_t n_left;      // NL: bins with leftover table multiplicity
    uint32_t repeated;    // R = usable - NZ
    uint32_t k_zero;      // bin of the all-zero tuple (inactive rows)
    uint32_t bad;         // != 0: an input value has no table row (cannot happen by construction)
    uint32_t k_pad;       // bin of table row 0 (the value halo2 pads the table with)
    uint32_t pad[2];
};

// Debug build (`make -C halo2_aes_b200/csrc debug`, -DH2AES_DEBUG): bounds assertions on every
// data-dependent index of the lookup / grand-product kernels (bins, runs, list entries, block
// totals).  compute-sanitizer is not available on the pool's boxes; the parity suite run against
// the debug library (H2AES_LIB=.../libh2aes_debug.so pytest -m gpu) is the memcheck substitute.
#if defined(H2AES_DEBUG)
#include <cassert>
#define H2_ASSERT(cond) assert(cond)
#else
#define H2_ASSERT(cond) ((void)0)
#endif

#if defined(__CUDACC__)
// ---- device helpers shared by lookup.cu and gprod.cu ---------------------------------------
__device__ __forceinline__ fr_t fr_from_tab(const uint4 *__restrict__ frtab, uint32_t b) {
    uint4 lo = frtab[2 * b], hi = frtab[2 * b + 1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
__device__ __forceinline__ fr_t fr_ld(const fr_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
__device__ __forceinline__ void fr_st(fr_t *p, const fr_t &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// bins (k of A', k' of S') of usable row i, and how S' relates to A'
constexpr uint32_t CODE_FIRST = 0u, CODE_PAD = 1u, CODE_OTHER = 2u;  // stored in bits 30..31

// factor code of a row for the grand product (bits 31..30 kind, 29..28 selector, 23..0 index):
//   kind 0  factor exactly one (constant row)                       -> nothing to do
//   kind 1  factor = one entry of a merged table: selector 0 nacp[bin of A], 1 ibgz[bin of A'],
//           2 ibpz[bin of A'], 3 sgz[table entry of the row]
//   kind 2  factor = nac[bin of A] * den(bin of A', class in the selector bits)
//   kind 3  the same with S a table row: (A+beta)(S+gamma) * den
constexpr uint32_t FK_ONE = 0u, FK_TABLE = 1u, FK_GENERAL = 2u, FK_TABLE_ROW = 3u;
constexpr uint32_t FSEL_NACP = 0u, FSEL_IBGZ = 1u, FSEL_IBPZ = 2u, FSEL_SGZ = 3u;
__device__ __forceinline__ uint32_t fcode_pack(uint32_t kind, uint32_t sel, uint32_t idx) {
    return (kind << 30) | (sel << 28) | idx;
}


__device__ __forceinline__ fr_t fr_shfl_up(const fr_t &a, int d) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
    return r;
}
__device__ __forceinline__ fr_t fr_shfl_down(const fr_t &a, int d) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
    return r;
}
__device__ __forceinline__ fr_t fr_shfl_xor(const fr_t &a, int d) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], d);
    return r;
}


__device__ __forceinline__ fr_t fr_shfl_idx(const fr_t &a, int src) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
    return r;
}
#endif

// ---- grand-product scan (gprod.cu) ----------------------------------------------------------------
// A column is cut into blocks of ZB_R consecutive rows (one thread each).  The 32 lanes of a warp
// take blocks that are AES_ROWS rows apart, so they sit at the SAME block-relative rows of 32
// different AES blocks: same selector pattern, same number of multiplies.  A row tile = 32 x
// AES_ROWS rows = ZB_W warps' worth of blocks.
#ifndef ZB_ROWS
#define ZB_ROWS 40
#endif
constexpr uint32_t ZT_LANE = (uint32_t)AES_ROWS;          // rows between two lanes of a warp
// block totals per look-back tile of k_z_prefix (gprod.cu asserts its tile is this; api.cu sizes
// the look-back workspace by it)
constexpr int ZP_TILE_TOTALS = 1024;
constexpr uint32_t ZB_R = ZB_ROWS;                        // consecutive rows per thread
constexpr uint32_t ZB_W = ZT_LANE / ZB_R;                 // warps per row tile (34)
constexpr uint32_t ZT_ROWS = 32u * ZT_LANE;               // rows per row tile (43520)
static_assert(ZB_R * ZB_W == ZT_LANE, "the blocks of a lane partition one AES block's rows");
static_assert(ZB_R <= 96, "three 32-bit mask words per thread");
H2_HD uint64_t zscan_blocks(uint64_t rows) { return (rows + ZB_R - 1) / ZB_R; }
H2_HD uint64_t zscan_tiles(uint64_t rows) { return (rows + ZT_ROWS - 1) / ZT_ROWS; }

}  // namespace h2

struct h2aes_prof_rec {
    const char *name;
    cudaEvent_t a, b;
};

struct h2aes_ctx {
    int device = 0;
    bool profiling = false;                // h2aes_ctx_profile: CUDA events around every kernel
    std::vector<h2aes_prof_rec> prof;
    cudaStream_t stream = nullptr;
    int sbox_mode = 0;
    int sm_count = 148;
    uint8_t h_sbox[256];
    uint8_t *d_sbox = nullptr;   // 256 bytes
    uint4 *d_frtab = nullptr;    // 256 entries x 32 bytes: byte -> Fr (Montgomery)
    uint64_t launches = 0;
    std::string last_error;
    cudaStream_t side = nullptr;  // second stream + fork/join events (comm.cu)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void *comm = nullptr;        // ncclComm_t of h2aes_comm_init (comm.cu)
    int comm_rank = 0, comm_world = 1;
};

struct h2aes_circuit {
    h2aes_ctx *ctx = nullptr;
    h2::Shape sh{};
    uint64_t n_blocks = 0;
    uint8_t key[16] = {0};         // key of the last h2aes_witness_* call (h2aes_encrypt_append_dev)
    bool have_witness = false;
    int shard = 0, n_shards = 1;
    uint64_t shard_ts = 0;         // 0: contiguous rows, else rows per ownership tile (dealt round-robin)

    // witness workspace
    uint8_t *d_rk = nullptr;       // 176
    uint8_t *d_key = nullptr;      // 16
    uint8_t *d_states = nullptr;   // capacity * ST_BYTES
    uint64_t states_cap = 0;
    uint32_t *d_trace = nullptr;   // n_sets * n words (a0 | a1<<8 | a2<<16 | op<<24)
    uint32_t *d_tracek = nullptr;  // n_sets * n words (op<<24 | bin of the row's tuple), per theta
    uint32_t *d_codek = nullptr;   // [L][n] factor code of every row (fcode_pack), written by k_permuted
    uint32_t *d_codep = nullptr;   // [L][n] bin of S' where class == OTHER
    uint32_t *d_glist = nullptr;   // local rows whose factor is not a table entry (k_general_list)
    uint32_t *d_gbase = nullptr;   // [L+1] first list entry of every lookup (capacity bound, api.cu)
    uint32_t *d_gcount = nullptr;  // [L] entries written by the last k_permuted
    bool *code_valid = nullptr;    // [L] host flags: codes of lookup l match the current bins

    // lookup workspace (allocated on first use)
    bool lk_alloc = false;
    bool have_theta = false, have_hist = false, have_bins = false, have_beta_gamma = false,
         have_partial = false;
    uint64_t theta[4] = {0, 0, 0, 0}, beta[4] = {0, 0, 0, 0}, gamma[4] = {0, 0, 0, 0};
    h2::fr_t *d_chal = nullptr;      // theta, beta, gamma
    // per view
    h2::fr_t *d_val = nullptr;       // [3][BIN_STRIDE] compressed value of entry e
    h2::fr_t *d_canon = nullptr;     // [3][BIN_STRIDE] canonical form (sort key)
    h2::fr_t *d_canon_b = nullptr;   // [3][BIN_STRIDE] the same in bucket order
    uint32_t *d_active = nullptr;    // [L] rows with the lookup's selector enabled
    uint32_t *d_jidx = nullptr;      // [L][idx_stride] coarse index into offz (k_run_index)
    uint32_t *d_lidx = nullptr;      // [L][idx_stride] coarse index into loffz
    uint64_t idx_stride = 0;
    uint4 *d_lbsum = nullptr;        // [L][tiles] per-tile sums for the run-structure kernels
    h2::fr_t *d_dval = nullptr;      // [3][BIN_STRIDE] distinct values, ascending
    uint32_t *d_rank = nullptr;      // [3][BIN_STRIDE] entry -> bin
    uint32_t *d_ts = nullptr;        // [3][BIN_STRIDE] table multiplicity of bin
    uint32_t *d_sorted = nullptr;    // [3][BIN_STRIDE] entries in ascending value order
    uint32_t *d_slot = nullptr;      // [3][BIN_STRIDE] bucket-ordered entries
    uint32_t *d_head = nullptr;      // [3][BIN_STRIDE]
    uint32_t *d_bkt = nullptr;       // [3][BIN_STRIDE] bucket of every entry
    uint32_t *d_bkt_b = nullptr;     // [3][BIN_STRIDE] the same in bucket order
    h2::fr_t *d_splval = nullptr;    // [3][N_BUCKETS] sorted sample of the canonical values (splitters)
    uint32_t *d_splidx = nullptr;    // [3][N_BUCKETS] ... and their entry indices (ties)
    uint32_t *d_bhist = nullptr;     // [3][N_BUCKETS+1]
    uint32_t *d_bcur = nullptr;      // [3][N_BUCKETS]
    uint32_t *d_viewmeta = nullptr;  // [3][8]: [1] = number of bins
    h2::fr_t *d_ib = nullptr;        // [3][BIN_STRIDE] 1/(dval+beta)
    h2::fr_t *d_ig = nullptr;        // [3][BIN_STRIDE] 1/(dval+gamma)
    h2::fr_t *d_nac = nullptr;       // [3][BIN_STRIDE] (dval+beta)*(pad+gamma)
    h2::fr_t *d_ibg = nullptr;       // [3][BIN_STRIDE] ib[k]*ig[k]
    h2::fr_t *d_ibp = nullptr;       // [3][BIN_STRIDE] ib[k]*ig[k_pad]
    h2::fr_t *d_ftab = nullptr;      // [4][3][BIN_STRIDE] the four merged tables below, by FSEL_*
    h2::fr_t *d_nacp = nullptr;      // [3][BIN_STRIDE] nac[k]*ibp[k_zero]
    h2::fr_t *d_ibgz = nullptr;      // [3][BIN_STRIDE] nac[k_zero]*ibg[k]
    h2::fr_t *d_ibpz = nullptr;      // [3][BIN_STRIDE] nac[k_zero]*ibp[k]
    h2::fr_t *d_sgz = nullptr;       // [3][BIN_STRIDE] (val[e]+gamma)/(pad+gamma) per table entry
    // per lookup
    uint32_t *d_cnt = nullptr;       // [L][BIN_STRIDE] input rows per bin
    uint32_t *d_offz = nullptr;      // [L][BIN_STRIDE] first row of each nonempty bin (+ end)
    uint32_t *d_nzk = nullptr;       // [L][BIN_STRIDE] bin of each nonempty slot
    uint32_t *d_loffz = nullptr;     // [L][BIN_STRIDE] leftover prefix (+ end)
    uint32_t *d_lk = nullptr;        // [L][BIN_STRIDE] bin of each leftover slot
    h2::LookupMeta *d_meta = nullptr;  // [L]
    // grand-product scan workspace
    unsigned long long *d_znonconst = nullptr;  // [L] rows whose factor is not one (last product)
    uint32_t *d_zmask = nullptr;      // [2][L][zm] one bit per row: factor is not one ; factor is not a table entry
    h2::fr_t *d_ztot = nullptr;       // [L][zblocks] block totals, then their exclusive prefixes
    h2::fr_t *d_zagg = nullptr;       // [ztiles] look-back of k_z_prefix: tile aggregates
    h2::fr_t *d_zinc = nullptr;       // [ztiles] inclusive prefixes
    uint32_t *d_zflag = nullptr;      // [ztiles] 2*epoch + {1: aggregate, 2: inclusive}
    uint32_t *d_zctl = nullptr;       // [4] ticket, finished CTAs, epoch (device-resident: graph replay safe)
    uint64_t ztiles = 0, zblocks = 0; // look-back tiles / block totals the workspace holds
    uint32_t *d_err = nullptr;        // device error flags: [0] zero denominator, [1] not in table
    uint32_t *h_err = nullptr;        // pinned mirror of d_err
    h2::fr_t *d_partial = nullptr;    // [L] shard partial products
    h2::fr_t *d_carry = nullptr;      // [L] product of the partials of the shards before this one
    uint64_t *d_gathered = nullptr;   // [64][L][4] every shard's partial products (comm.cu)
    uint64_t zm = 0;
    uint64_t *partial_Z = nullptr;    // Z buffer holding the factors between phases 2a and 2b

    // permutation argument (allocated on first use)
    bool perm_alloc = false, perm_tables = false;
    uint64_t perm_blocks = ~0ull;     // n_blocks the mapping on the device was built for
    uint32_t *d_pmap = nullptr;       // [n_cols][n] packed mapping: position << 24 | row
    h2::PermTemplate *d_ptmpl = nullptr;  // successor templates of one block / the key schedule
    h2::fr_t *d_wlo = nullptr;        // [1024] omega^t
    h2::fr_t *d_whi = nullptr;        // [n/1024] omega^(1024 b)
    h2::fr_t *d_dpow = nullptr;       // [n_cols] delta^p
    h2::fr_t *d_omega_tab = nullptr;  // [n] omega^j
    h2::fr_t *d_bdp = nullptr;        // [n_cols] beta * delta^p
    h2::fr_t *d_pnum = nullptr;       // [n_sets][n] numerators
    uint32_t *d_pmask = nullptr;      // [n_sets][n/32] chunk masks of the permutation product
    h2::fr_t *d_ppre = nullptr;       // [n_sets][n] running denominator products

    // staging for the _host entry points
    void *d_stage = nullptr;
    uint64_t stage_bytes = 0;
    uint8_t *d_pt = nullptr;
    uint64_t pt_cap = 0;
    uint8_t *d_ct = nullptr;
};

// evaluation domain (ntt.cu); the extended Lagrange columns are built on first use (quotient.cu)
struct h2aes_domain {
    h2aes_ctx *ctx = nullptr;
    int j = 0, k = 0, extended_k = 0;
    uint64_t n = 0, en = 0, quotient_len = 0;
    h2::fr_t *d_tw = nullptr;       // [en] extended_omega^e
    h2::fr_t *d_ctw = nullptr;      // compact per-size tables omega_m^e at [m + e], m <= 2^11
    h2::fr_t *d_consts = nullptr;   // 4 groups of 3: coset in, 1/n, 1/en * zeta^-i, (spare)
    h2::fr_t *d_scratch = nullptr;  // [batch][en]
    int batch = 0;
    h2::fr_t *d_lagrange = nullptr; // [3][en] l0, l_last, l_active_row on the extended coset
    h2::fr_t *d_tinv = nullptr;     // [2^(extended_k-k)] 1 / ((zeta*extended_omega^i)^n - 1)
};

// per-kernel timing hooks (no-ops unless h2aes_ctx_profile(ctx, 1) was called)
static inline void prof_begin(h2aes_ctx *ctx, const char *name) {
    if (!ctx->profiling) return;
    h2aes_prof_rec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, ctx->stream);
    ctx->prof.push_back(r);
}
static inline void prof_end(h2aes_ctx *ctx) {
    if (!ctx->profiling || ctx->prof.empty()) return;
    cudaEventRecord(ctx->prof.back().b, ctx->stream);
}
