// gprod.cu — the running product of a column of factors (sm_100a).
//
// Replaces the scan of halo2_proofs v0.3.0 [UPSTREAM, un-vendored dependency of the reference]
//   plonk/lookup/prover.rs      Permuted::commit_product:  z[0]=1, z[i+1] = z[i]*f[i]   (call
//                               sites src/main.rs:94-101, benches/aes128.rs:100-107)
//   plonk/permutation/prover.rs Argument::commit:          the same over the column sets, each
//                               set starting from the previous set's last usable value
// which is sequential on the CPU.
//
// A factor is exactly one on most rows (k_permuted's chunk masks say where it is not) and one
// entry of an L2-resident per-bin table on most of the others (gathered through the row's 4-byte
// factor code).  Three kernels:
//   k_z_general  the remaining rows (active AND outside the zero run of A', or inside the table
//                rows): their factor is a product of two to four table entries.  The warps
//                compact these rows into shared-memory queues and evaluate 32 of them at a time
//                (every lane multiplies), storing the factor in the row's slot of the Z buffer;
//   k_z_totals   per tile of 32 x 1360 rows, every thread multiplies the factors of its ZT_R
//                consecutive rows; thread totals and the tile aggregate go to the workspace;
//   k_z_scan     per tile: exclusive product scan of the 512 thread totals (warp shuffles of
//                256-bit Montgomery products), decoupled look-back over the tile aggregates /
//                inclusive prefixes (tickets are handed out tile-major; the aggregates are all
//                there already, so nothing ever waits long), then every thread walks its rows
//                again from its exclusive prefix and stores Z: one 256-bit store per row, one
//                multiply per non-constant row.
// Lanes of a warp are 1360 rows (one AES block) apart: they see the same selector pattern, hence
// the same number of multiplies; the 16 warps of a CTA cover the 1360 block-relative rows.
// Bound by integer-multiply issue (one bn256 Montgomery product = 128 IMAD.WIDE), not by HBM:
// 2 multiplies per non-constant row, +1..3 for the rows of k_z_general.
#include "internal.h"

namespace h2 {

enum : int { ZSRC_LOOKUP = 0, ZSRC_BUFFER = 1 };

struct ZScanArgs {
    Shape sh;
    int first_lookup, count;
    // ZSRC_LOOKUP: factors from the per-row codes written by k_permuted
    const uint32_t *tracek;        // [sets][n]  op<<24 | bin of the row's input tuple
    const uint32_t *fcode;         // [L][n]     factor code of every row (k_permuted)
    const uint4 *glist;            // rows whose factor is not a table entry: (local row, code, bin
    const uint32_t *gbase, *gcount;//   of A, bin of S'); lookup l at [gbase[l], +gcount[l])
    const fr_t *chal;              // theta, beta, gamma
    const fr_t *val, *dval, *ib, *ig, *nac, *ibg, *ibp;  // [3][BIN_STRIDE]
    const fr_t *ftab;              // [4][3][BIN_STRIDE] nacp, ibgz, ibpz, sgz (k_inv_merge)
    const LookupMeta *meta;
    // both sources
    const uint32_t *mask;          // one bit per row: factor is not one; column li at li*mstride words
    uint64_t mstride;
    uint64_t *Z;                   // [count][rows][4]; ZSRC_BUFFER: holds the factors on entry
    uint64_t row_begin, rows;      // rows of every column handled by this launch
    uint64_t col_mask;             // n-1: row & col_mask = row inside its column (ZSRC_BUFFER scans
                                   // several columns as one long column)
    uint64_t n_tiles;
    const fr_t *carry_in;          // [L] or null: product of everything before row_begin
    fr_t *partial;                 // [L] or null: product of all factors (times carry_in)
    unsigned long long *nonconst;  // [L] or null: statistics
    // look-back workspace
    fr_t *tot;                     // [count*n_tiles][ZT_T] thread totals in row order
    fr_t *agg, *inc;               // [count*n_tiles]
    uint32_t *flag;                // [count*n_tiles]
    uint32_t *ctl;                 // ticket, finished, epoch
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ fr_t fr_ld_cg(const fr_t *p) {  // L2 (another SM wrote it)
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    fr_t r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
__device__ __forceinline__ void fr_st_cg(fr_t *p, const fr_t &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    __stcg(q, make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]));
    __stcg(q + 1, make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]));
}

// the (at most 96) mask bits of rows [lr0, lr0 + nr) of a column whose mask words start at `words`
__device__ __forceinline__ void span_mask(const uint32_t *__restrict__ words, uint64_t n_words,
                                          uint64_t lr0, uint32_t nr, uint32_t b[3]) {
    const uint64_t w0 = lr0 >> 5;
    const uint32_t sh = (uint32_t)lr0 & 31u;
    uint32_t x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = w0 + i < n_words ? __ldg(words + w0 + i) : 0u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b[i] = __funnelshift_r(x[i], x[i + 1], sh);
        const uint32_t lo = 32u * i;
        if (nr <= lo) b[i] = 0u;
        else if (nr < lo + 32u) b[i] &= (1u << (nr - lo)) - 1u;
    }
}

// the set bits of a thread's (up to 96) mask bits in ascending order
struct RowIter {
    uint32_t m0, m1, m2;
    __device__ __forceinline__ RowIter(const uint32_t b[3]) : m0(b[0]), m1(b[1]), m2(b[2]) {}
    __device__ __forceinline__ int next() {
        if (m0) { const int q = __ffs(m0) - 1; m0 &= m0 - 1u; return q; }
        if (m1) { const int q = 31 + __ffs(m1); m1 &= m1 - 1u; return q; }
        if (m2) { const int q = 63 + __ffs(m2); m2 &= m2 - 1u; return q; }
        return -1;
    }
};

// ZSRC_LOOKUP: what a thread keeps about its lookup (everything else is re-derived from the
// kernel parameters, which live in the constant bank and cost no registers)
struct LookupLane {
    uint64_t ln, sn;  // l * n, set * n
    uint32_t vo;      // view * BIN_STRIDE
    int l;
};

// factor of a row that is not a single table entry: active AND outside the zero run of A', or a
// table row.  `kin` = bin of the row's input value A, `kp` = bin of S' (k_permuted's list entry).
static __device__ __forceinline__ fr_t general_factor(const ZScanArgs &a, int l, uint64_t row,
                                                   uint32_t code, uint32_t kin, uint32_t kp) {
    const int ops = a.sh.ops_per_set();
    const uint32_t op = (uint32_t)(l % ops) + 1;
    const int v = view_of_op(op);
    const size_t vo = (size_t)v * BIN_STRIDE;
    const uint32_t kind = code >> 30, cls = (code >> 28) & 3u, k = code & 0xffffffu;
    fr_t den = fr_ldg_256_nc((cls == CODE_FIRST ? a.ibg : (cls == CODE_PAD ? a.ibp : a.ib)) + vo + k);
    fr_t num = fr_ldg_256_nc((kind == FK_TABLE_ROW ? a.dval : a.nac) + vo + kin);
    if (cls == CODE_OTHER) den = fr_mul(den, fr_ldg_256_nc(a.ig + vo + kp));
    if (kind == FK_TABLE_ROW) {  // S[i] is a table row: (A+beta)(S+gamma)
        const fr_t beta = fr_ld(a.chal + 1), gamma = fr_ld(a.chal + 2);
        const fr_t sg = fr_add(fr_ldg_256_nc(a.val + vo + entry_of_table_row(v, (uint32_t)row)), gamma);
        num = fr_mul(fr_add(num, beta), sg);
    }
    return fr_mul(num, den);
}

// The factor of a non-constant row is a table entry (ZSRC_LOOKUP, most rows) or stored in place
// in the Z buffer (ZSRC_BUFFER: every such row; ZSRC_LOOKUP: the rows k_z_general evaluated).
// Two-stage fetch so that the dependent loads of consecutive rows overlap the multiplies:
// row_code (the 4-byte factor code) two rows ahead, row_factor one row ahead.
template <int SRC>
__device__ __forceinline__ uint32_t row_code(const ZScanArgs &a, const LookupLane &cx, uint64_t row) {
    return SRC == ZSRC_LOOKUP ? __ldg(a.fcode + cx.ln + row) : 0u;
}
template <int SRC>
__device__ __forceinline__ fr_t row_factor(const ZScanArgs &a, const LookupLane &cx, uint32_t code,
                                           uint32_t q, const uint64_t *__restrict__ z) {
    if (SRC == ZSRC_LOOKUP && (code >> 30) == FK_TABLE)  // merged table, selector in bits 28..29
        return fr_ldg_256_nc(a.ftab + (size_t)((code >> 28) & 3u) * (N_VIEWS * BIN_STRIDE) + cx.vo +
                             (code & 0xffffffu));
    return fr_ldg_256(z + 4 * (size_t)q);
}

// ---------------------------------------------------------------------------------------------
// k_z_general: one thread per list entry, factor stored in the row's slot of the Z buffer
// ---------------------------------------------------------------------------------------------
constexpr int ZG_T = 128;
#ifndef ZG_OCC
#define ZG_OCC 8
#endif
__global__ void __launch_bounds__(ZG_T, ZG_OCC) k_z_general(const __grid_constant__ ZScanArgs a) {
    const int li = blockIdx.y, l = a.first_lookup + li;
    const uint32_t cnt = min(a.gcount[l], a.gbase[l + 1] - a.gbase[l]);
    const uint4 *__restrict__ gl = a.glist + a.gbase[l];
    uint64_t *__restrict__ zout = a.Z + (size_t)li * a.rows * 4;
    for (uint32_t i = blockIdx.x * ZG_T + threadIdx.x; i < cnt; i += gridDim.x * ZG_T) {
        const uint4 e = __ldg(gl + i);  // (local row, code, bin of A, bin of S')
        fr_store_256(zout + 4 * (size_t)e.x, general_factor(a, l, a.row_begin + e.x, e.y, e.z, e.w));
    }
}

// ---------------------------------------------------------------------------------------------
// k_z_totals / k_z_scan
// ---------------------------------------------------------------------------------------------
struct ZThread {
    int li, l, rk;
    uint64_t T, lr0, row0;
    size_t slot;
    uint32_t nr;
    uint64_t *z;
    LookupLane cx;
    uint32_t mb[3];
};
template <int SRC>
__device__ __forceinline__ void z_thread_setup(const ZScanArgs &a, uint32_t ticket, ZThread &t) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // tile-major, every column advances together; the tiles with the most multiplies first (the
    // table rows at the start of a column, then the end of the column, where A' is not zero), so
    // that the lightly loaded tiles fill the machine behind them.  Any order is deadlock free:
    // the look-back only needs the aggregates, which k_z_totals has published for every tile.
    t.li = (int)(ticket % (uint32_t)a.count);
    const uint64_t ord = ticket / (uint32_t)a.count;
    t.T = ord < 2 ? ord : a.n_tiles + 1 - ord;
    t.slot = (size_t)t.li * a.n_tiles + t.T;
    t.l = a.first_lookup + t.li;
    // my rows: local rows [lr0, lr0 + nr) of column li
    t.lr0 = t.T * ZT_ROWS + (uint64_t)lane * ZT_LANE + (uint64_t)wid * ZT_R;
    t.nr = t.lr0 >= a.rows ? 0u : (uint32_t)min((uint64_t)ZT_R, a.rows - t.lr0);
    t.row0 = a.row_begin + t.lr0;
    t.rk = lane * ZT_W + wid;  // rank of my rows among the CTA's threads
    t.z = a.Z + ((size_t)t.li * a.rows + t.lr0) * 4;
    t.cx.l = t.l;
    t.cx.ln = t.cx.sn = 0;
    t.cx.vo = 0;
    if (SRC == ZSRC_LOOKUP) {
        const int ops = a.sh.ops_per_set();
        t.cx.vo = (uint32_t)view_of_op((uint32_t)(t.l % ops) + 1) * BIN_STRIDE;
        t.cx.ln = (uint64_t)t.l * a.sh.n();
    }
    span_mask(a.mask + (size_t)t.li * a.mstride, (a.rows + 31) >> 5, t.lr0, t.nr, t.mb);
}

constexpr uint32_t ZS_BURST = 4;  // rows per thread staged between two cooperative stores (one line)
constexpr size_t ZS_SMEM = (size_t)ZT_W * ZS_BURST * 32 * sizeof(fr_t);
#ifndef ZTOT_OCC
#define ZTOT_OCC 2
#endif
#ifndef ZSCAN_OCC
#define ZSCAN_OCC 2
#endif

// thread totals in row order + the tile aggregate (flag state 1)
template <int SRC>
__global__ void __launch_bounds__(ZT_T, ZTOT_OCC) k_z_totals(const __grid_constant__ ZScanArgs a) {
    __shared__ fr_t s_w[ZT_W];
    __shared__ uint32_t s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_cnt = 0;
    ZThread t;
    z_thread_setup<SRC>(a, blockIdx.x, t);
    // software pipeline over the non-constant rows: code of row i+2, factor of row i+1, multiply i
    fr_t tot = fr_one();
    RowIter it(t.mb);
    int q1 = it.next(), q2 = it.next();
    uint32_t c1 = q1 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q1) : 0u;
    uint32_t c2 = q2 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q2) : 0u;
    fr_t f1 = fr_one();
    if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, t.z);
#pragma unroll 1
    while (q1 >= 0) {
        const fr_t f = f1;
        q1 = q2;
        c1 = c2;
        q2 = it.next();
        if (q2 >= 0) c2 = row_code<SRC>(a, t.cx, t.row0 + q2);
        if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, t.z);
        tot = fr_mul_inline(tot, f);
    }
    fr_st(a.tot + t.slot * ZT_T + t.rk, tot);
    // tile aggregate: product of everything (order is irrelevant)
    fr_t p = tot;
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) p = fr_mul(p, fr_shfl_xor(p, d));
    if (lane == 0) fr_st(&s_w[wid], p);
    uint32_t c = __popc(t.mb[0]) + __popc(t.mb[1]) + __popc(t.mb[2]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __syncthreads();
    if (lane == 0 && c) atomicAdd(&s_cnt, c);
    if (wid == 0) {
        fr_t w = lane < ZT_W ? fr_ld(&s_w[lane]) : fr_one();
#pragma unroll 1
        for (int d = ZT_W / 2; d > 0; d >>= 1) w = fr_mul(w, fr_shfl_xor(w, d));
        if (lane == 0) {
            fr_st_cg(a.agg + t.slot, w);
            __threadfence();
            st_volatile_u32(a.flag + t.slot, 2u * ld_volatile_u32(a.ctl + 2) + 1u);
        }
    }
    __syncthreads();
    if (a.nonconst && tid == 0 && s_cnt) atomicAdd(a.nonconst + t.l, (unsigned long long)s_cnt);
}

// partial[l] = product of the tile aggregates of lookup l (row-sharded circuit: the shard's
// partial product, exchanged before k_z_scan)
__global__ void __launch_bounds__(32) k_z_partial(const __grid_constant__ ZScanArgs a) {
    const int li = blockIdx.x, lane = threadIdx.x;
    fr_t p = fr_one();
    for (uint64_t T = lane; T < a.n_tiles; T += 32)
        p = fr_mul(p, fr_ld_cg(a.agg + (size_t)li * a.n_tiles + T));
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) p = fr_mul(p, fr_shfl_xor(p, d));
    if (lane == 0) fr_st(a.partial + a.first_lookup + li, p);
}

template <int SRC>
__global__ void __launch_bounds__(ZT_T, ZSCAN_OCC) k_z_scan(const __grid_constant__ ZScanArgs a) {
    __shared__ fr_t s_x[ZT_T];      // thread prefixes in row order
    extern __shared__ uint4 s_dyn[];
    fr_t *s_stage = reinterpret_cast<fr_t *>(s_dyn);  // [ZT_W][ZS_BURST][32] rows on their way out
    __shared__ fr_t s_w[ZT_W];      // warp totals, then warp exclusive prefixes (times the tile prefix)
    __shared__ uint32_t s_ticket, s_epoch;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(a.ctl, 1u);
        s_epoch = ld_volatile_u32(a.ctl + 2);
    }
    __syncthreads();
    const uint32_t ep2 = 2u * s_epoch;
    ZThread t;
    z_thread_setup<SRC>(a, s_ticket, t);
    const uint32_t usable = (uint32_t)a.sh.usable();

    // ---- exclusive scan of the tile's thread totals (rank order = thread order here) -----------
    const fr_t mine = fr_ld(a.tot + t.slot * ZT_T + tid);
    fr_t inc = mine;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const fr_t y = fr_shfl_up(inc, d);
        if (lane >= d) inc = fr_mul(inc, y);
    }
    if (lane == 31) fr_st(&s_w[wid], inc);
    __syncthreads();
    if (wid == 0) {
        const fr_t wt = lane < ZT_W ? fr_ld(&s_w[lane]) : fr_one();
        fr_t winc = wt;
#pragma unroll 1
        for (int d = 1; d < ZT_W; d <<= 1) {
            const fr_t y = fr_shfl_up(winc, d);
            if (lane >= d) winc = fr_mul(winc, y);
        }
        const fr_t aggv = fr_shfl_idx(winc, ZT_W - 1);
        fr_t wex = fr_shfl_up(winc, 1);
        if (lane == 0) wex = fr_one();
        // ---- decoupled look-back over the aggregates (state 1, all published by k_z_totals) and
        //      the inclusive prefixes (state 2) of the tiles before this one ----------------------
        fr_t excl = fr_one();
        if (t.T == 0) {
            if (a.carry_in) excl = fr_ld(a.carry_in + t.l);
        } else {
            int64_t p = (int64_t)t.T - 1;
            while (true) {
                const int64_t tt = p - lane;
                uint32_t st;
                do {
                    st = tt >= 0 ? ld_volatile_u32(a.flag + (size_t)t.li * a.n_tiles + tt) - ep2 : 2u;
                } while (__any_sync(0xffffffffu, st - 1u > 1u));
                __threadfence();
                const uint32_t have_inc = __ballot_sync(0xffffffffu, st == 2u);
                const int j = have_inc ? __ffs(have_inc) - 1 : 31;  // nearest inclusive prefix
                fr_t v = fr_one();  // (the "tile" before tile 0 holds the carry-in as its inclusive prefix)
                if (lane <= j && tt >= 0)
                    v = fr_ld_cg((st == 2u ? a.inc : a.agg) + (size_t)t.li * a.n_tiles + tt);
                else if (lane <= j && tt == -1 && a.carry_in)
                    v = fr_ld(a.carry_in + t.l);
#pragma unroll 1
                for (int d = 16; d > 0; d >>= 1) v = fr_mul(v, fr_shfl_xor(v, d));
                excl = fr_mul(excl, v);
                if (have_inc) break;
                p -= 32;
            }
        }
        if (lane == 0) {
            fr_st_cg(a.inc + t.slot, fr_mul(excl, aggv));
            __threadfence();
            st_volatile_u32(a.flag + t.slot, ep2 + 2u);
        }
        if (lane < ZT_W) fr_st(&s_w[lane], fr_mul(excl, wex));
    }
    __syncthreads();
    fr_t before = fr_shfl_up(inc, 1);
    if (lane == 0) before = fr_one();
    fr_st(&s_x[tid], fr_mul(fr_ld(&s_w[wid]), before));  // exclusive prefix of rank `tid`
    __syncthreads();

    // ---- Z[row] = running product; it changes only after a non-constant row ---------------------
    // Z has n-5 deterministic entries (rows 0..usable); the last 5 are blinding rows (zero).
    // A thread's rows are consecutive but the lanes of a warp are 1360 rows apart, so the values go
    // through shared memory four rows at a time and leave as whole 128-byte lines: lane l stores
    // row l%4 of the burst of lane 8i + l/4 (i = 0..3), 8 lines per store instruction instead of
    // 32 sectors in 32 different lines.
    fr_t cur = fr_ld(&s_x[t.rk]);
    const fr_t zero = fr_zero();
    fr_t *__restrict__ stage = s_stage + wid * (ZS_BURST * 32);
    const uint64_t lrw0 = t.T * ZT_ROWS + (uint64_t)wid * ZT_R;          // lane 0's first local row
    uint64_t *__restrict__ zw = a.Z + ((size_t)t.li * a.rows + lrw0) * 4;  // ... and its Z slot
    const uint64_t *__restrict__ z = t.z;
    // software pipeline over the non-constant rows: code of row i+2, factor of row i+1
    RowIter it(t.mb);
    int q1 = it.next(), q2 = it.next();
    uint32_t c1 = q1 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q1) : 0u;
    uint32_t c2 = q2 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q2) : 0u;
    fr_t f1 = fr_one();
    if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
#pragma unroll 1
    for (uint32_t r = 0; r < ZT_R; ++r) {
        const uint32_t jj = r & (ZS_BURST - 1);
        const bool ok = r < t.nr && (uint32_t)((t.row0 + r) & a.col_mask) <= usable;
        fr_st(stage + jj * 32 + lane, ok ? cur : zero);
        if ((int)r == q1) {  // (the factor was fetched before any lane overwrites its slot)
            const fr_t f = f1;
            q1 = q2;
            c1 = c2;
            q2 = it.next();
            if (q2 >= 0) c2 = row_code<SRC>(a, t.cx, t.row0 + q2);
            if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
            cur = fr_mul_inline(cur, f);
        }
        if (jj == ZS_BURST - 1 || r == ZT_R - 1) {
            __syncwarp();
            const uint32_t j2 = lane & (ZS_BURST - 1), rr = r - jj + j2;
#pragma unroll
            for (int i = 0; i < ZS_BURST; ++i) {
                const uint32_t tt = (32 / ZS_BURST) * i + (lane / ZS_BURST);
                const uint64_t lrow = lrw0 + (uint64_t)tt * ZT_LANE + rr;
                if (j2 <= jj && lrow < a.rows)
                    fr_store_256(zw + ((size_t)tt * ZT_LANE + rr) * 4, fr_ld(stage + j2 * 32 + tt));
            }
            __syncwarp();
        }
    }

    // the last CTA to finish rearms the control block for the next scan
    if (tid == 0) {
        __threadfence();
        const uint32_t fin = atomicAdd(a.ctl + 1, 1u);
        if (fin + 1u == gridDim.x) {
            a.ctl[0] = 0u;
            a.ctl[1] = 0u;
            a.ctl[2] = s_epoch + 1u;
            __threadfence();
        }
    }
}

// carry[l] = product of the partial products of the shards before `shard` (row-sharded circuit)
__global__ void k_shard_carry(const fr_t *__restrict__ gathered, int n_lookups, int shard,
                              fr_t *__restrict__ carry) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lookups) return;
    fr_t acc = fr_one();
    for (int s = 0; s < shard; ++s) acc = fr_mul(acc, fr_ld(gathered + (size_t)s * n_lookups + l));
    fr_st(carry + l, acc);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// k_z_scan stages its stores in 64 KB of dynamic shared memory (two CTAs per SM)
static cudaError_t scan_smem_optin() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_z_scan<ZSRC_LOOKUP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ZS_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_z_scan<ZSRC_BUFFER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ZS_SMEM);
    done = e == cudaSuccess;
    return e;
}

static ZScanArgs lookup_args(h2aes_circuit *c, int first, int count, const fr_t *carry_in,
                             uint64_t *Z, uint64_t row_begin, uint64_t row_end) {
    ZScanArgs a;
    a.sh = c->sh;
    a.first_lookup = first;
    a.count = count;
    a.tracek = c->d_tracek;
    a.fcode = c->d_codek;
    a.glist = c->d_glist; a.gbase = c->d_gbase; a.gcount = c->d_gcount;
    a.chal = c->d_chal;
    a.val = c->d_val; a.dval = c->d_dval; a.ib = c->d_ib; a.ig = c->d_ig; a.nac = c->d_nac;
    a.ibg = c->d_ibg; a.ibp = c->d_ibp;
    a.ftab = c->d_ftab;
    a.meta = c->d_meta;
    a.mask = c->d_zmask + (size_t)first * c->zm;
    a.mstride = c->zm;
    a.Z = Z;
    a.row_begin = row_begin;
    a.rows = row_end - row_begin;
    a.col_mask = c->sh.n() - 1;
    a.n_tiles = zscan_tiles(a.rows);
    a.carry_in = carry_in;
    a.partial = c->d_partial;
    a.nonconst = c->d_znonconst;
    a.tot = c->d_ztot; a.agg = c->d_zagg; a.inc = c->d_zinc; a.flag = c->d_zflag;
    a.ctl = c->d_zctl;
    return a;
}

// factors of the rows that are not table entries, then thread totals + tile aggregates
static cudaError_t launch_totals(h2aes_circuit *c, const ZScanArgs &a) {
    h2aes_ctx *ctx = c->ctx;
    cudaStream_t s = ctx->stream;
    cudaError_t e0 = cudaMemsetAsync(c->d_znonconst + a.first_lookup, 0, (size_t)a.count * 8, s);
    if (e0) return e0;
    unsigned gx = cdiv(a.rows / 4, ZG_T);  // lists are a fraction of the rows; CTAs past the end exit
    const unsigned cap = (unsigned)ctx->sm_count * 4;
    if (gx > cap) gx = cap;
    prof_begin(ctx, "k_z_general");
    k_z_general<<<dim3(gx, a.count), ZG_T, 0, s>>>(a);
    prof_end(ctx);
    prof_begin(ctx, "k_z_totals");
    k_z_totals<ZSRC_LOOKUP><<<(unsigned)(a.n_tiles * a.count), ZT_T, 0, s>>>(a);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

// Z of lookups [first, first+count) over rows [row_begin, row_end)
cudaError_t launch_z_scan(h2aes_circuit *c, int first, int count, const fr_t *carry_in,
                          uint64_t *Z, uint64_t row_begin, uint64_t row_end) {
    h2aes_ctx *ctx = c->ctx;
    const ZScanArgs a = lookup_args(c, first, count, carry_in, Z, row_begin, row_end);
    cudaError_t e = launch_totals(c, a);
    if (e) return e;
    if (cudaError_t es = scan_smem_optin()) return es;
    prof_begin(ctx, "k_z_scan");
    k_z_scan<ZSRC_LOOKUP><<<(unsigned)(a.n_tiles * count), ZT_T, ZS_SMEM, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// row-sharded circuit, first half: thread totals and tile aggregates of the shard's rows stay in
// the workspace, partial[l] = product of the shard's factors
cudaError_t launch_z_partial(h2aes_circuit *c, uint64_t *Z, uint64_t row_begin, uint64_t row_end) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    const ZScanArgs a = lookup_args(c, 0, L, nullptr, Z, row_begin, row_end);
    cudaError_t e = launch_totals(c, a);
    if (e) return e;
    prof_begin(ctx, "k_z_partial");
    k_z_partial<<<L, 32, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// second half: Z from the stored totals, starting from the carry of the shards before this one
cudaError_t launch_z_finish(h2aes_circuit *c, const fr_t *carry_in, uint64_t *Z, uint64_t row_begin,
                            uint64_t row_end) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    const ZScanArgs a = lookup_args(c, 0, L, carry_in, Z, row_begin, row_end);
    if (cudaError_t es = scan_smem_optin()) return es;
    prof_begin(ctx, "k_z_scan");
    k_z_scan<ZSRC_LOOKUP><<<(unsigned)(a.n_tiles * L), ZT_T, ZS_SMEM, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// running product of factors stored in place: `cols` columns of n rows scanned as ONE column
// (the permutation argument: set s starts from the last usable value of set s-1)
cudaError_t launch_z_scan_buffer(h2aes_circuit *c, int cols, uint64_t *Z, const uint32_t *mask) {
    h2aes_ctx *ctx = c->ctx;
    ZScanArgs a = lookup_args(c, 0, 1, nullptr, Z, 0, (uint64_t)cols * c->sh.n());
    a.mask = mask;
    a.mstride = 0;
    a.partial = nullptr;
    a.nonconst = nullptr;
    prof_begin(ctx, "k_z_totals");
    k_z_totals<ZSRC_BUFFER><<<(unsigned)a.n_tiles, ZT_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    if (cudaError_t es = scan_smem_optin()) return es;
    prof_begin(ctx, "k_z_scan");
    k_z_scan<ZSRC_BUFFER><<<(unsigned)a.n_tiles, ZT_T, ZS_SMEM, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_shard_carry(h2aes_circuit *c, const uint64_t *gathered_dev) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    prof_begin(ctx, "k_shard_carry");
    k_shard_carry<<<cdiv(L, 64), 64, 0, ctx->stream>>>((const fr_t *)gathered_dev, L, c->shard,
                                                       c->d_carry);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace h2
