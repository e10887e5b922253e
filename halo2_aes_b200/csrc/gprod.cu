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
// factor code).  Four kernels, none of which keeps a CTA waiting on another:
//   k_z_general  the remaining rows (active AND outside the zero run of A', or inside the table
//                rows): product of two to four table entries, one thread per entry of the list
//                k_permuted wrote; the factor goes to the row's slot of the Z buffer;
//   k_z_totals   the column is cut into blocks of ZB_R consecutive rows; one thread multiplies
//                the factors of one block -> tot[block];
//   k_z_prefix   exclusive product scan of the block totals in one pass: per tile of 1024
//                totals a block-wide scan (warp shuffles of 256-bit Montgomery products) and a
//                decoupled look-back over the preceding tiles' aggregates / inclusive prefixes
//                (tickets are handed out in tile order, so a tile only waits on running tiles);
//   k_z_write    one thread per block again: walk the rows from the block's exclusive prefix,
//                one multiply per non-constant row; the values leave through shared memory as
//                whole 128-byte lines (one 256-bit store per row).
// Work unit = warp.  The 32 lanes of a warp take blocks that are 1360 rows (one AES block) apart:
// they sit at the SAME block-relative rows of 32 different AES blocks, see the same selector
// pattern and execute the same number of multiplies.  Warps are dealt heaviest region first
// (the table rows at the start of a column, then the end of the column, where A' is not zero).
// Bound by integer-multiply issue (one bn256 Montgomery product = 128 IMAD.WIDE), not by HBM:
// 2 multiplies per non-constant row, +1..3 for the rows of k_z_general, ~2 per block for the scan.
#include <algorithm>
#include <cstdlib>

#include "internal.h"

namespace h2 {

enum : int { ZSRC_LOOKUP = 0, ZSRC_BUFFER = 1 };

struct ZScanArgs {
    Shape sh;
    int first_lookup, count;
    // ZSRC_LOOKUP: factors from the per-row codes written by k_permuted
    const uint32_t *tracek;        // [sets][n]  op<<24 | bin of the row's input tuple
    const uint32_t *fcode;         // [L][n]     factor code of every row (k_permuted)
    const uint32_t *codep;         // [L][n]     bin of S' where it is neither A' nor the padding value
    const uint32_t *glist;         // local rows whose factor is not a table entry (k_general_list);
    const uint32_t *gbase, *gcount;//   lookup l at [gbase[l], +gcount[l])
    const fr_t *chal;              // theta, beta, gamma
    const fr_t *val, *dval, *ib, *ig, *nac, *ibg, *ibp;  // [3][BIN_STRIDE]
    const fr_t *ftab;              // [4][3][BIN_STRIDE] nacp, ibgz, ibpz, sgz (k_inv_merge)
    const LookupMeta *meta;
    // both sources
    const uint32_t *mask;          // one bit per row: factor is not one; column li at li*mstride words
    uint64_t mstride;
    uint64_t *Z;                   // [count][rows][4]; ZSRC_BUFFER: holds the factors on entry
    RowMap rm;                     // local row -> row of the column (a shard holds rm.rows rows)
    uint64_t rows;                 // = rm.rows: local rows of every column handled by this launch
    uint64_t col_mask;             // n-1: row & col_mask = row inside its column (ZSRC_BUFFER scans
                                   // several columns as one long column)
    uint64_t n_tiles;              // row tiles of 32 x 1360 local rows per column (work distribution)
    // The scan runs per "segment" = the local rows of one ownership tile of one column: rows that
    // are consecutive in the column.  Unsharded / contiguous shards: one segment per column.
    uint32_t segs;                 // segments (local ownership tiles) per column
    uint64_t n_blocks;             // blocks of ZB_R rows per segment = block totals per segment
    uint64_t n_ptiles;             // tiles of ZP_TILE block totals per segment (k_z_prefix)
    const fr_t *carry_in;          // [L][segs] or null: product of everything before the segment
    fr_t *partial;                 // [L][segs] or null: product of all factors of the segment
    unsigned long long *nonconst;  // [L] or null: statistics
    fr_t *tot;                     // [count][segs][n_blocks] block totals, then their exclusive prefixes
    // look-back workspace of k_z_prefix
    fr_t *agg, *inc;               // [count*n_ptiles]
    uint32_t *flag;                // [count*n_ptiles] 2*epoch + {1: aggregate, 2: inclusive prefix}
    uint32_t *ctl;                 // ticket, finished CTAs, epoch
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
// (three scalars, not an array: an indexed array would live in local memory)
__device__ __forceinline__ void span_mask(const uint32_t *__restrict__ words, uint64_t n_words,
                                          uint64_t lr0, uint32_t nr, uint32_t &b0, uint32_t &b1,
                                          uint32_t &b2) {
    const uint64_t w0 = lr0 >> 5;
    const uint32_t sh = (uint32_t)lr0 & 31u;
    const uint32_t x0 = w0 < n_words ? __ldg(words + w0) : 0u;
    const uint32_t x1 = w0 + 1 < n_words ? __ldg(words + w0 + 1) : 0u;
    const uint32_t x2 = w0 + 2 < n_words ? __ldg(words + w0 + 2) : 0u;
    const uint32_t x3 = w0 + 3 < n_words ? __ldg(words + w0 + 3) : 0u;
    auto clip = [nr](uint32_t v, uint32_t lo) {
        return nr <= lo ? 0u : (nr < lo + 32u ? v & ((1u << (nr - lo)) - 1u) : v);
    };
    b0 = clip(__funnelshift_r(x0, x1, sh), 0u);
    b1 = clip(__funnelshift_r(x1, x2, sh), 32u);
    b2 = clip(__funnelshift_r(x2, x3, sh), 64u);
}

// the set bits of a thread's (up to 96) mask bits in ascending order
struct RowIter {
    uint32_t m0, m1, m2;
    __device__ __forceinline__ RowIter(uint32_t b0, uint32_t b1, uint32_t b2) : m0(b0), m1(b1), m2(b2) {}
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
    H2_ASSERT(kind >= FK_GENERAL && k < BIN_STRIDE && kin < BIN_STRIDE && kp < BIN_STRIDE && row < a.sh.n());
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
    H2_ASSERT(q < ZB_R && (SRC != ZSRC_LOOKUP || (code >> 30) != FK_ONE));
    H2_ASSERT(SRC != ZSRC_LOOKUP || (code >> 30) != FK_TABLE || (code & 0xffffffu) < BIN_STRIDE);
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
    const int ops = a.sh.ops_per_set();
    const uint32_t op = (uint32_t)(l % ops) + 1;
    const uint64_t n = a.sh.n();
    const uint32_t cnt = min(a.gcount[l], a.gbase[l + 1] - a.gbase[l]);
    const uint32_t *__restrict__ gl = a.glist + a.gbase[l];
    const uint32_t *__restrict__ fc = a.fcode + (uint64_t)l * n;
    const uint32_t *__restrict__ cp = a.codep + (uint64_t)l * n;
    const uint32_t *__restrict__ tk = a.tracek + (uint64_t)(l / ops) * n;
    const uint32_t k_zero = a.meta[l].k_zero;
    uint64_t *__restrict__ zout = a.Z + (size_t)li * a.rows * 4;
    for (uint32_t i = blockIdx.x * ZG_T + threadIdx.x; i < cnt; i += gridDim.x * ZG_T) {
        const uint32_t lrow = __ldg(gl + i);
        H2_ASSERT(lrow < a.rows);
        const uint64_t row = a.rm.global(lrow);
        const uint32_t code = __ldg(fc + row), t = __ldg(tk + row);
        const uint32_t kin = (t >> 24) == op ? (t & 0xffffffu) : k_zero;  // bin of the row's input value
        const uint32_t kp = ((code >> 28) & 3u) == CODE_OTHER ? __ldg(cp + row) : 0u;
        fr_store_256(zout + 4 * (size_t)lrow, general_factor(a, l, row, code, kin, kp));
    }
}

// ---------------------------------------------------------------------------------------------
// k_z_totals / k_z_write: one thread per block of ZB_R rows, one warp per 32 blocks 1360 rows apart
// ---------------------------------------------------------------------------------------------
struct ZBlock {
    int li, l;
    uint64_t T, w;        // row tile and slice of the tile this warp works on
    uint64_t b, lr0, row0;  // my block (index into tot[]), its first local / absolute row
    uint32_t nr;          // rows of my block inside the column
    LookupLane cx;
    uint32_t mb0, mb1, mb2;
};
// warp task -> (column, row tile, slice); heaviest tiles first (see the file header)
template <int SRC>
__device__ __forceinline__ bool z_block_setup(const ZScanArgs &a, uint64_t task, ZBlock &t) {
    const int lane = threadIdx.x & 31;
    const uint64_t per_tile = (uint64_t)a.count * ZB_W;
    if (task >= a.n_tiles * per_tile) return false;
    const uint64_t ord = task / per_tile, rem = task - ord * per_tile;
    t.T = ord < 2 ? ord : a.n_tiles + 1 - ord;
    t.li = (int)(rem / ZB_W);
    t.w = rem - (uint64_t)t.li * ZB_W;
    t.l = a.first_lookup + t.li;
    t.lr0 = ((t.T * 32 + (uint64_t)lane) * ZB_W + t.w) * ZB_R;
    t.nr = t.lr0 >= a.rows ? 0u : (uint32_t)min((uint64_t)ZB_R, a.rows - t.lr0);
    t.row0 = a.rm.global(t.lr0);  // (a block never straddles two ownership tiles)
    // slot of my block's total: [column][segment][block of the segment]
    const uint64_t seg = a.segs > 1 ? t.lr0 / a.rm.ts : 0;
    t.b = seg < a.segs ? ((uint64_t)t.li * a.segs + seg) * a.n_blocks + (t.lr0 - seg * (a.segs > 1 ? a.rm.ts : 0)) / ZB_R
                       : ~0ull;
    t.cx.l = t.l;
    t.cx.ln = t.cx.sn = 0;
    t.cx.vo = 0;
    if (SRC == ZSRC_LOOKUP) {
        const int ops = a.sh.ops_per_set();
        t.cx.vo = (uint32_t)view_of_op((uint32_t)(t.l % ops) + 1) * BIN_STRIDE;
        t.cx.ln = (uint64_t)t.l * a.sh.n();
    }
    span_mask(a.mask + (size_t)t.li * a.mstride, (a.rows + 31) >> 5, t.lr0, t.nr, t.mb0, t.mb1, t.mb2);
    return true;
}

constexpr int ZB_T = 128;  // threads per CTA (4 independent warps)
#ifndef ZTOT_OCC
#define ZTOT_OCC 8
#endif
#ifndef ZWRITE_OCC
#define ZWRITE_OCC 6
#endif
#ifndef ZWRITE_SHALLOW
#define ZWRITE_DEEP 1  // codes three rows ahead, factors two: 4 % faster than two / one at 6 CTAs per SM
#endif

template <int SRC>
__global__ void __launch_bounds__(ZB_T, ZTOT_OCC) k_z_totals(const __grid_constant__ ZScanArgs a) {
    // (persistent form: with a grid smaller than the task count every warp walks several tasks,
    // heaviest first; see launch_totals)
    for (uint64_t task = (uint64_t)blockIdx.x * (ZB_T / 32) + (threadIdx.x >> 5);;
         task += (uint64_t)gridDim.x * (ZB_T / 32)) {
    ZBlock t;
    if (!z_block_setup<SRC>(a, task, t)) return;
    const uint64_t *__restrict__ z = a.Z + ((size_t)t.li * a.rows + t.lr0) * 4;
    // software pipeline over the non-constant rows: code of row i+2, factor of row i+1, multiply i
    fr_t tot = fr_one();
    RowIter it(t.mb0, t.mb1, t.mb2);
    int q1 = it.next(), q2 = it.next();
    uint32_t c1 = q1 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q1) : 0u;
    uint32_t c2 = q2 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q2) : 0u;
    fr_t f1 = fr_one();
    if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
#pragma unroll 1
    while (q1 >= 0) {
        const fr_t f = f1;
        q1 = q2;
        c1 = c2;
        q2 = it.next();
        if (q2 >= 0) c2 = row_code<SRC>(a, t.cx, t.row0 + q2);
        if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
        tot = fr_mul_inline(tot, f);
    }
    H2_ASSERT(!(t.b != ~0ull && (t.nr || a.segs > 1)) || t.b < (uint64_t)a.count * a.segs * a.n_blocks);
    if (t.b != ~0ull && (t.nr || a.segs > 1)) fr_st(a.tot + t.b, tot);  // (padding blocks of a partial segment: one)
    if (a.nonconst) {
        uint32_t c = __popc(t.mb0) + __popc(t.mb1) + __popc(t.mb2);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(a.nonconst + t.l, (unsigned long long)c);
    }
    }
}

// ---------------------------------------------------------------------------------------------
// k_z_prefix: tot[] -> exclusive prefixes (times the carry-in), single pass, decoupled look-back
// ---------------------------------------------------------------------------------------------
#ifndef ZP_THREADS
#define ZP_THREADS 128
#endif
#ifndef ZP_ELEMS
#define ZP_ELEMS 8
#endif
#ifndef ZP_STAGE
#define ZP_STAGE 0
#endif
constexpr int ZP_T = ZP_THREADS, ZP_E = ZP_ELEMS, ZP_TILE = ZP_T * ZP_E;
static_assert(ZP_TILE == ZP_TILE_TOTALS, "the look-back workspace (api.cu) is sized for tiles of ZP_TILE_TOTALS block totals");
__global__ void __launch_bounds__(ZP_T) k_z_prefix(const __grid_constant__ ZScanArgs a) {
    __shared__ fr_t s_w[ZP_T / 32];
    __shared__ uint32_t s_ticket, s_epoch;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(a.ctl, 1u);
        s_epoch = ld_volatile_u32(a.ctl + 2);
    }
    __syncthreads();
    // tickets in tile order, all columns advancing together: a tile only waits on running tiles
    const uint32_t ep2 = 2u * s_epoch;
    const uint32_t nseg = (uint32_t)a.count * a.segs;           // independent scans
    const int li = (int)(s_ticket % nseg);                      // (column, segment)
    const uint64_t T = s_ticket / nseg;
    const size_t slot = (size_t)li * a.n_ptiles + T;
    const int l = a.first_lookup * (int)a.segs + li;            // index into carry_in / partial
    fr_t *__restrict__ x = a.tot + (size_t)li * a.n_blocks;
    const uint64_t base = T * ZP_TILE + (uint64_t)tid * ZP_E;
    // ZP_STAGE: where a thread's elements wait for the output pass.  0: registers (164 of them:
    // 3 CTAs per SM, the 650 tiles of the bench shape take two waves); 1: shared memory (32 KiB per
    // CTA); 2: read again from global memory
#if ZP_STAGE == 1
    __shared__ uint4 s_v[ZP_E * 2][ZP_T];
#endif
#if ZP_STAGE == 0
    fr_t v[ZP_E];
#endif
    fr_t p = fr_one();
#pragma unroll
    for (int j = 0; j < ZP_E; ++j) {
        const fr_t vj = base + j < a.n_blocks ? fr_ld(x + base + j) : fr_one();
#if ZP_STAGE == 0
        v[j] = vj;
#elif ZP_STAGE == 1
        s_v[2 * j][tid] = make_uint4(vj.v[0], vj.v[1], vj.v[2], vj.v[3]);
        s_v[2 * j + 1][tid] = make_uint4(vj.v[4], vj.v[5], vj.v[6], vj.v[7]);
#endif
        p = j == 0 ? vj : fr_mul(p, vj);
    }
    fr_t inc = p;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const fr_t y = fr_shfl_up(inc, d);
        if (lane >= d) inc = fr_mul(inc, y);
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        constexpr int NW = ZP_T / 32;
        const fr_t wt = lane < NW ? s_w[lane] : fr_one();
        fr_t winc = wt;
#pragma unroll 1
        for (int d = 1; d < NW; d <<= 1) {
            const fr_t y = fr_shfl_up(winc, d);
            if (lane >= d) winc = fr_mul(winc, y);
        }
        const fr_t aggv = fr_shfl_idx(winc, NW - 1);
        fr_t wex = fr_shfl_up(winc, 1);
        if (lane == 0) wex = fr_one();
        // ---- decoupled look-back ----------------------------------------------------------------
        fr_t excl = fr_one();
        if (T == 0) {
            if (a.carry_in) excl = fr_ld(a.carry_in + l);
        } else {
            if (lane == 0) {
                fr_st_cg(a.agg + slot, aggv);
                __threadfence();
                st_volatile_u32(a.flag + slot, ep2 + 1u);
            }
            int64_t pp = (int64_t)T - 1;
            while (true) {
                const int64_t tt = pp - lane;
                uint32_t st;
                do {
                    st = tt >= 0 ? ld_volatile_u32(a.flag + (size_t)li * a.n_ptiles + tt) - ep2 : 2u;
                } while (__any_sync(0xffffffffu, st - 1u > 1u));
                __threadfence();
                const uint32_t have_inc = __ballot_sync(0xffffffffu, st == 2u);
                const int j = have_inc ? __ffs(have_inc) - 1 : 31;  // nearest inclusive prefix
                fr_t u = fr_one();  // (the "tile" before tile 0 holds the carry-in)
                if (lane <= j && tt >= 0)
                    u = fr_ld_cg((st == 2u ? a.inc : a.agg) + (size_t)li * a.n_ptiles + tt);
                else if (lane <= j && tt == -1 && a.carry_in)
                    u = fr_ld(a.carry_in + l);
#pragma unroll 1
                for (int d = 16; d > 0; d >>= 1) u = fr_mul(u, fr_shfl_xor(u, d));
                excl = fr_mul(excl, u);
                if (have_inc) break;
                pp -= 32;
            }
        }
        const fr_t incl = fr_mul(excl, aggv);
        if (lane == 0) {
            fr_st_cg(a.inc + slot, incl);
            __threadfence();
            st_volatile_u32(a.flag + slot, ep2 + 2u);
            if (a.partial && T + 1 == a.n_ptiles) fr_st(a.partial + l, incl);
        }
        if (lane < NW) s_w[lane] = fr_mul(excl, wex);
    }
    __syncthreads();
    fr_t run = fr_shfl_up(inc, 1);
    if (lane == 0) run = fr_one();
    run = fr_mul(run, s_w[wid]);
#pragma unroll
    for (int j = 0; j < ZP_E; ++j) {
#if ZP_STAGE == 0
        const fr_t vj = v[j];
#elif ZP_STAGE == 1
        const uint4 lo = s_v[2 * j][tid], hi = s_v[2 * j + 1][tid];
        const fr_t vj = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
#else
        const fr_t vj = base + j < a.n_blocks ? fr_ld(x + base + j) : fr_one();
#endif
        if (base + j < a.n_blocks) fr_st(x + base + j, run);
        if (j + 1 < ZP_E) run = fr_mul(run, vj);
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

// ---------------------------------------------------------------------------------------------
// k_z_write: Z[row] = running product from the block's exclusive prefix
// ---------------------------------------------------------------------------------------------
constexpr uint32_t ZS_BURST = 4;  // rows per thread staged between two cooperative stores (one line)
template <int SRC>
__global__ void __launch_bounds__(ZB_T, ZWRITE_OCC) k_z_write(const __grid_constant__ ZScanArgs a) {
    __shared__ uint4 s_stage[ZB_T / 32][ZS_BURST * 32 * 2];  // (uint4: the type it is accessed as)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t task = (uint64_t)blockIdx.x * (ZB_T / 32) + wid;; task += (uint64_t)gridDim.x * (ZB_T / 32)) {
    ZBlock t;
    if (!z_block_setup<SRC>(a, task, t)) return;
    const uint32_t usable = (uint32_t)a.sh.usable();
    // Z has n-5 deterministic entries (rows 0..usable); the last 5 are blinding rows (zero).
    // A thread's rows are consecutive but the lanes of a warp are 1360 rows apart, so the values go
    // through shared memory four rows at a time and leave as whole 128-byte lines: lane l stores
    // row l%4 of the burst of lane 8i + l/4 (i = 0..3), 8 lines per store instruction instead of
    // 32 sectors in 32 different lines.
    fr_t cur = fr_one();
    if (t.nr) cur = fr_ld(a.tot + t.b);
    const fr_t zero = fr_zero();
    uint4 *stage = s_stage[wid];
    const uint64_t lrw0 = (t.T * 32 * ZB_W + t.w) * ZB_R;                  // lane 0's first local row
    uint64_t *zw = a.Z + ((size_t)t.li * a.rows + lrw0) * 4;  // ... and its Z slot
    const uint64_t *z = a.Z + ((size_t)t.li * a.rows + t.lr0) * 4;
#ifdef ZWRITE_DEEP
    // software pipeline over the non-constant rows: code of row i+3, factor of row i+2
    RowIter it(t.mb0, t.mb1, t.mb2);
    int q1 = it.next(), q2 = it.next(), q3 = it.next();
    uint32_t c1 = q1 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q1) : 0u;
    uint32_t c2 = q2 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q2) : 0u;
    uint32_t c3 = q3 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q3) : 0u;
    fr_t f1 = fr_one(), f2 = fr_one();
    if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
    if (q2 >= 0) f2 = row_factor<SRC>(a, t.cx, c2, (uint32_t)q2, z);
#else
    // software pipeline over the non-constant rows: code of row i+2, factor of row i+1
    RowIter it(t.mb0, t.mb1, t.mb2);
    int q1 = it.next(), q2 = it.next();
    uint32_t c1 = q1 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q1) : 0u;
    uint32_t c2 = q2 >= 0 ? row_code<SRC>(a, t.cx, t.row0 + q2) : 0u;
    fr_t f1 = fr_one();
    if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
#endif
#pragma unroll 1
    for (uint32_t r = 0; r < ZB_R; ++r) {
        const uint32_t jj = r & (ZS_BURST - 1);
        const bool ok = r < t.nr && (uint32_t)((t.row0 + r) & a.col_mask) <= usable;
#ifdef ZWRITE_DIRECT
        if (r < t.nr) fr_store_256(a.Z + ((size_t)t.li * a.rows + t.lr0 + r) * 4, ok ? cur : zero);
#else
        {
            const fr_t o = ok ? cur : zero;
            stage[2 * (jj * 32 + lane)] = make_uint4(o.v[0], o.v[1], o.v[2], o.v[3]);
            stage[2 * (jj * 32 + lane) + 1] = make_uint4(o.v[4], o.v[5], o.v[6], o.v[7]);
        }
#endif
        if ((int)r == q1) {  // (the factor was fetched before any lane overwrites its slot)
            const fr_t f = f1;
#ifdef ZWRITE_DEEP
            f1 = f2;
            q1 = q2;
            q2 = q3;
            c2 = c3;
            q3 = it.next();
            if (q3 >= 0) c3 = row_code<SRC>(a, t.cx, t.row0 + q3);
            if (q2 >= 0) f2 = row_factor<SRC>(a, t.cx, c2, (uint32_t)q2, z);
#else
            q1 = q2;
            c1 = c2;
            q2 = it.next();
            if (q2 >= 0) c2 = row_code<SRC>(a, t.cx, t.row0 + q2);
            if (q1 >= 0) f1 = row_factor<SRC>(a, t.cx, c1, (uint32_t)q1, z);
#endif
            cur = fr_mul_inline(cur, f);
        }
#ifndef ZWRITE_DIRECT
        if (jj == ZS_BURST - 1 || r == ZB_R - 1) {
            __syncwarp();
            const uint32_t j2 = lane & (ZS_BURST - 1), rr = r - jj + j2;
#pragma unroll
            for (int i = 0; i < (int)ZS_BURST; ++i) {
                const uint32_t tt = (32 / ZS_BURST) * i + (lane / ZS_BURST);
                const uint64_t lrow = lrw0 + (uint64_t)tt * ZT_LANE + rr;
                if (j2 <= jj && lrow < a.rows)
                    store_256_cs(zw + ((size_t)tt * ZT_LANE + rr) * 4, stage[2 * (j2 * 32 + tt)],
                              stage[2 * (j2 * 32 + tt) + 1]);
            }
            __syncwarp();
        }
#endif
    }
    }
}

// carry[l][t] = product of the partial products of every segment that precedes segment t of this
// shard in the column: segment t' of shard g' is ownership tile t' * G + g' (row-sharded circuit).
// gathered = [G][L][segs] (segments a shard does not hold: one).
__global__ void __launch_bounds__(128)
k_shard_carry(const fr_t *__restrict__ gathered, int n_lookups, uint32_t segs, uint32_t G,
              uint32_t shard, fr_t *__restrict__ carry) {
    // one warp per (lookup, segment): the lanes split the preceding ownership tiles, then a
    // shuffle tree (<= 4 + 5 dependent multiplies instead of one chain over up to segs * G tiles)
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // (lookup, segment)
    const int lane = threadIdx.x & 31;
    if (i >= (uint32_t)n_lookups * segs) return;  // (whole warps)
    const uint32_t l = i / segs, t = i - l * segs;
    const uint32_t mine = t * G + shard;          // ownership tiles of the column before mine
    fr_t acc = fr_one();
    for (uint32_t j = lane; j < mine; j += 32) {
        const uint32_t tt = j / G, gg = j - tt * G;
        acc = fr_mul(acc, fr_ld(gathered + ((size_t)gg * n_lookups + l) * segs + tt));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = fr_mul(acc, fr_shfl_xor(acc, d));
    if (lane == 0) fr_st(carry + i, acc);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

static ZScanArgs lookup_args(h2aes_circuit *c, int first, int count, const fr_t *carry_in,
                             uint64_t *Z, const RowMap &rm) {
    ZScanArgs a;
    a.sh = c->sh;
    a.first_lookup = first;
    a.count = count;
    a.tracek = c->d_tracek;
    a.fcode = c->d_codek;
    a.codep = c->d_codep;
    a.glist = c->d_glist; a.gbase = c->d_gbase; a.gcount = c->d_gcount;
    a.chal = c->d_chal;
    a.val = c->d_val; a.dval = c->d_dval; a.ib = c->d_ib; a.ig = c->d_ig; a.nac = c->d_nac;
    a.ibg = c->d_ibg; a.ibp = c->d_ibp;
    a.ftab = c->d_ftab;
    a.meta = c->d_meta;
    a.mask = c->d_zmask + (size_t)first * c->zm;
    a.mstride = c->zm;
    a.Z = Z;
    a.rm = rm;
    a.rows = rm.rows;
    a.col_mask = c->sh.n() - 1;
    a.segs = rm.segments();
    a.n_tiles = a.segs > 1 ? a.segs : zscan_tiles(a.rows);  // (tiled shards: ownership tile = row tile)
    a.n_blocks = zscan_blocks(a.segs > 1 ? rm.ts : a.rows);
    a.n_ptiles = (a.n_blocks + ZP_TILE - 1) / ZP_TILE;
    a.carry_in = carry_in;
    a.partial = c->d_partial;
    a.nonconst = c->d_znonconst;
    a.tot = c->d_ztot; a.agg = c->d_zagg; a.inc = c->d_zinc; a.flag = c->d_zflag;
    a.ctl = c->d_zctl;
    return a;
}
// H2AES_Z_CTAS_PER_SM (tuning, default 0 = one warp per task): run the multiply-bound kernels as
// persistent grids of that many CTAs per SM, leaving room for the HBM-bound kernels of the other
// circuit in flight to be dispatched beside them.  Measured (profiles/persist.sh): the persistent
// form is slower by itself (static striding over tasks of very different weight: k_z_totals 0.22
// -> 0.29 ms, k_z_write 0.31 -> 0.38 ms) and the co-scheduled step does not win it back (1.48 ms
// per circuit at best against 1.30 ms), so it is off.
#ifndef Z_CTAS_PER_SM
#define Z_CTAS_PER_SM 0
#endif
static int z_ctas_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("H2AES_Z_CTAS_PER_SM");
        v = s ? atoi(s) : Z_CTAS_PER_SM;
    }
    return v;
}
static inline unsigned block_grid(const h2aes_ctx *ctx, const ZScanArgs &a) {
    const unsigned full = cdiv(a.n_tiles * (uint64_t)a.count * ZB_W, ZB_T / 32);  // one warp per task
    const unsigned cap = (unsigned)(ctx->sm_count * z_ctas_per_sm());
    return cap && cap < full ? cap : full;
}

// factors of the rows that are not table entries, then the block totals
static cudaError_t launch_totals(h2aes_circuit *c, const ZScanArgs &a) {
    h2aes_ctx *ctx = c->ctx;
    cudaStream_t s = ctx->stream;
    cudaError_t e0 = cudaMemsetAsync(c->d_znonconst + a.first_lookup, 0, (size_t)a.count * 8, s);
    if (e0) return e0;
    unsigned gx = cdiv(a.rows / 4, ZG_T);  // lists are a fraction of the rows; CTAs past the end exit
    const unsigned cap = z_ctas_per_sm() ? std::max(1u, (unsigned)(ctx->sm_count * z_ctas_per_sm()) / (unsigned)a.count)
                                         : (unsigned)ctx->sm_count * 4;
    if (gx > cap) gx = cap;
    static int zg = -1;  // H2AES_ZG_CTAS (tuning): CTAs per lookup of k_z_general
    if (zg < 0) {
        const char *e = getenv("H2AES_ZG_CTAS");
        zg = e ? atoi(e) : 0;
    }
    if (zg) gx = std::min(gx, (unsigned)zg);
    prof_begin(ctx, "k_z_general");
    k_z_general<<<dim3(gx, a.count), ZG_T, 0, s>>>(a);
    prof_end(ctx);
    prof_begin(ctx, "k_z_totals");
    k_z_totals<ZSRC_LOOKUP><<<block_grid(ctx, a), ZB_T, 0, s>>>(a);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}
static cudaError_t launch_prefix(h2aes_circuit *c, const ZScanArgs &a) {
    h2aes_ctx *ctx = c->ctx;
    prof_begin(ctx, "k_z_prefix");
    k_z_prefix<<<(unsigned)(a.n_ptiles * a.count * a.segs), ZP_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// Z of lookups [first, first+count) over rows [row_begin, row_end)
cudaError_t launch_z_scan(h2aes_circuit *c, int first, int count, const fr_t *carry_in,
                          uint64_t *Z, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    const ZScanArgs a = lookup_args(c, first, count, carry_in, Z, rm);
    cudaError_t e = launch_totals(c, a);
    if (e) return e;
    if ((e = launch_prefix(c, a))) return e;
    prof_begin(ctx, "k_z_write");
    k_z_write<ZSRC_LOOKUP><<<block_grid(ctx, a), ZB_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// row-sharded circuit, first half: the block prefixes of the shard's rows (without the carry of
// the shards before it) stay in the workspace, partial[l] = product of the shard's factors
cudaError_t launch_z_partial(h2aes_circuit *c, uint64_t *Z, const RowMap &rm) {
    const int L = c->sh.n_lookups();
    const ZScanArgs a = lookup_args(c, 0, L, nullptr, Z, rm);
    cudaError_t e = launch_totals(c, a);
    if (e) return e;
    return launch_prefix(c, a);
}

// second half: fold the carry of everything before a segment into its prefixes, then write Z
__global__ void __launch_bounds__(256) k_z_carry(const __grid_constant__ ZScanArgs a) {
    const int li = blockIdx.y;  // (column, segment)
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_blocks) return;
    fr_t *p = a.tot + (size_t)li * a.n_blocks + i;
    fr_st(p, fr_mul(fr_ld(p), fr_ld(a.carry_in + (size_t)a.first_lookup * a.segs + li)));
}
cudaError_t launch_z_finish(h2aes_circuit *c, const fr_t *carry_in, uint64_t *Z, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    const ZScanArgs a = lookup_args(c, 0, L, carry_in, Z, rm);
    if (carry_in) {
        prof_begin(ctx, "k_z_carry");
        k_z_carry<<<dim3(cdiv(a.n_blocks, 256), L * a.segs), 256, 0, ctx->stream>>>(a);
        prof_end(ctx);
        ctx->launches += 1;
    }
    prof_begin(ctx, "k_z_write");
    k_z_write<ZSRC_LOOKUP><<<block_grid(ctx, a), ZB_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// running product of factors stored in place: `cols` columns of n rows scanned as ONE column
// (the permutation argument: set s starts from the last usable value of set s-1)
cudaError_t launch_z_scan_buffer(h2aes_circuit *c, int cols, uint64_t *Z, const uint32_t *mask) {
    h2aes_ctx *ctx = c->ctx;
    ZScanArgs a = lookup_args(c, 0, 1, nullptr, Z, row_map_of((uint64_t)cols * c->sh.n(), 0, 1, 0));
    a.mask = mask;
    a.mstride = 0;
    a.partial = nullptr;
    a.nonconst = nullptr;
    prof_begin(ctx, "k_z_totals");
    k_z_totals<ZSRC_BUFFER><<<block_grid(ctx, a), ZB_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    cudaError_t e = launch_prefix(c, a);
    if (e) return e;
    prof_begin(ctx, "k_z_write");
    k_z_write<ZSRC_BUFFER><<<block_grid(ctx, a), ZB_T, 0, ctx->stream>>>(a);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_shard_carry(h2aes_circuit *c, const uint64_t *gathered_dev, uint32_t segs) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    prof_begin(ctx, "k_shard_carry");
    k_shard_carry<<<cdiv((uint64_t)L * segs * 32, 128), 128, 0, ctx->stream>>>(
        (const fr_t *)gathered_dev, L, segs, (uint32_t)c->n_shards, (uint32_t)c->shard, c->d_carry);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace h2
