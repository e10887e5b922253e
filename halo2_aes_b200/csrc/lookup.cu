// lookup.cu — halo2 lookup-argument helper columns on the GPU (sm_100a).
//
// Replaces halo2_proofs v0.3.0 plonk/lookup/prover.rs [UPSTREAM, un-vendored dependency of the
// reference; call sites src/main.rs:94-101, benches/aes128.rs:100-107,
// benches/key_schedule.rs:117-124]:
//   commit_permuted  : A = fold(q*tag, q*x, ...; theta), S = fold(table cols; theta),
//                      A' = sort(A[..u]), S' = permuted table (first occurrences aligned, the
//                      leftover table multiset filled ascending into the repeated rows taken
//                      from the END), u = 2^k - 6;
//   commit_product   : Z[0]=1, Z[i+1] = Z[i] (A[i]+beta)(S[i]+gamma) / ((A'[i]+beta)(S'[i]+gamma)).
//
// Design (no sort over n rows, no per-row inversion):
//   * every input value is the compression of one of <= 66561 table tuples, so the rows are
//     counted per *bin* (distinct compressed value) from the 4-byte trace word of each row;
//   * the bins are ordered once per theta per view (2-, 3-, 4-column tables) with a bucket sort
//     (splitters from a sorted sample of the values themselves) + exact 256-bit ranking inside
//     each bucket;
//   * A'/S' are expanded from the run structure (binary search of the row in the run offsets);
//     the same kernel classifies every row's grand-product factor (4-byte code, one mask bit per
//     row for "factor is not one" and one for "factor is not a single table entry");
//   * Z: per-bin inverse tables here (Montgomery-trick batch inversion over the bins, not the
//     rows), the running product itself in gprod.cu (k_z_general / k_z_totals / k_z_prefix with
//     a single-pass decoupled-look-back scan / k_z_write).
#include <algorithm>
#include <cstdlib>

#include "internal.h"

namespace h2 {
size_t stream_pad_bytes(const void *func);
cudaError_t launch_z_scan_buffer(h2aes_circuit *c, int cols, uint64_t *Z, const uint32_t *mask);

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// Exclusive scan of data[0..n) in place by ONE CTA (<= 1024 threads); returns the total to all.
// Rounds of SCAN_E*blockDim consecutive elements, thread t on SCAN_E consecutive elements of the
// round (64 contiguous bytes per thread); few rounds because every round costs a block barrier
// chain and a global-memory round trip.
constexpr int SCAN_E = 16;
__device__ uint32_t block_exclusive_scan_u32(uint32_t *data, uint32_t n) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t t = threadIdx.x, nt = blockDim.x;
    const int lane = t & 31, wid = t >> 5, nw = nt >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += SCAN_E * nt) {
        const uint32_t i = base + SCAN_E * t;
        uint32_t v[SCAN_E];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < SCAN_E; ++j) {
            v[j] = i + j < n ? data[i + j] : 0u;
            sum += v[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = lane < nw ? s_warp[lane] : 0u;
            uint32_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += y;
            }
            s_warp[lane] = winc - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[wid] + inc - sum;
#pragma unroll
        for (int j = 0; j < SCAN_E; ++j) {
            if (i + j < n) data[i + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (t == nt - 1) s_carry = run;  // inclusive total through this round
        __syncthreads();
    }
    return s_carry;
}

// ---------------------------------------------------------------------------------------------
// phase 1a: per-view bins (once per theta)
// ---------------------------------------------------------------------------------------------
// viewmeta[v*8 + 1] = number of bins
__global__ void __launch_bounds__(256)
k_view_values(const uint8_t *__restrict__ sbox, const uint4 *__restrict__ frtab,
              const fr_t *__restrict__ chal, fr_t *__restrict__ val, fr_t *__restrict__ canon,
              uint32_t *__restrict__ viewmeta) {
    const int v = blockIdx.y;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= view_entries(v)) return;
    const fr_t theta = fr_ld(chal);
    const uint32_t trow = table_row_of_entry(v, e);
    fr_t acc = fr_from_tab(frtab, table_cell(sbox, 0, trow, ~0ull));
    for (int j = 1; j < v + 2; ++j)
        acc = fr_add(fr_mul(acc, theta), fr_from_tab(frtab, table_cell(sbox, j, trow, ~0ull)));
    fr_st(val + (size_t)v * BIN_STRIDE + e, acc);
    const fr_t c = fr_to_canonical(acc);
    fr_st(canon + (size_t)v * BIN_STRIDE + e, c);
}

// branch-free 256-bit compare of canonical values: *lt = (a < b) from one borrow chain, *eq from
// the OR of the difference words
__device__ __forceinline__ void fr_lt_eq(const fr_t &a, const fr_t &b, uint32_t *lt, uint32_t *eq) {
    uint32_t d0, d1, d2, d3, d4, d5, d6, d7, borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(d4), "=r"(d5), "=r"(d6), "=r"(d7),
          "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
          "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
          "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    *lt = borrow & 1u;  // 0xffffffff when a - b borrowed
    *eq = ((d0 | d1 | d2 | d3 | d4 | d5 | d6 | d7) == 0u) ? 1u : 0u;
}
// sort key = (canonical value, entry index): all keys are distinct, so equal values — a tiny theta
// makes thousands of table tuples compress to the same element — spread over the buckets too
__device__ __forceinline__ bool key_less(const fr_t &a, uint32_t ia, const fr_t &b, uint32_t ib) {
    uint32_t lt, eq;
    fr_lt_eq(a, b, &lt, &eq);
    return lt || (eq && ia < ib);
}

// Bucket boundaries from the data (sample sort): N_SPLIT entries taken at regular index intervals
// and sorted (k_splitters) by their (value, index) keys.  Whatever the
// distribution of the values — 256 consecutive integers per (tag, x) for a 3-column view, a few
// huge clusters for a theta below 2^64, everything equal for theta = 0 — a bucket then holds about
// entries / N_SPLIT keys, and the exact in-bucket ranking below stays linear overall.
constexpr uint32_t N_SPLIT = 2048;  // buckets = N_SPLIT + 1 <= N_BUCKETS
static_assert(N_SPLIT + 1 <= N_BUCKETS, "bucket arrays");
// the samples, contiguous (so that k_splitters reads them coalesced)
__global__ void __launch_bounds__(256)
k_sample(const fr_t *__restrict__ canon, fr_t *__restrict__ smp) {
    const int v = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N_SPLIT) return;
    const uint32_t e = (uint32_t)(((uint64_t)i * view_entries(v)) / N_SPLIT);
    fr_st(smp + (size_t)v * N_SPLIT + i, fr_ld(canon + (size_t)v * BIN_STRIDE + e));
}
// one warp per sample: its position among the N_SPLIT samples by counting (lane l compares with
// samples l, l+32, ...; ties — the same entry sampled twice when a view has fewer entries than
// samples — by sample number), then the key goes to that position.  4 M cheap compares spread over
// the whole GPU instead of a sorting network on one SM.
__global__ void __launch_bounds__(128)
k_splitters(const fr_t *__restrict__ smp, fr_t *__restrict__ spl_val, uint32_t *__restrict__ spl_idx) {
    const int v = blockIdx.y;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // my sample
    if (i >= N_SPLIT) return;
    const int lane = threadIdx.x & 31;
    const uint32_t np = view_entries(v);
    const fr_t *sv = smp + (size_t)v * N_SPLIT;
    const uint32_t ei = (uint32_t)(((uint64_t)i * np) / N_SPLIT);
    const fr_t me = fr_ld(sv + i);
    uint32_t less = 0;
#pragma unroll 4
    for (uint32_t j = lane; j < N_SPLIT; j += 32) {
        const uint32_t ej = (uint32_t)(((uint64_t)j * np) / N_SPLIT);
        uint32_t lt, eq;
        fr_lt_eq(fr_ldg_256_nc(sv + j), me, &lt, &eq);
        less += lt | (eq & ((ej < ei || (ej == ei && j < i)) ? 1u : 0u));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) less += __shfl_xor_sync(0xffffffffu, less, d);
    if (lane == 0) {
        fr_st(spl_val + (size_t)v * N_SPLIT + less, me);
        spl_idx[(size_t)v * N_SPLIT + less] = ei;
    }
}

// bucket of an entry = number of splitters <= its key; stored for the scatter and ranking kernels
__global__ void __launch_bounds__(256)
k_bucket_hist(const fr_t *__restrict__ canon, const fr_t *__restrict__ spl_val,
              const uint32_t *__restrict__ spl_idx, uint32_t *__restrict__ bkt,
              uint32_t *__restrict__ bhist) {
    const int v = blockIdx.y;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= view_entries(v)) return;
    const fr_t me = fr_ld(canon + (size_t)v * BIN_STRIDE + e);
    const fr_t *sv = spl_val + (size_t)v * N_SPLIT;
    const uint32_t *si = spl_idx + (size_t)v * N_SPLIT;
    uint32_t lo = 0, hi = N_SPLIT;  // first splitter > me
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key_less(me, e, fr_ld(sv + mid), si[mid])) hi = mid; else lo = mid + 1;
    }
    bkt[(size_t)v * BIN_STRIDE + e] = lo;
    atomicAdd(bhist + (size_t)v * (N_BUCKETS + 1) + lo, 1u);
}

__global__ void __launch_bounds__(1024) k_bucket_scan(uint32_t *__restrict__ bhist) {
    const int v = blockIdx.x;
    uint32_t *h = bhist + (size_t)v * (N_BUCKETS + 1);
    uint32_t total = block_exclusive_scan_u32(h, N_BUCKETS);
    if (threadIdx.x == 0) h[N_BUCKETS] = total;
}

// entries (and their keys and buckets) in bucket order, so that bucket mates are contiguous
__global__ void __launch_bounds__(256)
k_bucket_scatter(const fr_t *__restrict__ canon, const uint32_t *__restrict__ bkt,
                 const uint32_t *__restrict__ bhist, uint32_t *__restrict__ bcur,
                 uint32_t *__restrict__ slot, fr_t *__restrict__ canon_b,
                 uint32_t *__restrict__ bkt_b) {
    const int v = blockIdx.y;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= view_entries(v)) return;
    const fr_t me = fr_ld(canon + (size_t)v * BIN_STRIDE + e);
    const uint32_t b = bkt[(size_t)v * BIN_STRIDE + e];
    const uint32_t pos = bhist[(size_t)v * (N_BUCKETS + 1) + b] +
                         atomicAdd(bcur + (size_t)v * N_BUCKETS + b, 1u);
    slot[(size_t)v * BIN_STRIDE + pos] = e;
    bkt_b[(size_t)v * BIN_STRIDE + pos] = b;
    fr_st(canon_b + (size_t)v * BIN_STRIDE + pos, me);
}

// exact position of the entry at bucket slot p: bucket offset + number of bucket mates that order
// before it (ties broken by entry index).  The CTA stages the canonical values of every bucket its
// 256 slots touch in shared memory (window of RANK_WIN entries; anything beyond falls back to
// global loads — with splitters from the data a bucket holds ~32 entries, so that is rare).
constexpr uint32_t RANK_WIN = 1280;
__global__ void __launch_bounds__(256)
k_bucket_rank(const fr_t *__restrict__ canon_b, const uint32_t *__restrict__ bkt_b,
              const uint32_t *__restrict__ bhist, const uint32_t *__restrict__ slot,
              uint32_t *__restrict__ sorted) {
    __shared__ uint4 s_c[2 * RANK_WIN];
    const int v = blockIdx.y;
    const uint32_t np = view_entries(v);
    const uint32_t p0 = blockIdx.x * blockDim.x;
    if (p0 >= np) return;
    const fr_t *__restrict__ cb = canon_b + (size_t)v * BIN_STRIDE;
    const uint32_t *__restrict__ sl = slot + (size_t)v * BIN_STRIDE;
    const uint32_t *__restrict__ bb = bkt_b + (size_t)v * BIN_STRIDE;
    const uint32_t *__restrict__ bh = bhist + (size_t)v * (N_BUCKETS + 1);
    const uint32_t w_lo = bh[bb[p0]];
    const uint32_t w_hi = bh[bb[min(p0 + blockDim.x, np) - 1] + 1];
    const uint32_t w_n = min(w_hi - w_lo, RANK_WIN);
    for (uint32_t i = threadIdx.x; i < w_n; i += blockDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(cb + w_lo + i);
        s_c[2 * i] = src[0];
        s_c[2 * i + 1] = src[1];
    }
    __syncthreads();
    const uint32_t p = p0 + threadIdx.x;
    if (p >= np) return;
    const fr_t me = fr_ld(cb + p);
    const uint32_t e = sl[p];
    const uint32_t b = bb[p];
    const uint32_t lo = bh[b], hi = bh[b + 1];
    uint32_t rank = 0;
    for (uint32_t q = lo; q < hi; ++q) {
        const uint32_t wq = q - w_lo;  // lo >= w_lo by construction
        const fr_t other = wq < w_n ? fr_load_s(s_c + 2 * wq) : fr_ld(cb + q);
        uint32_t lt, eq;
        fr_lt_eq(other, me, &lt, &eq);
        rank += lt;
        if (eq && q != p && sl[q] < e) rank += 1;  // ties (equal values) ordered by entry index
    }
    sorted[(size_t)v * BIN_STRIDE + lo + rank] = e;
}

// block-wide sum of one u32 per thread (all threads get the result); blockDim <= 1024
__device__ __forceinline__ uint32_t block_sum_u32(uint32_t x, uint32_t *s_red /* [33] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) s_red[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t y = lane < nw ? s_red[lane] : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
        if (lane == 0) s_red[32] = y;
    }
    __syncthreads();
    const uint32_t r = s_red[32];
    __syncthreads();
    return r;
}

constexpr uint32_t HEAD_TILES = (VIEW_MAX_ENTRIES + 255) / 256;  // tiles of 256 sorted positions

// head[p] = 1 where the value at sorted position p differs from its predecessor (a new bin
// starts); tilesum[v][tile] = heads in the tile
__global__ void __launch_bounds__(256)
k_bin_heads(const fr_t *__restrict__ canon, const uint32_t *__restrict__ sorted,
            uint32_t *__restrict__ head, uint32_t *__restrict__ tilesum) {
    const int v = blockIdx.y;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const fr_t *cv = canon + (size_t)v * BIN_STRIDE;
    const uint32_t *sv = sorted + (size_t)v * BIN_STRIDE;
    uint32_t h = 0;
    if (p < view_entries(v)) {
        h = 1;
        if (p > 0) h = fr_eq(fr_ld(cv + sv[p]), fr_ld(cv + sv[p - 1])) ? 0u : 1u;
        head[(size_t)v * BIN_STRIDE + p] = h;
    }
    const int cnt = __syncthreads_count((int)h);
    if (threadIdx.x == 0) tilesum[v * HEAD_TILES + blockIdx.x] = (uint32_t)cnt;
}

// bin index of every sorted position = heads before it (tile prefix summed on the fly from the
// tile sums + ballot scan inside the tile); fills rank / dval / ts; CTA 0 records the bin count
__global__ void __launch_bounds__(256)
k_bin_fill(const fr_t *__restrict__ val, const uint32_t *__restrict__ sorted,
           const uint32_t *__restrict__ head, const uint32_t *__restrict__ tilesum,
           uint64_t usable, fr_t *__restrict__ dval, uint32_t *__restrict__ rank,
           uint32_t *__restrict__ ts, uint32_t *__restrict__ viewmeta) {
    __shared__ uint32_t s_red[33];
    __shared__ uint32_t s_wcnt[8];
    const int v = blockIdx.y;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t p = blockIdx.x * blockDim.x + t;
    const uint32_t np = view_entries(v);
    const uint32_t ntiles = (np + 255) / 256;
    if (blockIdx.x >= ntiles) return;
    const uint32_t *tsum = tilesum + v * HEAD_TILES;
    uint32_t before = 0, all = 0;
    for (uint32_t i = t; i < ntiles; i += blockDim.x) {
        const uint32_t x = tsum[i];
        all += x;
        if (i < blockIdx.x) before += x;
    }
    const uint32_t tile_prefix = block_sum_u32(before, s_red);
    if (blockIdx.x == 0) {
        const uint32_t total = block_sum_u32(all, s_red);
        if (t == 0) viewmeta[v * 8 + 1] = total;
    }
    const size_t o = (size_t)v * BIN_STRIDE;
    const uint32_t h = p < np ? head[o + p] : 0u;
    const uint32_t bal = __ballot_sync(0xffffffffu, h != 0);
    if (lane == 0) s_wcnt[wid] = __popc(bal);
    __syncthreads();
    uint32_t in_tile = __popc(bal & ((1u << lane) - 1u));
    for (uint32_t w = 0; w < wid; ++w) in_tile += s_wcnt[w];
    if (p >= np) return;
    const uint32_t e = sorted[o + p];
    const uint32_t k = tile_prefix + in_tile + h - 1;
    rank[o + e] = k;
    if (h) fr_st(dval + o + k, fr_ld(val + o + e));
    // table rows 66561..u-1 repeat row 0 (halo2 pads with the offset-0 value) -> entry 0
    uint32_t mult = entry_multiplicity(v, e) + (e == 0 ? (uint32_t)(usable - TABLE_ROWS) : 0u);
    atomicAdd(ts + o + k, mult);
}

// ---------------------------------------------------------------------------------------------
// phase 1b: per-lookup histogram of the rows over the bins
// ---------------------------------------------------------------------------------------------
// Also writes tracek[set][row] = op<<24 | bin of the row's tuple (0 for inactive rows), which the
// grand-product kernel reads instead of re-deriving the bin from the trace word.
__global__ void __launch_bounds__(256)
k_lookup_hist(Shape sh, const uint32_t *__restrict__ trace, const uint32_t *__restrict__ rank,
              uint32_t *__restrict__ cnt, uint32_t *__restrict__ tracek, RowMap rm) {
    const uint64_t n = sh.n();
    const int set = blockIdx.y;
    const int ops = sh.ops_per_set();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t usable = sh.usable();
    for (uint64_t lrow = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; lrow < rm.rows;
         lrow += stride) {
        const uint64_t row = rm.global(lrow);
        const uint32_t w = trace[(uint64_t)set * n + row];
        const uint32_t op = w >> 24;
        uint32_t tk = 0;
        if (op && row < usable) {
            const int v = view_of_op(op);
            const uint32_t e = entry_of_table_row(v, table_row_of_word(w));
            H2_ASSERT(e < view_entries(v));
            const uint32_t k = rank[(size_t)v * BIN_STRIDE + e];
            H2_ASSERT(k < BIN_STRIDE);
            atomicAdd(cnt + (size_t)(set * ops + (int)op - 1) * BIN_STRIDE + k, 1u);
            tk = (op << 24) | k;
        }
        tracek[(uint64_t)set * n + row] = tk;
    }
}

// Run structure of a lookup from its bin histogram: add the inactive rows to the zero bin, then
// compact the nonempty input bins (offz/nzk: first row and bin of each run of A') and the
// leftover table bins (loffz/lk).  Two kernels over tiles of LB_TILE bins: per-tile sums of
// (rows, nonempty bins, leftover rows, leftover bins), then each tile sums the tiles before it
// on the fly, scans its own bins and writes.  `active[l]` = rows on which the lookup's selector
// is enabled: input independent, computed on the host from the layout.
constexpr int LB_T = 256, LB_E = 4, LB_TILE = LB_T * LB_E;
constexpr uint32_t LB_TILES = (VIEW_MAX_ENTRIES + LB_TILE - 1) / LB_TILE;  // 66

struct BinItem {
    uint32_t ck, lf;
};
__device__ __forceinline__ BinItem bin_item(uint32_t k, uint32_t D, uint32_t k0, uint32_t zero_extra,
                                            const uint32_t *__restrict__ c,
                                            const uint32_t *__restrict__ tsv, uint32_t *bad) {
    BinItem it = {0u, 0u};
    if (k < D) {
        it.ck = c[k] + (k == k0 ? zero_extra : 0u);
        const uint32_t nzf = it.ck ? 1u : 0u;
        const uint32_t tk = tsv[k];
        if (tk < nzf) *bad = 1u;  // an input value without a table row
        it.lf = tk - nzf;
    }
    return it;
}

__global__ void __launch_bounds__(LB_T)
k_lookup_tiles(Shape sh, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ rank,
               const uint32_t *__restrict__ ts, const uint32_t *__restrict__ viewmeta,
               const uint32_t *__restrict__ active, uint4 *__restrict__ tsum,
               uint32_t *__restrict__ err) {
    __shared__ uint32_t s_red[33];
    const int l = blockIdx.y;
    const uint32_t op = (uint32_t)(l % sh.ops_per_set()) + 1;
    const int v = view_of_op(op);
    const uint32_t D = viewmeta[v * 8 + 1];
    const uint32_t k0 = rank[(size_t)v * BIN_STRIDE + (v == 0 ? 1280u : TROW_ZERO)];
    const uint32_t zero_extra = (uint32_t)sh.usable() - active[l];
    const uint32_t *__restrict__ c = cnt + (size_t)l * BIN_STRIDE;
    const uint32_t *__restrict__ tsv = ts + (size_t)v * BIN_STRIDE;
    const uint32_t kb = blockIdx.x * LB_TILE + threadIdx.x * LB_E;
    uint4 mine = make_uint4(0, 0, 0, 0);
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < LB_E; ++j) {
        const BinItem it = bin_item(kb + j, D, k0, zero_extra, c, tsv, &bad);
        mine.x += it.ck;
        mine.y += it.ck ? 1u : 0u;
        mine.z += it.lf;
        mine.w += it.lf ? 1u : 0u;
    }
    uint4 tot;
    tot.x = block_sum_u32(mine.x, s_red);
    tot.y = block_sum_u32(mine.y, s_red);
    tot.z = block_sum_u32(mine.z, s_red);
    tot.w = block_sum_u32(mine.w, s_red);
    if (threadIdx.x == 0) tsum[(size_t)l * LB_TILES + blockIdx.x] = tot;
    if (bad) atomicOr(err + 1, 1u);
}

__global__ void __launch_bounds__(LB_T)
k_lookup_compact(Shape sh, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ rank,
                 const uint32_t *__restrict__ ts, const uint32_t *__restrict__ viewmeta,
                 const uint32_t *__restrict__ active, const uint4 *__restrict__ tsum,
                 uint32_t *__restrict__ offz, uint32_t *__restrict__ nzk,
                 uint32_t *__restrict__ loffz, uint32_t *__restrict__ lk,
                 LookupMeta *__restrict__ meta, uint32_t *__restrict__ err) {
    __shared__ uint32_t s_red[33];
    __shared__ uint4 s_warp[LB_T / 32];
    const int l = blockIdx.y;
    const uint32_t op = (uint32_t)(l % sh.ops_per_set()) + 1;
    const int v = view_of_op(op);
    const uint32_t D = viewmeta[v * 8 + 1];
    const uint32_t usable = (uint32_t)sh.usable();
    const uint32_t k0 = rank[(size_t)v * BIN_STRIDE + (v == 0 ? 1280u : TROW_ZERO)];
    const uint32_t zero_extra = usable - active[l];
    const uint32_t *__restrict__ c = cnt + (size_t)l * BIN_STRIDE;
    const uint32_t *__restrict__ tsv = ts + (size_t)v * BIN_STRIDE;
    uint32_t *__restrict__ o_offz = offz + (size_t)l * BIN_STRIDE;
    uint32_t *__restrict__ o_nzk = nzk + (size_t)l * BIN_STRIDE;
    uint32_t *__restrict__ o_loffz = loffz + (size_t)l * BIN_STRIDE;
    uint32_t *__restrict__ o_lk = lk + (size_t)l * BIN_STRIDE;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // sums of the tiles before this one (LB_TILES <= blockDim: one tile per thread)
    uint4 before = make_uint4(0, 0, 0, 0);
    if (t < blockIdx.x) before = tsum[(size_t)l * LB_TILES + t];
    uint4 carry;
    carry.x = block_sum_u32(before.x, s_red);
    carry.y = block_sum_u32(before.y, s_red);
    carry.z = block_sum_u32(before.z, s_red);
    carry.w = block_sum_u32(before.w, s_red);
    const uint32_t kb = blockIdx.x * LB_TILE + t * LB_E;
    BinItem it[LB_E];
    uint4 mine = make_uint4(0, 0, 0, 0);
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < LB_E; ++j) {
        it[j] = bin_item(kb + j, D, k0, zero_extra, c, tsv, &bad);
        mine.x += it[j].ck;
        mine.y += it[j].ck ? 1u : 0u;
        mine.z += it[j].lf;
        mine.w += it[j].lf ? 1u : 0u;
    }
    uint4 inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint4 y;
        y.x = __shfl_up_sync(0xffffffffu, inc.x, d);
        y.y = __shfl_up_sync(0xffffffffu, inc.y, d);
        y.z = __shfl_up_sync(0xffffffffu, inc.z, d);
        y.w = __shfl_up_sync(0xffffffffu, inc.w, d);
        if ((int)lane >= d) { inc.x += y.x; inc.y += y.y; inc.z += y.z; inc.w += y.w; }
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint4 ex = make_uint4(carry.x + inc.x - mine.x, carry.y + inc.y - mine.y,
                          carry.z + inc.z - mine.z, carry.w + inc.w - mine.w);
    for (uint32_t w = 0; w < wid; ++w) {
        const uint4 x = s_warp[w];
        ex.x += x.x; ex.y += x.y; ex.z += x.z; ex.w += x.w;
    }
#pragma unroll
    for (int j = 0; j < LB_E; ++j) {
        if (it[j].ck) {
            o_offz[ex.y] = ex.x;
            o_nzk[ex.y] = kb + j;
            ex.x += it[j].ck;
            ex.y += 1;
        }
        if (it[j].lf) {
            o_loffz[ex.w] = ex.z;
            o_lk[ex.w] = kb + j;
            ex.z += it[j].lf;
            ex.w += 1;
        }
    }
    // the thread holding the last bin knows the totals
    if (kb < D && kb + LB_E >= D) {
        LookupMeta m;
        m.n_nonempty = ex.y;
        m.n_left = ex.w;
        m.repeated = ex.z;
        m.k_zero = k0;
        m.k_pad = rank[(size_t)v * BIN_STRIDE];  // bin of table row 0 = the padding value
        // totals: rows must equal usable, leftover must equal usable - nonempty
        m.bad = (ex.x != usable || ex.z != usable - ex.y) ? 2u : 0u;
        m.pad[0] = m.pad[1] = 0;
        meta[l] = m;
        if (m.bad) atomicOr(err + 1, 1u);
        o_offz[ex.y] = usable;
        o_loffz[ex.w] = ex.z;
    }
}

// largest j in [0, cnt) with arr[j] <= x   (arr ascending, arr[0] == 0 <= x)
__device__ __forceinline__ uint32_t search_le(const uint32_t *__restrict__ arr, uint32_t cnt,
                                              uint32_t x) {
    uint32_t lo = 0, hi = cnt;  // invariant: arr[lo] <= x, (hi == cnt or arr[hi] > x)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (arr[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}
// the same inside [lo, hi): the caller guarantees arr[lo] <= x and (hi == cnt or arr[hi] > x)
__device__ __forceinline__ uint32_t search_le_in(const uint32_t *__restrict__ arr, uint32_t lo,
                                                 uint32_t hi, uint32_t x) {
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (arr[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// Coarse indices into the run tables, one entry per RUN_STEP rows, so that the per-row searches
// of k_permuted look at <= RUN_STEP+1 neighbouring entries (L1-resident) instead of bisecting the
// whole table through L2:  jidx[t] = run of A' containing row t*RUN_STEP,  lidx[t] = leftover
// entry containing leftover position t*RUN_STEP.
constexpr uint32_t RUN_SHIFT = 8, RUN_STEP = 1u << RUN_SHIFT;
__global__ void __launch_bounds__(256)
k_run_index(Shape sh, const LookupMeta *__restrict__ meta, const uint32_t *__restrict__ offz_all,
            const uint32_t *__restrict__ loffz_all, uint32_t *__restrict__ jidx_all,
            uint32_t *__restrict__ lidx_all, uint32_t stride) {
    const int l = blockIdx.y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    const LookupMeta m = meta[l];
    const uint32_t usable = (uint32_t)sh.usable();
    const uint64_t pos = (uint64_t)t << RUN_SHIFT;
    const uint32_t i = pos < usable ? (uint32_t)pos : usable - 1;
    jidx_all[(size_t)l * stride + t] = search_le(offz_all + (size_t)l * BIN_STRIDE, m.n_nonempty, i);
    uint32_t li = 0;
    if (m.repeated) {
        const uint32_t jj = pos < m.repeated ? (uint32_t)pos : m.repeated - 1;
        li = search_le(loffz_all + (size_t)l * BIN_STRIDE, m.n_left, jj);
    }
    lidx_all[(size_t)l * stride + t] = li;
}

struct RowBins {
    uint32_t k, kp, cls;
};
__device__ __forceinline__ RowBins row_bins(uint32_t i, const LookupMeta &m,
                                            const uint32_t *__restrict__ offz,
                                            const uint32_t *__restrict__ nzk,
                                            const uint32_t *__restrict__ loffz,
                                            const uint32_t *__restrict__ lk,
                                            const uint32_t *__restrict__ jidx,
                                            const uint32_t *__restrict__ lidx) {
    RowBins rb;
    const uint32_t t = i >> RUN_SHIFT;
    const uint32_t j = search_le_in(offz, jidx[t], jidx[t + 1] + 1, i);
    H2_ASSERT(j < m.n_nonempty && offz[j] <= i && i < offz[j + 1]);
    rb.k = nzk[j];
    if (offz[j] == i) {
        rb.kp = rb.k;  // first occurrence: S' = A'
        rb.cls = CODE_FIRST;
    } else {
        // repeated rows, in row order, receive the leftover multiset from its END
        const uint32_t rho = i - (j + 1);
        const uint32_t jj = m.repeated - 1 - rho;
        const uint32_t tt = jj >> RUN_SHIFT;
        H2_ASSERT(rho < m.repeated);
        const uint32_t le = search_le_in(loffz, lidx[tt], lidx[tt + 1] + 1, jj);
        H2_ASSERT(le < m.n_left && loffz[le] <= jj && jj < loffz[le + 1]);
        rb.kp = lk[le];
        rb.cls = rb.kp == m.k_pad ? CODE_PAD : CODE_OTHER;
    }
    return rb;
}

// ---------------------------------------------------------------------------------------------
// phase 1c: expand A, S, A', S' (one lane per row, 256-bit stores).  Also records per row the
// 4-byte code (bin of A' | class<<30) — and the bin of S' where it is neither A' nor the
// padding value — so that the grand product does not repeat the searches.  With Ap == null only
// the codes are produced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_permuted(Shape sh, int first_lookup, const uint32_t *__restrict__ trace,
           const uint32_t *__restrict__ tracek, const fr_t *__restrict__ val,
           const fr_t *__restrict__ dval, const uint32_t *__restrict__ offz_all,
           const uint32_t *__restrict__ nzk_all, const uint32_t *__restrict__ loffz_all,
           const uint32_t *__restrict__ lk_all, const LookupMeta *__restrict__ meta,
           uint64_t *__restrict__ A, uint64_t *__restrict__ S, uint64_t *__restrict__ Ap,
           uint64_t *__restrict__ Sp, uint32_t *__restrict__ fcode, uint32_t *__restrict__ codep,
           uint32_t *__restrict__ zmask, uint64_t zm, const uint32_t *__restrict__ jidx_all,
           const uint32_t *__restrict__ lidx_all, uint32_t idx_stride, RowMap rm) {
    const int li = blockIdx.y;            // index inside this call
    const int l = first_lookup + li;      // lookup index in the circuit
    const int ops = sh.ops_per_set();
    const int set = l / ops;
    const uint32_t op = (uint32_t)(l % ops) + 1;
    const int v = view_of_op(op);
    const uint64_t n = sh.n(), usable = sh.usable();
    const uint64_t out_rows = rm.rows;
    const LookupMeta m = meta[l];
    const uint32_t *offz = offz_all + (size_t)l * BIN_STRIDE;
    const uint32_t *nzk = nzk_all + (size_t)l * BIN_STRIDE;
    const uint32_t *loffz = loffz_all + (size_t)l * BIN_STRIDE;
    const uint32_t *lk = lk_all + (size_t)l * BIN_STRIDE;
    const fr_t *valv = val + (size_t)v * BIN_STRIDE;
    const fr_t *dv = dval + (size_t)v * BIN_STRIDE;
    uint32_t *__restrict__ fc = fcode + (size_t)l * n;
    uint32_t *__restrict__ zmk = zmask + (size_t)l * zm;
    uint32_t *__restrict__ gmk = zmask + ((size_t)sh.n_lookups() + l) * zm;  // rows for k_z_general
    uint32_t *__restrict__ cp = codep + (size_t)l * n;
    const uint32_t *__restrict__ tk = tracek + (uint64_t)set * n;
    const uint32_t *__restrict__ jidx = jidx_all + (size_t)l * idx_stride;
    const uint32_t *__restrict__ lidx = lidx_all + (size_t)l * idx_stride;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const fr_t zero = fr_zero();
    // row ranges are multiples of 32 and a warp's 32 lanes sit on 32 consecutive rows, so one
    // ballot per iteration is the non-constant mask of one chunk of the grand product
    // (ownership tiles and shards are multiples of the CTA's 256 rows: one local -> global
    // translation per CTA iteration, in the uniform datapath, instead of a division per row)
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < out_rows; base += stride) {
        const uint64_t lrow = base + threadIdx.x;
        if (lrow >= out_rows) break;  // (never splits a warp: out_rows is a multiple of 32)
        const uint64_t row = rm.global(base) + threadIdx.x;
        const size_t o = ((size_t)li * out_rows + lrow) * 4;
        uint32_t code = fcode_pack(FK_ONE, 0, 0);
        if (row >= usable) {
            if (A) fr_store_256_cs(A + o, zero);
            if (S) fr_store_256_cs(S + o, zero);
            if (Ap) {
                fr_store_256_cs(Ap + o, zero);
                fr_store_256_cs(Sp + o, zero);
            }
        } else {
            if (A) {
                const uint32_t w = trace[(uint64_t)set * n + row];
                fr_t a = zero;
                if ((w >> 24) == op) a = fr_ld(valv + entry_of_table_row(v, table_row_of_word(w)));
                fr_store_256_cs(A + o, a);
            }
            if (S) {
                const uint32_t e = row < TABLE_ROWS ? entry_of_table_row(v, (uint32_t)row) : 0u;
                fr_store_256_cs(S + o, fr_ld(valv + e));
            }
            const RowBins rb = row_bins((uint32_t)row, m, offz, nzk, loffz, lk, jidx, lidx);
            H2_ASSERT(rb.k < BIN_STRIDE && rb.kp < BIN_STRIDE && row < n);
            if (rb.cls == CODE_OTHER) cp[row] = rb.kp;  // (k_z_general reads it for these rows)
            if (Ap) {
                fr_store_256_cs(Ap + o, fr_ldg_256_nc(dv + rb.k));
                fr_store_256_cs(Sp + o, fr_ldg_256_nc(dv + rb.kp));
            }
            // classify the row's grand-product factor (see fcode_pack)
            const uint32_t t = tk[row];
            const bool active = (t >> 24) == op;
            const bool in_zero_run = rb.k == m.k_zero && rb.cls == CODE_PAD;
            if (row < TABLE_ROWS) {
                code = (!active && in_zero_run)
                           ? fcode_pack(FK_TABLE, FSEL_SGZ, entry_of_table_row(v, (uint32_t)row))
                           : fcode_pack(FK_TABLE_ROW, rb.cls, rb.k);
            } else if (in_zero_run) {
                if (active) code = fcode_pack(FK_TABLE, FSEL_NACP, t & 0xffffffu);
            } else if (!active && rb.cls != CODE_OTHER) {
                code = fcode_pack(FK_TABLE, rb.cls == CODE_FIRST ? FSEL_IBGZ : FSEL_IBPZ, rb.k);
            } else {
                code = fcode_pack(FK_GENERAL, rb.cls, rb.k);
            }
        }
        fc[row] = code;
        const int lane = threadIdx.x & 31;
        const uint32_t bal = __ballot_sync(0xffffffffu, (code >> 30) != FK_ONE);
        if (lane == 0) zmk[lrow >> 5] = bal;
        // rows whose factor is not a table entry: one more mask, compacted into a list by
        // k_general_list before k_z_general evaluates them densely
        const uint32_t gbal = __ballot_sync(0xffffffffu, (code >> 30) >= FK_GENERAL);
        if (lane == 0) gmk[lrow >> 5] = gbal;
    }
}

// rows flagged in the second mask -> dense list of local row indices per lookup (one atomic per
// warp = per 1024 rows; the order of the entries is irrelevant)
__global__ void __launch_bounds__(256)
k_general_list(int first_lookup, int n_lookups, const uint32_t *__restrict__ zmask, uint64_t zm,
               uint64_t words, uint32_t *__restrict__ glist, const uint32_t *__restrict__ gbase,
               uint32_t *__restrict__ gcount, uint32_t *__restrict__ err) {
    const int l = first_lookup + blockIdx.y;
    const uint32_t *__restrict__ gmk = zmask + ((size_t)n_lookups + l) * zm;
    uint32_t *__restrict__ gl = glist + gbase[l];
    const uint32_t gcap = gbase[l + 1] - gbase[l];
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); w0 < words; w0 += stride) {
        uint32_t m = w0 + lane < words ? gmk[w0 + lane] : 0u;
        const uint32_t cnt = __popc(m);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (!total) continue;
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(gcount + l, total);
        pos = __shfl_sync(0xffffffffu, pos, 0) + inc - cnt;
        const uint32_t row0 = (uint32_t)((w0 + lane) << 5);
        for (; m; m &= m - 1u, ++pos) {
            H2_ASSERT(pos < gcap);
            if (pos < gcap) gl[pos] = row0 + (uint32_t)__ffs(m) - 1u;
            else atomicOr(err + 2, 1u);  // cannot happen: gcap bounds the count (api.cu)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// phase 2a: per-bin tables for the grand product (once per beta,gamma)
//   ib[k] = 1/(d_k + beta), ig[k] = 1/(d_k + gamma): Montgomery-trick batch inversion with ONE
//   field inversion per CTA of INV_TILE bins (block-wide forward + reverse product scans);
//   nac[k] = (d_k + beta) * (pad + gamma) with pad = value of table row 0 (numerator of every
//   row past the table);  then k_inv_merge: ibg[k] = ib[k]*ig[k] (first occurrence: S' = A'),
//   ibp[k] = ib[k]*ig[k_pad] (repeated row that received the padding value).
// ---------------------------------------------------------------------------------------------
constexpr int INV_T = 256;
constexpr int INV_CH = 8;
constexpr int INV_TILE = INV_T * INV_CH;

__global__ void __launch_bounds__(INV_T)
k_inv_tables(const fr_t *__restrict__ chal, const fr_t *__restrict__ dval,
             const fr_t *__restrict__ val, const uint32_t *__restrict__ viewmeta,
             fr_t *__restrict__ ib, fr_t *__restrict__ ig, fr_t *__restrict__ nac,
             uint32_t *__restrict__ err) {
    __shared__ fr_t s_wtot[INV_T / 32];
    __shared__ fr_t s_inv;
    const int v = blockIdx.y, which = blockIdx.z;
    const uint32_t D = viewmeta[v * 8 + 1];
    const uint32_t base = blockIdx.x * INV_TILE;
    if (base >= D) return;  // whole CTA
    const size_t o = (size_t)v * BIN_STRIDE;
    const fr_t ch = fr_ld(chal + 1 + which);
    const fr_t padg = fr_add(fr_ld(val + o), fr_ld(chal + 2));
    fr_t *out = which == 0 ? ib : ig;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t k0 = base + (uint32_t)t * INV_CH;
    const fr_t one = fr_one();

    fr_t x[INV_CH], pre[INV_CH];  // pre[j] = x[0] * ... * x[j-1]
    uint32_t zero_mask = 0;
    fr_t tot = one;
#pragma unroll
    for (int j = 0; j < INV_CH; ++j) {
        const uint32_t k = k0 + j;
        fr_t xv = one;
        if (k < D) {
            xv = fr_add(fr_ld(dval + o + k), ch);
            if (which == 0) fr_st(nac + o + k, fr_mul(xv, padg));
            if (fr_is_zero(xv)) {  // halo2's batch_invert would leave a zero here
                zero_mask |= 1u << j;
                xv = one;
                atomicOr(err, 1u);
            }
        }
        x[j] = xv;
        pre[j] = tot;
        tot = fr_mul(tot, xv);
    }
    // forward / reverse inclusive product scans of the thread totals inside the warp
    fr_t finc = tot, rinc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const fr_t yu = fr_shfl_up(finc, d);
        const fr_t yd = fr_shfl_down(rinc, d);
        if (lane >= d) finc = fr_mul(finc, yu);
        if (lane + d < 32) rinc = fr_mul(rinc, yd);
    }
    if (lane == 31) s_wtot[wid] = finc;
    __syncthreads();
    fr_t before = fr_shfl_up(finc, 1), after = fr_shfl_down(rinc, 1);
    if (lane == 0) before = one;
    if (lane == 31) after = one;
    for (int w = 0; w < INV_T / 32; ++w) {
        const fr_t wt = s_wtot[w];
        if (w < wid) before = fr_mul(before, wt);
        if (w > wid) after = fr_mul(after, wt);
    }
    if (wid == 0) {  // one inversion per CTA
        fr_t all = s_wtot[0];
        for (int w = 1; w < INV_T / 32; ++w) all = fr_mul(all, s_wtot[w]);
        const fr_t inv = fr_inv(all);
        if (lane == 0) s_inv = inv;
    }
    __syncthreads();
    // 1 / (product of this thread's elements) = inv_all * (everything before) * (everything after)
    fr_t run = fr_mul(fr_mul(s_inv, before), after);
#pragma unroll
    for (int j = INV_CH - 1; j >= 0; --j) {
        const uint32_t k = k0 + j;
        const fr_t r = fr_mul(run, pre[j]);
        run = fr_mul(run, x[j]);
        if (k < D) fr_st(out + o + k, ((zero_mask >> j) & 1u) ? fr_zero() : r);
    }
}

__global__ void __launch_bounds__(256)
k_inv_merge(const uint32_t *__restrict__ viewmeta, const uint32_t *__restrict__ rank,
            const fr_t *__restrict__ ib, const fr_t *__restrict__ ig, const fr_t *__restrict__ nac,
            fr_t *__restrict__ ibg, fr_t *__restrict__ ibp, fr_t *__restrict__ nacp,
            fr_t *__restrict__ ibgz, fr_t *__restrict__ ibpz, const fr_t *__restrict__ val,
            const fr_t *__restrict__ chal, fr_t *__restrict__ sgz) {
    const int v = blockIdx.y;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t o = (size_t)v * BIN_STRIDE;
    const uint32_t k_pad = rank[o];  // bin of table row 0, the padding value
    const uint32_t k_zero = rank[o + (v == 0 ? 1280u : TROW_ZERO)];  // bin of the all-zero tuple
    const fr_t ig_pad = fr_ld(ig + o + k_pad);
    // an inactive row inside the table rows AND inside the zero run of A' (S' = padding value):
    // (0+beta)(S+gamma) / ((0+beta)(pad+gamma)) = (S+gamma)/(pad+gamma), per table entry
    if (k < view_entries(v)) fr_st(sgz + o + k, fr_mul(fr_add(fr_ld(val + o + k), fr_ld(chal + 2)), ig_pad));
    if (k >= viewmeta[v * 8 + 1]) return;
    const fr_t b = fr_ld(ib + o + k);
    const fr_t g = fr_mul(b, fr_ld(ig + o + k));
    const fr_t p = fr_mul(b, ig_pad);
    fr_st(ibg + o + k, g);
    fr_st(ibp + o + k, p);
    // one operand constant: an active row inside the zero run of A' (denominator ibp[k_zero]),
    // an inactive row outside it (numerator nac[k_zero]) -> the factor is a table entry
    const fr_t ibp_zero = fr_mul(fr_ld(ib + o + k_zero), ig_pad);
    const fr_t nac_zero = fr_ld(nac + o + k_zero);
    fr_st(nacp + o + k, fr_mul(fr_ld(nac + o + k), ibp_zero));
    fr_st(ibgz + o + k, fr_mul(nac_zero, g));
    fr_st(ibpz + o + k, fr_mul(nac_zero, p));
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

cudaError_t launch_view_bins(h2aes_circuit *c) {
    h2aes_ctx *ctx = c->ctx;
    cudaStream_t s = ctx->stream;
    cudaError_t e;
    if ((e = cudaMemsetAsync(c->d_viewmeta, 0, N_VIEWS * 8 * 4, s))) return e;
    if ((e = cudaMemsetAsync(c->d_bhist, 0, (size_t)N_VIEWS * (N_BUCKETS + 1) * 4, s))) return e;
    if ((e = cudaMemsetAsync(c->d_bcur, 0, (size_t)N_VIEWS * N_BUCKETS * 4, s))) return e;
    if ((e = cudaMemsetAsync(c->d_ts, 0, (size_t)N_VIEWS * BIN_STRIDE * 4, s))) return e;
    dim3 g(cdiv(VIEW_MAX_ENTRIES, 256), N_VIEWS);
    prof_begin(ctx, "k_view_values");
    k_view_values<<<g, 256, 0, s>>>(ctx->d_sbox, ctx->d_frtab, c->d_chal, c->d_val, c->d_canon,
                                    c->d_viewmeta);
    prof_end(ctx);
    fr_t *smp = c->d_splval + (size_t)N_VIEWS * N_SPLIT;  // (second half of the allocation)
    prof_begin(ctx, "k_sample");
    k_sample<<<dim3(N_SPLIT / 256, N_VIEWS), 256, 0, s>>>(c->d_canon, smp);
    prof_end(ctx);
    prof_begin(ctx, "k_splitters");
    k_splitters<<<dim3(N_SPLIT / 4, N_VIEWS), 128, 0, s>>>(smp, c->d_splval, c->d_splidx);
    prof_end(ctx);
    prof_begin(ctx, "k_bucket_hist");
    k_bucket_hist<<<g, 256, 0, s>>>(c->d_canon, c->d_splval, c->d_splidx, c->d_bkt, c->d_bhist);
    prof_end(ctx);
    prof_begin(ctx, "k_bucket_scan");
    k_bucket_scan<<<N_VIEWS, 1024, 0, s>>>(c->d_bhist);
    prof_end(ctx);
    prof_begin(ctx, "k_bucket_scatter");
    k_bucket_scatter<<<g, 256, 0, s>>>(c->d_canon, c->d_bkt, c->d_bhist, c->d_bcur, c->d_slot,
                                       c->d_canon_b, c->d_bkt_b);
    prof_end(ctx);
    prof_begin(ctx, "k_bucket_rank");
    k_bucket_rank<<<g, 256, 0, s>>>(c->d_canon_b, c->d_bkt_b, c->d_bhist, c->d_slot, c->d_sorted);
    prof_end(ctx);
    // (d_slot is free again after k_bucket_rank: reused for the per-tile head counts)
    prof_begin(ctx, "k_bin_heads");
    k_bin_heads<<<g, 256, 0, s>>>(c->d_canon, c->d_sorted, c->d_head, c->d_slot);
    prof_end(ctx);
    prof_begin(ctx, "k_bin_fill");
    k_bin_fill<<<g, 256, 0, s>>>(c->d_val, c->d_sorted, c->d_head, c->d_slot, c->sh.usable(),
                                 c->d_dval, c->d_rank, c->d_ts, c->d_viewmeta);
    prof_end(ctx);
    ctx->launches += 9;
    return cudaGetLastError();
}

cudaError_t launch_lookup_hist(h2aes_circuit *c, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    cudaStream_t s = ctx->stream;
    const int L = c->sh.n_lookups();
    cudaError_t e;
    if ((e = cudaMemsetAsync(c->d_cnt, 0, (size_t)L * BIN_STRIDE * 4, s))) return e;
    dim3 g(ctx->sm_count * 2, c->sh.n_sets());
    prof_begin(ctx, "k_lookup_hist");
    k_lookup_hist<<<g, 256, 0, s>>>(c->sh, c->d_trace, c->d_rank, c->d_cnt, c->d_tracek, rm);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lookup_bins(h2aes_circuit *c) {
    h2aes_ctx *ctx = c->ctx;
    const int L = c->sh.n_lookups();
    dim3 g(LB_TILES, L);
    prof_begin(ctx, "k_lookup_tiles");
    k_lookup_tiles<<<g, LB_T, 0, ctx->stream>>>(c->sh, c->d_cnt, c->d_rank, c->d_ts, c->d_viewmeta,
                                                c->d_active, c->d_lbsum, c->d_err);
    prof_end(ctx);
    prof_begin(ctx, "k_lookup_compact");
    k_lookup_compact<<<g, LB_T, 0, ctx->stream>>>(c->sh, c->d_cnt, c->d_rank, c->d_ts, c->d_viewmeta,
                                                  c->d_active, c->d_lbsum, c->d_offz, c->d_nzk,
                                                  c->d_loffz, c->d_lk, c->d_meta, c->d_err);
    prof_end(ctx);
    prof_begin(ctx, "k_run_index");
    k_run_index<<<dim3(cdiv(c->idx_stride, 256), L), 256, 0, ctx->stream>>>(
        c->sh, c->d_meta, c->d_offz, c->d_loffz, c->d_jidx, c->d_lidx, (uint32_t)c->idx_stride);
    prof_end(ctx);
    ctx->launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_permuted(h2aes_circuit *c, int first, int count, uint64_t *A, uint64_t *S,
                            uint64_t *Ap, uint64_t *Sp, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    const uint64_t rows = rm.rows;
    unsigned gx = cdiv(rows, 256);
    // H2AES_STREAM_CTAS_PER_SM: total resident CTAs per SM of the streaming kernels (all lookups
    // together); with H2AES_STREAM_PAD_KB this keeps them to a share of every SM
    static int per_sm = -1;
    if (per_sm < 0) {
        const char *s = getenv("H2AES_STREAM_CTAS_PER_SM");
        per_sm = s ? atoi(s) : 0;
    }
    unsigned cap = (unsigned)ctx->sm_count * 8;
    if (per_sm) cap = std::max(1u, (unsigned)(ctx->sm_count * per_sm) / (unsigned)count);
    // H2AES_PERM_WAVES: all lookups together fill the machine this many times (0: every lookup
    // gets a machine-filling grid of its own, i.e. `count` waves of short-lived CTAs)
    static int waves = -1;
    if (waves < 0) {
        const char *s = getenv("H2AES_PERM_WAVES");
        waves = s ? atoi(s) : 0;
    }
    if (waves) cap = std::max(1u, (unsigned)(ctx->sm_count * 8 * waves) / (unsigned)count);
    if (gx > cap) gx = cap;
    dim3 g(gx, count);
    cudaError_t e0 = cudaMemsetAsync(c->d_gcount + first, 0, (size_t)count * 4, ctx->stream);
    if (e0) return e0;
    const size_t pad = stream_pad_bytes((const void *)k_permuted);
    prof_begin(ctx, "k_permuted");
    k_permuted<<<g, 256, pad, ctx->stream>>>(c->sh, first, c->d_trace, c->d_tracek, c->d_val, c->d_dval,
                                           c->d_offz, c->d_nzk, c->d_loffz, c->d_lk, c->d_meta, A, S,
                                           Ap, Sp, c->d_codek, c->d_codep, c->d_zmask, c->zm,
                                           c->d_jidx, c->d_lidx, (uint32_t)c->idx_stride, rm);
    prof_end(ctx);
    const uint64_t words = (rows + 31) >> 5;
    unsigned lx = cdiv(words, 256);
    if (lx > 64) lx = 64;
    prof_begin(ctx, "k_general_list");
    k_general_list<<<dim3(lx, count), 256, 0, ctx->stream>>>(first, c->sh.n_lookups(), c->d_zmask, c->zm,
                                                               words, c->d_glist, c->d_gbase,
                                                               c->d_gcount, c->d_err);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_inv_tables(h2aes_circuit *c) {
    h2aes_ctx *ctx = c->ctx;
    dim3 g(cdiv(VIEW_MAX_ENTRIES, INV_TILE), N_VIEWS, 2);
    prof_begin(ctx, "k_inv_tables");
    k_inv_tables<<<g, INV_T, 0, ctx->stream>>>(c->d_chal, c->d_dval, c->d_val, c->d_viewmeta,
                                               c->d_ib, c->d_ig, c->d_nac, c->d_err);
    prof_end(ctx);
    dim3 g2(cdiv(VIEW_MAX_ENTRIES, 256), N_VIEWS);
    prof_begin(ctx, "k_inv_merge");
    k_inv_merge<<<g2, 256, 0, ctx->stream>>>(c->d_viewmeta, c->d_rank, c->d_ib, c->d_ig, c->d_nac,
                                             c->d_ibg, c->d_ibp, c->d_nacp, c->d_ibgz, c->d_ibpz,
                                             c->d_val, c->d_chal, c->d_sgz);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

__global__ void k_dbg_fr_mul(const fr_t *__restrict__ a, const fr_t *__restrict__ b, uint64_t n,
                             fr_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_st(out + i, fr_mul(fr_ld(a + i), fr_ld(b + i)));
}
cudaError_t launch_dbg_fr_mul(h2aes_ctx *ctx, const fr_t *a, const fr_t *b, uint64_t n, fr_t *out) {
    prof_begin(ctx, "k_dbg_fr_mul");
    k_dbg_fr_mul<<<cdiv(n, 128), 128, 0, ctx->stream>>>(a, b, n, out);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// permutation (copy-constraint) argument — SURVEY 8(f) item 1
//   halo2_proofs v0.3.0 [UPSTREAM] plonk/permutation: sigma_p(omega^i) = delta^{p'} omega^{i'}
//   with (p', i') = mapping[p][i] (keygen), and per chunk of 3 columns the grand product
//     Z[i+1] = Z[i] * prod_t (v_t(i) + beta*delta^{p_t}*omega^i + gamma)
//                   / prod_t (v_t(i) + beta*sigma_{p_t}(omega^i) + gamma),
//   chunk s starting from the previous chunk's Z[n-6].
//   A term whose cell maps to itself cancels; a row where all terms cancel has factor one.  The
//   other rows: k_p_terms (numerator, denominator), k_p_invert (Montgomery-trick batch inversion,
//   one field inversion per CTA), then the same chunk-total / scan / propagate kernels as the
//   lookup grand product.
// ---------------------------------------------------------------------------------------------
constexpr int ZI = 32;  // rows per mask word of the permutation product
constexpr uint32_t W_LO_BITS = 10, W_LO = 1u << W_LO_BITS;  // omega^j = w_hi[j >> 10] * w_lo[j & 1023]
constexpr uint32_t PMAP_ROW_BITS = 24;                       // packed mapping: position << 24 | row

// w_lo[t] = omega^t (t < 1024), w_hi[b] = omega^(1024 b) (b < n/1024), dpow[p] = delta^p
__global__ void __launch_bounds__(256)
k_p_tables(fr_t omega, fr_t delta, uint32_t n_hi, int n_cols, fr_t *__restrict__ w_lo,
           fr_t *__restrict__ w_hi, fr_t *__restrict__ dpow) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    auto pow_u32 = [](fr_t base, uint32_t e) {
        fr_t acc = fr_one();
        for (; e; e >>= 1) {
            if (e & 1u) acc = fr_mul(acc, base);
            base = fr_mul(base, base);
        }
        return acc;
    };
    if (t < W_LO) fr_st(w_lo + t, pow_u32(omega, t));
    if (t < n_hi) {
        fr_t step = omega;  // omega^1024
        for (uint32_t i = 0; i < W_LO_BITS; ++i) step = fr_mul(step, step);
        fr_st(w_hi + t, pow_u32(step, t));
    }
    if (t < (uint32_t)n_cols) fr_st(dpow + t, pow_u32(delta, t));
}

// bdp[p] = beta * delta^p
__global__ void k_p_beta(const fr_t *__restrict__ chal, const fr_t *__restrict__ dpow, int n_cols,
                         fr_t *__restrict__ bdp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_cols) fr_st(bdp + p, fr_mul(fr_ld(chal + 1), fr_ld(dpow + p)));
}

__device__ __forceinline__ fr_t omega_pow(const fr_t *__restrict__ w_lo, const fr_t *__restrict__ w_hi,
                                          uint32_t j) {
    return fr_mul(fr_ld(w_hi + (j >> W_LO_BITS)), fr_ld(w_lo + (j & (W_LO - 1))));
}

// omega_tab[j] = omega^j for the whole domain (keygen side; 32 B per row)
__global__ void __launch_bounds__(256)
k_p_omega_tab(const fr_t *__restrict__ w_lo, const fr_t *__restrict__ w_hi, uint64_t n,
              fr_t *__restrict__ tab) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        fr_st(tab + j, omega_pow(w_lo, w_hi, (uint32_t)j));
}

// sigma columns [n_cols][n] from the packed mapping (keygen side)
__global__ void __launch_bounds__(256)
k_p_sigma(const uint32_t *__restrict__ pmap, uint64_t cells, const fr_t *__restrict__ w_lo,
          const fr_t *__restrict__ w_hi, const fr_t *__restrict__ dpow, uint64_t *__restrict__ sigma) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        const uint32_t m = pmap[i];
        const fr_t s = fr_mul(fr_ld(dpow + (m >> PMAP_ROW_BITS)),
                              omega_pow(w_lo, w_hi, m & ((1u << PMAP_ROW_BITS) - 1)));
        fr_store_256(sigma + 4 * i, s);
    }
}

// mapping of every cell in closed form (permmap.h), packed position << 24 | row
__global__ void __launch_bounds__(256)
k_p_mapping(Shape sh, uint64_t n_blocks, const PermTemplate *__restrict__ tmpl, int n_cols,
            uint32_t *__restrict__ pmap) {
    const uint64_t n = sh.n(), cells = (uint64_t)n_cols * n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        uint32_t op;
        uint64_t orow;
        perm_successor(sh, n_blocks, *tmpl, (uint32_t)(i / n), i % n, &op, &orow);
        pmap[i] = (op << PMAP_ROW_BITS) | (uint32_t)orow;
    }
}
cudaError_t launch_perm_mapping(h2aes_circuit *c, uint64_t n_blocks, const PermTemplate *tmpl_dev) {
    h2aes_ctx *ctx = c->ctx;
    prof_begin(ctx, "k_p_mapping");
    k_p_mapping<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->sh, n_blocks, tmpl_dev,
                                                            perm_n_columns(c->sh), c->d_pmap);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

struct PArgs {
    Shape sh;
    int n_cols, n_sets;
    const uint32_t *pmap;      // [n_cols][n] packed mapping
    const uint64_t *advice;    // [n_advice][n][4] as produced by h2aes_witness_dev
    const uint4 *frtab;        // byte -> Fr
    const fr_t *chal;          // [1] beta, [2] gamma
    const fr_t *w_lo, *w_hi, *bdp;
    const fr_t *omega_tab;     // [n] omega^j
    uint64_t *Z;               // [n_sets][n][4]: denominators, then factors, then Z
    fr_t *num;                 // [n_sets][n] numerators of the non-constant rows
    fr_t *pre;                 // [n_sets][n] running denominator products inside a chunk
    uint32_t *chunkmask;       // [n_sets][m]
    uint64_t m;
};

// value of permutation column `p` at `row`
__device__ __forceinline__ fr_t perm_value(const PArgs &a, int p, uint64_t row) {
    int is_fixed, idx;
    perm_column(a.sh, p, &is_fixed, &idx);
    if (is_fixed) return fr_from_tab(a.frtab, rc_cell(row));  // round_constants
    return fr_load(a.advice + ((uint64_t)idx * a.sh.n() + row) * 4);
}

// per row of a column set: numerator and denominator over the terms whose cell is in a cycle
__global__ void __launch_bounds__(256) k_p_terms(const PArgs a) {
    const int s = blockIdx.y;
    const uint64_t n = a.sh.n();
    const uint32_t usable = (uint32_t)a.sh.usable();
    const fr_t gamma = fr_ld(a.chal + 2);
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    uint64_t *__restrict__ zout = a.Z + (size_t)s * n * 4;
    fr_t *__restrict__ numout = a.num + (size_t)s * n;
    for (uint64_t c = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < a.m;
         c += warps) {
        const uint64_t row = c * ZI + lane;
        uint32_t moved = 0;  // bit t: term t's cell does not map to itself
        uint32_t mp[PERM_CHUNK];
        if (row < usable) {
#pragma unroll
            for (int t = 0; t < PERM_CHUNK; ++t) {
                const int p = PERM_CHUNK * s + t;
                mp[t] = 0;
                if (p < a.n_cols) {
                    mp[t] = a.pmap[(uint64_t)p * n + row];
                    if (mp[t] != (((uint32_t)p << PMAP_ROW_BITS) | (uint32_t)row)) moved |= 1u << t;
                }
            }
        }
        if (moved) {
            const fr_t wi = fr_ld(a.omega_tab + row);
            fr_t num = fr_one(), den = fr_one();
            bool first = true;
#pragma unroll
            for (int t = 0; t < PERM_CHUNK; ++t) {
                if (!((moved >> t) & 1u)) continue;
                const int p = PERM_CHUNK * s + t;
                const fr_t vg = fr_add(perm_value(a, p, row), gamma);
                const fr_t tn = fr_add(fr_mul(fr_ld(a.bdp + p), wi), vg);
                const fr_t ws = fr_ld(a.omega_tab + (mp[t] & ((1u << PMAP_ROW_BITS) - 1)));
                const fr_t td = fr_add(fr_mul(fr_ld(a.bdp + (mp[t] >> PMAP_ROW_BITS)), ws), vg);
                num = first ? tn : fr_mul(num, tn);
                den = first ? td : fr_mul(den, td);
                first = false;
            }
            fr_store_256(zout + row * 4, den);
            fr_st(numout + row, num);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, moved != 0);
        if (lane == 0) a.chunkmask[(size_t)s * a.m + c] = bal;
    }
}

// factor = numerator / denominator for the non-constant rows: one thread per chunk walks its
// rows (forward: running denominator products; the block shares ONE field inversion through
// forward/reverse product scans of the thread totals; backward: individual inverses).
#ifndef PI_THREADS
#define PI_THREADS 128
#endif
constexpr int PI_T = PI_THREADS;
__global__ void __launch_bounds__(PI_T) k_p_invert(const PArgs a) {
    __shared__ fr_t s_wtot[PI_T / 32];
    __shared__ fr_t s_inv;
    const int s = blockIdx.y;
    const uint64_t n = a.sh.n();
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint64_t e = (uint64_t)blockIdx.x * PI_T + t;
    const bool live = e < a.m;
    const uint32_t mask = live ? a.chunkmask[(size_t)s * a.m + e] : 0u;
    uint64_t *__restrict__ z = a.Z + ((size_t)s * n + (live ? e : 0) * ZI) * 4;
    fr_t *__restrict__ num = a.num + (size_t)s * n + (live ? e : 0) * ZI;
    fr_t *__restrict__ pre = a.pre + (size_t)s * n + (live ? e : 0) * ZI;
    const fr_t one = fr_one();
    fr_t tot = one;
    for (uint32_t mm = mask; mm; mm &= mm - 1) {
        const int q = __ffs(mm) - 1;
        fr_st(pre + q, tot);
        tot = fr_mul(tot, fr_load(z + 4 * q));
    }
    // forward / reverse inclusive product scans of the thread totals
    fr_t finc = tot, rinc = tot;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const fr_t yu = fr_shfl_up(finc, d);
        const fr_t yd = fr_shfl_down(rinc, d);
        if (lane >= d) finc = fr_mul(finc, yu);
        if (lane + d < 32) rinc = fr_mul(rinc, yd);
    }
    if (lane == 31) s_wtot[wid] = finc;
    __syncthreads();
    fr_t before = fr_shfl_up(finc, 1), after = fr_shfl_down(rinc, 1);
    if (lane == 0) before = one;
    if (lane == 31) after = one;
    for (int w = 0; w < PI_T / 32; ++w) {
        const fr_t wt = s_wtot[w];
        if (w < wid) before = fr_mul(before, wt);
        if (w > wid) after = fr_mul(after, wt);
    }
    if (wid == 0) {
        fr_t all = s_wtot[0];
        for (int w = 1; w < PI_T / 32; ++w) all = fr_mul(all, s_wtot[w]);
        const fr_t inv = fr_inv(all);  // denominators are nonzero with overwhelming probability
        if (lane == 0) s_inv = inv;
    }
    __syncthreads();
    if (!mask) return;
    fr_t run = fr_mul(fr_mul(s_inv, before), after);  // 1 / (product of this thread's denominators)
    for (uint32_t mm = mask; mm;) {
        const int q = 31 - __clz(mm);  // backward
        mm &= ~(1u << q);
        const fr_t den = fr_load(z + 4 * q);
        const fr_t inv_den = fr_mul(run, fr_ld(pre + q));
        run = fr_mul(run, den);
        fr_store_256(z + 4 * q, fr_mul(fr_ld(num + q), inv_den));
    }
}

cudaError_t launch_perm_tables(h2aes_circuit *c, const fr_t &omega, const fr_t &delta) {
    h2aes_ctx *ctx = c->ctx;
    const uint32_t n_hi = (uint32_t)(c->sh.n() >> W_LO_BITS);
    const int n_cols = perm_n_columns(c->sh);
    const uint32_t threads = n_hi > W_LO ? n_hi : W_LO;
    prof_begin(ctx, "k_p_tables");
    k_p_tables<<<cdiv(threads, 256), 256, 0, ctx->stream>>>(omega, delta, n_hi, n_cols, c->d_wlo,
                                                            c->d_whi, c->d_dpow);
    prof_end(ctx);
    prof_begin(ctx, "k_p_omega_tab");
    k_p_omega_tab<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->d_wlo, c->d_whi, c->sh.n(),
                                                              c->d_omega_tab);
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_perm_sigma(h2aes_circuit *c, uint64_t *sigma_dev) {
    h2aes_ctx *ctx = c->ctx;
    const uint64_t cells = (uint64_t)perm_n_columns(c->sh) * c->sh.n();
    prof_begin(ctx, "k_p_sigma");
    k_p_sigma<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->d_pmap, cells, c->d_wlo, c->d_whi,
                                                          c->d_dpow, sigma_dev);
    prof_end(ctx);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_perm_product(h2aes_circuit *c, const uint64_t *advice_dev, uint64_t *Zp) {
    h2aes_ctx *ctx = c->ctx;
    cudaStream_t s = ctx->stream;
    const int n_cols = perm_n_columns(c->sh), n_sets = perm_n_sets(c->sh);
    const uint64_t n = c->sh.n();
    PArgs p;
    p.sh = c->sh;
    p.n_cols = n_cols;
    p.n_sets = n_sets;
    p.pmap = c->d_pmap;
    p.advice = advice_dev;
    p.frtab = ctx->d_frtab;
    p.chal = c->d_chal;
    p.w_lo = c->d_wlo; p.w_hi = c->d_whi; p.bdp = c->d_bdp;
    p.omega_tab = c->d_omega_tab;
    p.Z = Zp;
    p.num = c->d_pnum;
    p.pre = c->d_ppre;
    p.chunkmask = c->d_pmask;
    p.m = n / ZI;
    prof_begin(ctx, "k_p_beta");
    k_p_beta<<<cdiv(n_cols, 64), 64, 0, s>>>(c->d_chal, c->d_dpow, n_cols, c->d_bdp);
    prof_end(ctx);
    prof_begin(ctx, "k_p_terms");
    k_p_terms<<<dim3((unsigned)ctx->sm_count * 4, n_sets), 256, 0, s>>>(p);
    prof_end(ctx);
    prof_begin(ctx, "k_p_invert");
    k_p_invert<<<dim3(cdiv(p.m, PI_T), n_sets), PI_T, 0, s>>>(p);
    prof_end(ctx);
    // the n_sets columns are ONE running product (set s starts from the last usable value of set
    // s-1): a single look-back scan over n_sets * n rows (gprod.cu)
    cudaError_t e = launch_z_scan_buffer(c, n_sets, Zp, c->d_pmask);
    if (e) return e;
    ctx->launches += 3;
    return cudaGetLastError();
}

uint32_t lookup_bin_tiles() { return LB_TILES; }
uint64_t run_index_stride(uint64_t n) { return (n >> RUN_SHIFT) + 2; }

}  // namespace h2
