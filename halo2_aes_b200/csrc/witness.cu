// witness.cu — advice / fixed / selector / copy-constraint kernels (sm_100a).
//
// Replaces, for a whole circuit at a time, what the reference does region by region on one CPU
// thread: schedule_key (src/key_schedule.rs:80-224), encrypt + lcon (src/aes128.rs:154-301),
// the chips' assign closures (src/chips/*.rs), load_enc_full_table (src/table.rs:18-192).
//
// Kernels
//   k_block_states   FIPS-197 key expansion with the context's S-box (warp 0 of every CTA, into
//                    shared memory), then half a warp per AES block: the 31 16-byte round states.
//   k_advice         the streaming kernel: one lane per row, computes the row's trace word from
//                    the block's round states (closed form, layout.h), stores the word (4 B) and
//                    the three Fr cells of the column set with one 256-bit store each, from a
//                    256-entry byte->Fr(Montgomery) table in shared memory.  HBM-store bound:
//                    (3*sets+1) * 2^k * 32 bytes per circuit.
//   k_fixed          table columns, round_constants and selector columns (input independent).
//   k_copies         the copy-constraint list (closed form).
#include <cstdlib>

#include "internal.h"

namespace h2 {

struct Key16 {
    uint8_t b[16];
};

// Key expansion (FIPS-197 with the context's S-box) by warp 0 of every CTA into shared memory
// (176 bytes; cheaper than a separate launch), then half a warp per AES block: lane i of the
// half-warp owns state byte i = 4w+j.  CTA 0 also publishes the round keys for k_advice.
__global__ void __launch_bounds__(128)
k_block_states(const uint8_t *__restrict__ sbox, Key16 key, uint8_t *__restrict__ rk_out,
               const uint8_t *__restrict__ pts, uint64_t n_blocks, uint8_t *__restrict__ states,
               uint8_t *__restrict__ cts) {
    __shared__ uint8_t s_sbox[256];
    __shared__ uint8_t s_rk[176];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_sbox[i] = sbox[i];
    if (threadIdx.x < 16) s_rk[threadIdx.x] = key.b[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int l = threadIdx.x;
        for (int r = 1; r <= 10; ++r) {
            const uint8_t *p = s_rk + 16 * (r - 1);
            uint8_t *o = s_rk + 16 * r;
            if (l < 4) {
                const int rot = (13 + l) - (l == 3 ? 4 : 0);
                o[l] = (uint8_t)(p[l] ^ s_sbox[p[rot]] ^ (l == 0 ? rcon(r - 1) : 0));
            }
            __syncwarp();
            for (int i = 1; i < 4; ++i) {
                if (l < 4) o[4 * i + l] = (uint8_t)(p[4 * i + l] ^ o[4 * (i - 1) + l]);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 176; i += blockDim.x) rk_out[i] = s_rk[i];
    if (!n_blocks) return;
    const int lane = threadIdx.x & 31, l = lane & 15, hb = lane & 16;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t blk = 2 * warp + (lane >> 4);
    const bool live = blk < n_blocks;
    const uint64_t bsafe = live ? blk : 0;
    uint8_t *st = states + bsafe * ST_BYTES;
    const int w = l >> 2, c = l & 3;
    uint32_t pt = pts[bsafe * 16 + l];
    uint32_t s = pt ^ s_rk[l];
    if (live) st[16 * ST_PT + l] = (uint8_t)pt;
#pragma unroll 1
    for (int r = 1; r <= 10; ++r) {
        uint32_t sub = s_sbox[s];
        if (live) {
            st[16 * (ST_IN + r - 1) + l] = (uint8_t)s;
            st[16 * (ST_SUB + r - 1) + l] = (uint8_t)sub;
        }
        if (r < 10) {
            uint32_t acc = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint32_t x = __shfl_sync(0xffffffffu, sub, hb + 4 * ((w + t) & 3) + t);
                int m = mix_coeff(c, t);
                acc ^= m == 1 ? x : (m == 2 ? gf_mul2((uint8_t)x) : gf_mul3((uint8_t)x));
            }
            if (live) st[16 * (ST_MIX + r - 1) + l] = (uint8_t)acc;
            s = acc ^ s_rk[16 * r + l];
        } else {
            uint32_t x = __shfl_sync(0xffffffffu, sub, hb + 4 * ((w + c) & 3) + c);
            s = x ^ s_rk[160 + l];
        }
    }
    if (live) {
        st[16 * ST_CT + l] = (uint8_t)s;
        st[16 * 31 + l] = 0;
        if (cts) cts[blk * 16 + l] = (uint8_t)s;
    }
}

// trace word of (set,row): key-schedule rows, AES block rows, or nothing
__device__ __forceinline__ uint32_t advice_word(const Shape &sh, uint64_t n_blocks, int set,
                                                uint64_t row, const uint8_t *s_sbox,
                                                const uint8_t *s_rk,
                                                const uint8_t *__restrict__ states) {
    if (set == 0 && row < KEY_SCHEDULE_ROWS) return ks_trace_word(s_sbox, s_rk, (uint32_t)row);
    if (sh.kind != KIND_AES) return 0;
    uint32_t base = (uint32_t)sh.set_base(set);
    if (row < base) return 0;
    uint32_t rel = (uint32_t)row - base;
    uint32_t b = rel / (uint32_t)AES_ROWS;
    if (b >= sh.blocks_in_set(set, n_blocks)) return 0;
    uint32_t q = rel - b * (uint32_t)AES_ROWS;
    const uint8_t *st = states + (sh.first_block_of_set(set) + b) * ST_BYTES;
    return block_trace_word(st, s_rk, q);
}

// One lane per row.  Work item = (column set | words column, chunk of 32 consecutive rows).
// Output column c of the shard starts at advice + c * out_rows * 4 (u64 units).
__global__ void __launch_bounds__(256)
k_advice(Shape sh, uint64_t n_blocks, const uint8_t *__restrict__ rk,
         const uint8_t *__restrict__ states, const uint8_t *__restrict__ sbox,
         const uint4 *__restrict__ frtab, uint64_t *__restrict__ advice,
         uint32_t *__restrict__ trace, RowMap rm) {
    __shared__ uint4 s_fr[512];
    __shared__ uint8_t s_rk[176];
    __shared__ uint8_t s_sbox[256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_fr[i] = frtab[i];
    for (int i = threadIdx.x; i < 176; i += blockDim.x) s_rk[i] = rk[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_sbox[i] = sbox[i];
    __syncthreads();
    const uint64_t n = sh.n();
    const int nsets = sh.n_sets();
    const uint64_t out_rows = rm.rows;
    const uint64_t chunks = (out_rows + 31) >> 5;
    const uint64_t total = chunks * (uint64_t)(nsets + 1);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = threadIdx.x & 31;
    for (uint64_t item = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         item < total; item += nwarps) {
        const int set = (int)(item / chunks);
        const uint64_t lrow = (item - (uint64_t)set * chunks) * 32 + lane;
        if (lrow >= out_rows) continue;
        const uint64_t row = rm.global(lrow);
        if (set < nsets) {
            const uint32_t word = advice_word(sh, n_blocks, set, row, s_sbox, s_rk, states);
            trace[(uint64_t)set * n + row] = word;
            uint64_t *o = advice + ((uint64_t)(3 * set) * out_rows + lrow) * 4;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint32_t b = (word >> (8 * j)) & 255u;
                store_256_cs(o + (uint64_t)j * out_rows * 4, s_fr[2 * b], s_fr[2 * b + 1]);
            }
        } else {
            const uint32_t b = row < (uint64_t)WORDS_ROWS ? ks_words_byte(s_rk, (uint32_t)row) : 0;
            uint64_t *o = advice + ((uint64_t)(3 * nsets) * out_rows + lrow) * 4;
            store_256_cs(o, s_fr[2 * b], s_fr[2 * b + 1]);
        }
    }
}

// Fixed columns (4 table columns + round_constants) and selector columns as Fr 0/1.
__global__ void __launch_bounds__(256)
k_fixed(Shape sh, uint64_t n_blocks, const uint8_t *__restrict__ sbox,
        const uint4 *__restrict__ frtab, uint64_t *__restrict__ fixed,
        uint64_t *__restrict__ selectors, RowMap rm) {
    __shared__ uint4 s_fr[512];
    __shared__ uint8_t s_sbox[256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_fr[i] = frtab[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_sbox[i] = sbox[i];
    __syncthreads();
    const uint64_t usable = sh.usable();
    const uint64_t out_rows = rm.rows;
    const uint64_t chunks = (out_rows + 31) >> 5;
    const int ncols = sh.n_fixed() + sh.n_selectors();
    const uint64_t total = chunks * (uint64_t)ncols;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = threadIdx.x & 31;
    const int ops = sh.ops_per_set();
    for (uint64_t item = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         item < total; item += nwarps) {
        const int col = (int)(item / chunks);
        const uint64_t lrow = (item - (uint64_t)col * chunks) * 32 + lane;
        if (lrow >= out_rows) continue;
        const uint64_t row = rm.global(lrow);
        uint32_t v;
        uint64_t *o;
        if (col < 5) {
            if (!fixed) continue;
            v = col < 4 ? table_cell(s_sbox, col, row, usable) : rc_cell(row);
            o = fixed + ((uint64_t)col * out_rows + lrow) * 4;
        } else {
            if (!selectors) continue;
            const int q = col - 5;
            if (q < sh.n_complex()) {
                const int set = q / ops, t = q - set * ops;
                v = row_op(sh, n_blocks, set, row) == (uint32_t)(t + 1) ? 1u : 0u;
            } else {
                int ri;
                v = ks_is_rcon_row(row, &ri) ? 1u : 0u;
            }
            o = selectors + ((uint64_t)q * out_rows + lrow) * 4;
        }
        store_256(o, s_fr[2 * v], s_fr[2 * v + 1]);
    }
}

__global__ void __launch_bounds__(256)
k_copies(Shape sh, uint64_t n_copies, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_copies; i += stride) {
        CopyPair cp;
        if (!copy_pair(sh, i, &cp)) continue;
        ulonglong4 v;
        v.x = cp.dst.col;
        v.y = cp.dst.row;
        v.z = cp.src.col;
        v.w = cp.src.row;
        uint4 lo = make_uint4((uint32_t)v.x, 0, (uint32_t)v.y, (uint32_t)(v.y >> 32));
        uint4 hi = make_uint4((uint32_t)v.z, 0, (uint32_t)v.w, (uint32_t)(v.w >> 32));
        store_256(out + 4 * i, lo, hi);
    }
}

// byte -> Fr(Montgomery) table, built on the device once per context
__global__ void k_build_frtab(uint4 *__restrict__ frtab) {
    const int b = threadIdx.x;
    fr_t f = fr_from_u32((uint32_t)b);
    frtab[2 * b] = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
    frtab[2 * b + 1] = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
}

// n one-row chip regions (the chips' stand-alone assign methods): row i holds x[i] in the chip's
// first column, then y[i] (xor) and the chip's output byte; the output is also returned as a byte
__global__ void __launch_bounds__(256)
k_chip_rows(int tag, const uint8_t *__restrict__ sbox, const uint4 *__restrict__ frtab,
            const uint8_t *__restrict__ x, const uint8_t *__restrict__ y, uint64_t n,
            uint64_t *__restrict__ cells, uint8_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t a = x[i], b = tag == 2 ? y[i] : (uint8_t)0;
        const uint8_t o = tag == 1 ? a
                        : tag == 2 ? (uint8_t)(a ^ b)
                        : tag == 3 ? sbox[a]
                        : tag == 4 ? gf_mul2(a) : gf_mul3(a);
        if (out) out[i] = o;
        if (!cells) continue;
        // columns of the region: range (x), xor (x, y, z), sbox / mul (x, y = out); the rest zero
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int col = 0; col < 3; ++col) {
            const bool used = col == 0 || (tag == 2) || (tag >= 3 && col == 1);
            const uint8_t v = col == 0 ? a : (tag == 2 ? (col == 1 ? b : o) : o);
            uint4 *dst = reinterpret_cast<uint4 *>(cells + ((size_t)col * n + i) * 4);
            dst[0] = used ? __ldg(frtab + 2 * v) : z4;
            dst[1] = used ? __ldg(frtab + 2 * v + 1) : z4;
        }
    }
}

// ---- launchers -----------------------------------------------------------------------------
// H2AES_STREAM_PAD_KB (tuning): unused dynamic shared memory requested by the HBM-bound streaming
// kernels, which caps how many of their CTAs an SM holds and leaves registers / warp slots for
// the multiply-bound kernels of the next circuit in flight.
size_t stream_pad_bytes(const void *func) {
    static int kb = -1;
    if (kb < 0) {
        const char *s = getenv("H2AES_STREAM_PAD_KB");
        kb = s ? atoi(s) : 0;
    }
    if (kb > 48) cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
    return (size_t)kb * 1024;
}

static inline int grid_for(const h2aes_ctx *ctx, int ctas_per_sm) { return ctx->sm_count * ctas_per_sm; }
// resident CTAs per SM of the streaming kernels (H2AES_STREAM_CTAS_PER_SM, 0 = the default 8)
static inline int stream_ctas(int dflt) {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("H2AES_STREAM_CTAS_PER_SM");
        v = s ? atoi(s) : 0;
    }
    return v ? v : dflt;
}

cudaError_t launch_build_frtab(h2aes_ctx *ctx) {
    prof_begin(ctx, "k_build_frtab");
    k_build_frtab<<<1, 256, 0, ctx->stream>>>(ctx->d_frtab);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_chip_rows(h2aes_ctx *ctx, int tag, const uint8_t *x, const uint8_t *y, uint64_t n,
                             uint64_t *cells, uint8_t *out) {
    prof_begin(ctx, "k_chip_rows");
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    k_chip_rows<<<grid, 256, 0, ctx->stream>>>(tag, ctx->d_sbox, ctx->d_frtab, x, y, n, cells, out);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_witness(h2aes_circuit *c, const uint8_t *key, const uint8_t *pts_dev,
                           uint64_t n_blocks, uint64_t *advice_dev, uint8_t *cts_dev, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    Key16 k16;
    for (int i = 0; i < 16; ++i) k16.b[i] = key[i];
    const uint64_t blocks_per_cta = 8;
    unsigned grid = n_blocks ? (unsigned)((n_blocks + blocks_per_cta - 1) / blocks_per_cta) : 1u;
    prof_begin(ctx, "k_block_states");
    k_block_states<<<grid, 128, 0, ctx->stream>>>(ctx->d_sbox, k16, c->d_rk, pts_dev, n_blocks,
                                                  c->d_states, cts_dev);
    prof_end(ctx);
    ctx->launches++;
    const size_t pad = stream_pad_bytes((const void *)k_advice);
    prof_begin(ctx, "k_advice");
    k_advice<<<grid_for(ctx, stream_ctas(8)), 256, pad, ctx->stream>>>(c->sh, n_blocks, c->d_rk, c->d_states,
                                                        ctx->d_sbox, ctx->d_frtab, advice_dev,
                                                        c->d_trace, rm);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
}

// `n` more blocks after the `first_block` already held: their round states, then the advice
// columns again from the states of ALL blocks (the streaming kernel writes whole columns: one
// pass, 0.11 ms at K=20, N=5 — cheaper than a second kernel that patches row ranges)
cudaError_t launch_witness_append(h2aes_circuit *c, const uint8_t *key, const uint8_t *pts_dev,
                                  uint64_t first_block, uint64_t n, uint64_t *advice_dev,
                                  uint8_t *cts_dev) {
    h2aes_ctx *ctx = c->ctx;
    Key16 k16;
    for (int i = 0; i < 16; ++i) k16.b[i] = key[i];
    const unsigned grid = (unsigned)((n + 7) / 8);
    prof_begin(ctx, "k_block_states");
    k_block_states<<<grid, 128, 0, ctx->stream>>>(ctx->d_sbox, k16, c->d_rk, pts_dev, n,
                                                  c->d_states + first_block * ST_BYTES, cts_dev);
    prof_end(ctx);
    const size_t pad = stream_pad_bytes((const void *)k_advice);
    prof_begin(ctx, "k_advice");
    k_advice<<<grid_for(ctx, stream_ctas(8)), 256, pad, ctx->stream>>>(c->sh, first_block + n, c->d_rk, c->d_states,
                                                          ctx->d_sbox, ctx->d_frtab, advice_dev,
                                                          c->d_trace, row_map_of(c->sh.n(), 0, 1, 0));
    prof_end(ctx);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_fixed(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed_dev,
                         uint64_t *selectors_dev, const RowMap &rm) {
    h2aes_ctx *ctx = c->ctx;
    prof_begin(ctx, "k_fixed");
    k_fixed<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(c->sh, n_blocks, ctx->d_sbox, ctx->d_frtab,
                                                       fixed_dev, selectors_dev, rm);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_copies(h2aes_circuit *c, uint64_t n_copies, uint64_t *out) {
    h2aes_ctx *ctx = c->ctx;
    prof_begin(ctx, "k_copies");
    k_copies<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(c->sh, n_copies, out);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
}

}  // namespace h2
