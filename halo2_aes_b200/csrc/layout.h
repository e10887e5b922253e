// layout.h — closed-form description of every cell the halo2-aes chips assign.
//
// Shared by the CUDA kernels (__device__) and by the host-side C ABI (__host__), so the same
// source is exercised by the CPU-only tests (through the h2aes_dbg_* hooks) and by the GPU.
//
// Nothing here replays regions: each function answers "what is in (column,row)?" directly.
// The derivation (reference file:line, and [UPSTREAM] = halo2_proofs v0.3.0 SimpleFloorPlanner:
// a region starts at the max next-free row over the columns/selectors it touches):
//
//   key schedule (src/key_schedule.rs:80-224), column set 0 = (a0,a1,a2), w = words_column:
//     w[0..16)                       key bytes                                     (:98-118)
//     round r=1..10, prev = round key r-1, b = 40(r-1):
//       w[16+8(r-1)+j]              = prev[{13,14,15,12}[j]]          "shift"      (:138-154)
//       w[20+8(r-1)]                = RCON[r-1], next 3 rows 0; q_eq_rcon + fixed rc (:162-188)
//       a0/a1[b+j]          j<4      sbox(rot[j])                                  (:156-159)
//       a0/a1/a2[b+4+j]              xor(subbed[j], rc_j)                          (:190-194)
//       a0/a1/a2[b+8+j]              xor(prev[j], rconned[j])                      (:197-202)
//       a0/a1/a2[b+12+4(i-1)+j] i=1..3  xor(prev[4i+j], new[4(i-1)+j])             (:207-216)
//       a0[b+24+m]          m<16     range_check(new[m])                           (:218-221)
//   one AES block (src/aes128.rs:154-301), rows relative to the block start, 1360 rows:
//     [0,16)     plaintext in a0                                                   (:176-192)
//     [16,32)    xor(pt[j], rk0[j])                                                (:194-198)
//     round r=1..9 at R = 32+144(r-1):
//       [R,R+16)              sbox(state[j])                                       (:203-209)
//       R+16+7(4w+c)+t, t<4   term t of lcon(word w, matrix row c): copy / mul2 / mul3
//       ...+4, +5, +6         xor(t0,t1), xor(t2,t3), xor(i1,i2)                   (:268-301)
//       [R+128,R+144)         xor(mixed[i], rk_r[i])                               (:250-261)
//     [1328,1344) sbox (round 10), [1344,1360) xor(shifted[i], rk10[i]) = ciphertext
//   block -> (column set, first row): aes_callable (src/aes128.rs:303-325).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define H2_HD __host__ __device__ __forceinline__
#else
#define H2_HD inline
#endif

namespace h2 {

enum : int { KIND_AES = 0, KIND_KEYSCHED = 1 };
enum : uint32_t { OP_NONE = 0, OP_RANGE = 1, OP_XOR = 2, OP_SBOX = 3, OP_MUL2 = 4, OP_MUL3 = 5 };

constexpr uint64_t KEY_SCHEDULE_ROWS_ACCOUNTED = 1760;  // src/constant.rs:113 (capacity rule only)
constexpr uint64_t KEY_SCHEDULE_ROWS = 400;             // rows of advices[0][0] really used
constexpr uint64_t AES_ROWS = 1360;                     // src/constant.rs:114
constexpr uint64_t TABLE_ROWS = 66561;                  // src/table.rs: 256+256+65536+256+256+1
constexpr uint64_t UNUSABLE_ROWS = 6;                   // blinding_factors()+1  [UPSTREAM]
constexpr uint64_t COPIES_KEY_SCHEDULE = 640;
constexpr uint64_t COPIES_PER_BLOCK = 1952;
constexpr int WORDS_ROWS = 96;                          // 16 + 8*10 rows of words_column
constexpr int MAX_SETS = 64;

// table row ranges (src/table.rs, in assignment order — NOT Tag order)
constexpr uint32_t TROW_U8 = 0, TROW_SBOX = 256, TROW_XOR = 512, TROW_MUL2 = 66048,
                   TROW_MUL3 = 66304, TROW_ZERO = 66560;

struct Shape {
    int kind, k, N;
    H2_HD uint64_t n() const { return (uint64_t)1 << k; }
    H2_HD uint64_t usable() const { return n() - UNUSABLE_ROWS; }
    H2_HD int n_sets() const { return kind == KIND_AES ? N : 1; }
    H2_HD int n_advice() const { return 3 * n_sets() + 1; }
    H2_HD int n_fixed() const { return 5; }
    H2_HD int n_complex() const { return kind == KIND_AES ? 5 * N : 3; }
    H2_HD int n_selectors() const { return n_complex() + 1; }
    H2_HD int n_lookups() const { return n_complex(); }
    H2_HD int ops_per_set() const { return kind == KIND_AES ? 5 : 3; }
    H2_HD int words_column() const { return 3 * n_sets(); }
    H2_HD uint64_t cap0() const { return (n() - KEY_SCHEDULE_ROWS_ACCOUNTED) / AES_ROWS; }
    H2_HD uint64_t cap() const { return n() / AES_ROWS; }
    H2_HD uint64_t capacity() const {
        return kind == KIND_AES ? cap0() + (uint64_t)(N - 1) * cap() : 0;
    }
    // first row of set `s` that AES blocks use
    H2_HD uint64_t set_base(int s) const { return s == 0 ? KEY_SCHEDULE_ROWS : 0; }
    // number of AES blocks living in set `s` when n_blocks were encrypted
    H2_HD uint64_t blocks_in_set(int s, uint64_t n_blocks) const {
        uint64_t c0 = cap0(), c = cap();
        if (s == 0) return n_blocks < c0 ? n_blocks : c0;
        uint64_t before = c0 + (uint64_t)(s - 1) * c;
        if (n_blocks <= before) return 0;
        uint64_t r = n_blocks - before;
        return r < c ? r : c;
    }
    // global index of the first block of set `s`
    H2_HD uint64_t first_block_of_set(int s) const {
        return s == 0 ? 0 : cap0() + (uint64_t)(s - 1) * cap();
    }
};

// Which rows of every column a (shard of a) circuit holds.  The rows are cut into ownership tiles
// of `ts` consecutive rows dealt round-robin to the G shards; shard g keeps its tiles back to back
// ("local rows").  ts = n / G is the contiguous split (one tile per shard), G = 1 the whole column.
struct RowMap {
    uint64_t n, ts, rows;  // rows of a column, rows per ownership tile, local rows of this shard
    uint32_t G, g;
    H2_HD uint64_t global(uint64_t lr) const {
        if (G == 1) return lr;
        const uint32_t t = (uint32_t)lr / (uint32_t)ts;  // (lr < 2^32: k <= 26)
        return ((uint64_t)t * G + g) * ts + (lr - (uint64_t)t * ts);
    }
    // segments = ownership tiles per shard, the same on every shard (a shard holding one tile less
    // pads with an empty one): what the exchange of the partial products is sized by
    H2_HD uint32_t segments() const {
        if (G == 1) return 1;
        const uint64_t nt = (n + ts - 1) / ts;
        return (uint32_t)((nt + G - 1) / G);
    }
};
H2_HD RowMap row_map_of(uint64_t n, uint64_t ts, uint32_t G, uint32_t g) {
    RowMap m;
    m.n = n;
    m.ts = ts ? ts : n / G;
    m.G = G;
    m.g = g;
    const uint64_t nt = (n + m.ts - 1) / m.ts;       // ownership tiles of a column
    const uint64_t mine = nt > g ? (nt - g + G - 1) / G : 0;  // ... held by shard g
    uint64_t rows = mine * m.ts;
    if (mine && (nt - 1) % G == g) rows -= nt * m.ts - n;  // the last tile of the column is partial
    m.rows = rows;
    return m;
}

// aes_callable + get_advices: block -> (set, first row).  false past capacity.
H2_HD bool block_position(const Shape &sh, uint64_t block, int *set, uint64_t *row) {
    uint64_t c0 = sh.cap0(), c = sh.cap();
    if (block < c0) {
        *set = 0;
        *row = KEY_SCHEDULE_ROWS + AES_ROWS * block;
        return true;
    }
    uint64_t r = block - c0;
    if (c == 0) return false;
    uint64_t s = 1 + r / c;
    if (s >= (uint64_t)sh.N) return false;
    *set = (int)s;
    *row = (r % c) * AES_ROWS;
    return true;
}

// MixColumns matrix of aes128.rs:228-233: M[c][t] = 2 if t==c, 3 if t==c+1 (mod 4), else 1.
H2_HD int mix_coeff(int c, int t) { return t == c ? 2 : (t == ((c + 1) & 3) ? 3 : 1); }

H2_HD uint8_t xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }
// == src/constant.rs MUL_BY_2 / MUL_BY_3 (checked against the reference data in tests/golden)
H2_HD uint8_t gf_mul2(uint8_t x) { return xtime(x); }
H2_HD uint8_t gf_mul3(uint8_t x) { return (uint8_t)(xtime(x) ^ x); }

// src/utils.rs:28
H2_HD uint8_t rcon(int i) {
    const uint8_t t[10] = {1, 2, 4, 8, 16, 32, 64, 128, 27, 54};
    return t[i];
}

H2_HD uint32_t pack_word(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t op) {
    return a0 | (a1 << 8) | (a2 << 16) | (op << 24);
}

// ---------------------------------------------------------------------------------------------
// key schedule
// ---------------------------------------------------------------------------------------------
// FIPS-197 key expansion with the caller's S-box: rk[16r+i] = byte i of round key r.
H2_HD void expand_key(const uint8_t *sbox, const uint8_t key[16], uint8_t rk[176]) {
    for (int i = 0; i < 16; ++i) rk[i] = key[i];
    for (int r = 1; r <= 10; ++r) {
        const uint8_t *p = rk + 16 * (r - 1);
        uint8_t *o = rk + 16 * r;
        const int rot[4] = {13, 14, 15, 12};
        for (int j = 0; j < 4; ++j)
            o[j] = (uint8_t)(p[j] ^ sbox[p[rot[j]]] ^ (j == 0 ? rcon(r - 1) : 0));
        for (int i = 1; i < 4; ++i)
            for (int j = 0; j < 4; ++j) o[4 * i + j] = (uint8_t)(p[4 * i + j] ^ o[4 * (i - 1) + j]);
    }
}

// trace word of row `row` (< 400) of column set 0 produced by schedule_key
H2_HD uint32_t ks_trace_word(const uint8_t *sbox, const uint8_t *rk, uint32_t row) {
    uint32_t r = row / 40 + 1, o = row % 40;
    const uint8_t *p = rk + 16 * (r - 1), *nw = rk + 16 * r;
    if (o < 4) {
        uint32_t x = p[(13 + o) - (o == 3 ? 4 : 0)];  // {13,14,15,12}[o]
        return pack_word(x, sbox[x], 0, OP_SBOX);
    }
    if (o < 8) {
        uint32_t j = o - 4;
        uint32_t x = sbox[p[(13 + j) - (j == 3 ? 4 : 0)]];
        uint32_t y = j == 0 ? rcon(r - 1) : 0;
        return pack_word(x, y, x ^ y, OP_XOR);
    }
    if (o < 12) {
        uint32_t j = o - 8;
        uint32_t x = p[j], z = nw[j];
        return pack_word(x, x ^ z, z, OP_XOR);  // y = rconned[j] = prev[j] ^ new[j]
    }
    if (o < 24) {
        uint32_t m = o - 12 + 4;  // index 4i+j of the new word byte, i=1..3
        return pack_word(p[m], nw[m - 4], nw[m], OP_XOR);
    }
    return pack_word(nw[o - 24], 0, 0, OP_RANGE);
}

// byte in row `row` (< 96) of words_column
H2_HD uint32_t ks_words_byte(const uint8_t *rk, uint32_t row) {
    if (row < 16) return rk[row];
    uint32_t r = (row - 16) / 8 + 1, o = (row - 16) % 8;
    if (o < 4) return rk[16 * (r - 1) + (13 + o) - (o == 3 ? 4 : 0)];
    return o == 4 ? rcon(r - 1) : 0;
}
// rows of words_column / round_constants where q_eq_rcon is enabled: 20+8(r-1)
H2_HD bool ks_is_rcon_row(uint64_t row, int *round_index) {
    if (row < 20 || row >= 100 || ((row - 20) & 7)) return false;
    *round_index = (int)((row - 20) >> 3);
    return true;
}

// ---------------------------------------------------------------------------------------------
// one AES block: round states, then any row's trace word from the states
// ---------------------------------------------------------------------------------------------
// 32 slots of 16 bytes per block:
//   slot r-1      (r=1..10)  input of SubBytes in round r (slot 0 = pt ^ rk0)
//   slot 10+r-1   (r=1..10)  S[state]  (before ShiftRows)
//   slot 20+r-1   (r=1..9)   MixColumns output (after ShiftRows)
//   slot 29 plaintext, slot 30 ciphertext, slot 31 unused
constexpr int ST_IN = 0, ST_SUB = 10, ST_MIX = 20, ST_PT = 29, ST_CT = 30, ST_SLOTS = 32;
constexpr int ST_BYTES = ST_SLOTS * 16;

H2_HD void aes_block_states(const uint8_t *sbox, const uint8_t *rk, const uint8_t pt[16],
                            uint8_t *st /* ST_BYTES */) {
    uint8_t s[16];
    for (int i = 0; i < 16; ++i) {
        st[16 * ST_PT + i] = pt[i];
        s[i] = (uint8_t)(pt[i] ^ rk[i]);
    }
    for (int r = 1; r <= 10; ++r) {
        uint8_t sub[16], sh[16];
        for (int i = 0; i < 16; ++i) {
            st[16 * (ST_IN + r - 1) + i] = s[i];
            sub[i] = sbox[s[i]];
            st[16 * (ST_SUB + r - 1) + i] = sub[i];
        }
        // aes128.rs:216-223: shifted[w][t] = subbed[(w+t)%4][t]
        for (int w = 0; w < 4; ++w)
            for (int t = 0; t < 4; ++t) sh[4 * w + t] = sub[4 * ((w + t) & 3) + t];
        if (r < 10) {
            for (int w = 0; w < 4; ++w)
                for (int c = 0; c < 4; ++c) {
                    uint8_t acc = 0;
                    for (int t = 0; t < 4; ++t) {
                        uint8_t x = sh[4 * w + t];
                        int m = mix_coeff(c, t);
                        acc ^= m == 1 ? x : (m == 2 ? gf_mul2(x) : gf_mul3(x));
                    }
                    st[16 * (ST_MIX + r - 1) + 4 * w + c] = acc;
                    s[4 * w + c] = (uint8_t)(acc ^ rk[16 * r + 4 * w + c]);
                }
        } else {
            for (int i = 0; i < 16; ++i) s[i] = (uint8_t)(sh[i] ^ rk[160 + i]);
        }
    }
    for (int i = 0; i < 16; ++i) {
        st[16 * ST_CT + i] = s[i];
        st[16 * 31 + i] = 0;
    }
}

// trace word of block-relative row q (< 1360)
H2_HD uint32_t block_trace_word(const uint8_t *st, const uint8_t *rk, uint32_t q) {
    if (q < 16) return pack_word(st[16 * ST_PT + q], 0, 0, OP_NONE);
    if (q < 32) {
        uint32_t j = q - 16;
        return pack_word(st[16 * ST_PT + j], rk[j], st[16 * ST_IN + j], OP_XOR);
    }
    if (q >= 1328) {
        if (q < 1344) {
            uint32_t j = q - 1328;
            return pack_word(st[16 * (ST_IN + 9) + j], st[16 * (ST_SUB + 9) + j], 0, OP_SBOX);
        }
        uint32_t i = q - 1344, w = i >> 2, j = i & 3;
        uint32_t x = st[16 * (ST_SUB + 9) + 4 * ((w + j) & 3) + j];
        return pack_word(x, rk[160 + i], st[16 * ST_CT + i], OP_XOR);
    }
    uint32_t r = (q - 32) / 144 + 1, rr = (q - 32) % 144;  // round 1..9
    const uint8_t *sub = st + 16 * (ST_SUB + r - 1);
    if (rr < 16) return pack_word(st[16 * (ST_IN + r - 1) + rr], sub[rr], 0, OP_SBOX);
    if (rr >= 128) {
        uint32_t i = rr - 128;
        uint32_t x = st[16 * (ST_MIX + r - 1) + i];
        uint32_t y = rk[16 * r + i];
        return pack_word(x, y, x ^ y, OP_XOR);  // == next round's input (or ST_IN + r)
    }
    uint32_t l = (rr - 16) / 7, t7 = (rr - 16) % 7;  // lcon index 4w+c, row within lcon
    uint32_t w = l >> 2, c = l & 3;
    uint32_t tmp[4];
    uint32_t xs[4];
    for (int t = 0; t < 4; ++t) {
        uint32_t x = sub[4 * ((w + t) & 3) + t];
        int m = mix_coeff((int)c, t);
        xs[t] = x;
        tmp[t] = m == 1 ? x : (m == 2 ? gf_mul2((uint8_t)x) : gf_mul3((uint8_t)x));
    }
    if (t7 < 4) {
        int m = mix_coeff((int)c, (int)t7);
        if (m == 1) return pack_word(xs[t7], 0, 0, OP_NONE);
        return pack_word(xs[t7], tmp[t7], 0, m == 2 ? OP_MUL2 : OP_MUL3);
    }
    uint32_t i1 = tmp[0] ^ tmp[1], i2 = tmp[2] ^ tmp[3];
    if (t7 == 4) return pack_word(tmp[0], tmp[1], i1, OP_XOR);
    if (t7 == 5) return pack_word(tmp[2], tmp[3], i2, OP_XOR);
    return pack_word(i1, i2, i1 ^ i2, OP_XOR);
}

// op code only (input independent) of block-relative row q
H2_HD uint32_t block_row_op(uint32_t q) {
    if (q < 16) return OP_NONE;
    if (q < 32) return OP_XOR;
    if (q >= 1328) return q < 1344 ? OP_SBOX : OP_XOR;
    uint32_t rr = (q - 32) % 144;
    if (rr < 16) return OP_SBOX;
    if (rr >= 128) return OP_XOR;
    uint32_t l = (rr - 16) / 7, t7 = (rr - 16) % 7;
    if (t7 >= 4) return OP_XOR;
    int m = mix_coeff((int)(l & 3), (int)t7);
    return m == 1 ? OP_NONE : (m == 2 ? OP_MUL2 : OP_MUL3);
}
H2_HD uint32_t ks_row_op(uint32_t row) {
    uint32_t o = row % 40;
    return o < 4 ? OP_SBOX : (o < 24 ? OP_XOR : OP_RANGE);
}

// op code of (set,row) for a circuit with n_blocks encrypted blocks (selector pattern)
H2_HD uint32_t row_op(const Shape &sh, uint64_t n_blocks, int set, uint64_t row) {
    if (set == 0 && row < KEY_SCHEDULE_ROWS) return ks_row_op((uint32_t)row);
    if (sh.kind != KIND_AES) return OP_NONE;
    uint64_t base = sh.set_base(set);
    if (row < base) return OP_NONE;
    uint64_t rel = row - base;
    uint64_t b = rel / AES_ROWS;
    if (b >= sh.blocks_in_set(set, n_blocks)) return OP_NONE;
    return block_row_op((uint32_t)(rel - b * AES_ROWS));
}

// rows on which the selector of (set, op) is enabled after schedule_key + n_blocks encrypts:
// per key schedule 160 range / 200 xor / 40 sbox rows, per block 608 xor / 160 sbox / 144 mul2 /
// 144 mul3 rows
H2_HD uint64_t active_rows(const Shape &sh, uint64_t n_blocks, int set, uint32_t op) {
    const uint64_t per_block[6] = {0, 0, 608, 160, 144, 144};
    const uint64_t per_ks[6] = {0, 160, 200, 40, 0, 0};
    uint64_t r = set == 0 ? per_ks[op] : 0;
    if (sh.kind == KIND_AES) r += sh.blocks_in_set(set, n_blocks) * per_block[op];
    return r;
}

// ---------------------------------------------------------------------------------------------
// fixed columns: 4 table columns (src/table.rs:18-192 + [UPSTREAM] padding with the offset-0
// value over the usable rows), then round_constants.  Value is a small integer (< 256).
// ---------------------------------------------------------------------------------------------
H2_HD uint32_t table_cell(const uint8_t *sbox, int col, uint64_t row, uint64_t usable) {
    if (row >= usable) return 0;
    if (row >= TABLE_ROWS) return col == 0 ? 1 : 0;  // default = row 0 = (Tag::U8, 0, 0, 0)
    uint32_t r = (uint32_t)row;
    uint32_t tag, x, y = 0, z = 0;
    if (r < TROW_SBOX) { tag = 1; x = r; }
    else if (r < TROW_XOR) { tag = 3; x = r - TROW_SBOX; y = sbox[x]; }
    else if (r < TROW_MUL2) { tag = 2; x = (r - TROW_XOR) >> 8; y = (r - TROW_XOR) & 255; z = x ^ y; }
    else if (r < TROW_MUL3) { tag = 4; x = r - TROW_MUL2; y = gf_mul2((uint8_t)x); }
    else if (r < TROW_ZERO) { tag = 5; x = r - TROW_MUL3; y = gf_mul3((uint8_t)x); }
    else { tag = 0; x = 0; }
    return col == 0 ? tag : (col == 1 ? x : (col == 2 ? y : z));
}
H2_HD uint32_t rc_cell(uint64_t row) {
    int ri;
    return ks_is_rcon_row(row, &ri) ? rcon(ri) : 0;
}

// table row index hit by an active row with trace word `w` (op != NONE)
H2_HD uint32_t table_row_of_word(uint32_t w) {
    uint32_t op = w >> 24, x = w & 255, y = (w >> 8) & 255;
    switch (op) {
    case OP_RANGE: return TROW_U8 + x;
    case OP_XOR: return TROW_XOR + (x << 8) + y;
    case OP_SBOX: return TROW_SBOX + x;
    case OP_MUL2: return TROW_MUL2 + x;
    case OP_MUL3: return TROW_MUL3 + x;
    default: return TROW_ZERO;
    }
}

// ---------------------------------------------------------------------------------------------
// copy constraints (every copy_advice in the chips), closed form.
// A cell is (advice column, absolute row).  Columns: set s -> 3s, 3s+1, 3s+2; words = 3*sets.
// ---------------------------------------------------------------------------------------------
struct CellRef {
    uint32_t col;
    uint64_t row;
};
struct CopyPair {
    CellRef dst, src;  // dst = the freshly assigned copy, src = the cell it was copied from
};

// cell holding byte i of round key r (what schedule_keys returns, key_schedule.rs:87-95)
H2_HD CellRef round_key_cell(const Shape &sh, int r, int i) {
    if (r == 0) return CellRef{(uint32_t)sh.words_column(), (uint64_t)i};
    return CellRef{2u, (uint64_t)(40 * (r - 1) + 8 + i)};  // a2 of set 0
}

// copy `idx` (0..639) made by schedule_key
H2_HD CopyPair ks_copy(const Shape &sh, uint32_t idx) {
    int r = (int)(idx / 64) + 1;
    uint32_t o = idx % 64;
    uint64_t b = 40 * (uint64_t)(r - 1);
    uint32_t wcol = (uint32_t)sh.words_column();
    uint64_t wrow = 16 + 8 * (uint64_t)(r - 1);
    CopyPair cp;
    if (o < 4) {  // shift: words[wrow+j] <- prev[{13,14,15,12}[j]]
        int j = (int)o;
        cp.dst = CellRef{wcol, wrow + j};
        cp.src = round_key_cell(sh, r - 1, (13 + j) - (j == 3 ? 4 : 0));
    } else if (o < 8) {  // sbox: a0[b+j] <- words[wrow+j]
        uint32_t j = o - 4;
        cp.dst = CellRef{0, b + j};
        cp.src = CellRef{wcol, wrow + j};
    } else if (o < 16) {  // xor(subbed, rc): a0 <- a1[b+j] ; a1 <- words[wrow+4+j]
        uint32_t j = (o - 8) >> 1, side = (o - 8) & 1;
        cp.dst = CellRef{side, b + 4 + j};
        cp.src = side == 0 ? CellRef{1, b + j} : CellRef{wcol, wrow + 4 + j};
    } else if (o < 24) {  // xor(prev[j], rconned[j])
        uint32_t j = (o - 16) >> 1, side = (o - 16) & 1;
        cp.dst = CellRef{side, b + 8 + j};
        cp.src = side == 0 ? round_key_cell(sh, r - 1, (int)j) : CellRef{2, b + 4 + j};
    } else if (o < 48) {  // xor(prev[4i+j], new[4(i-1)+j]), m = 4i+j in 4..15
        uint32_t m = 4 + ((o - 24) >> 1), side = (o - 24) & 1;
        cp.dst = CellRef{side, b + 8 + m};
        cp.src = side == 0 ? round_key_cell(sh, r - 1, (int)m) : CellRef{2, b + 8 + m - 4};
    } else {  // range: a0[b+24+m] <- new[m]
        uint32_t m = o - 48;
        cp.dst = CellRef{0, b + 24 + m};
        cp.src = CellRef{2, b + 8 + m};
    }
    return cp;
}

// copy `idx` (0..1951) made by one `encrypt` whose block starts at (set, base)
H2_HD CopyPair block_copy(const Shape &sh, int set, uint64_t base, uint32_t idx) {
    uint32_t a0 = 3 * (uint32_t)set, a1 = a0 + 1, a2 = a0 + 2;
    CopyPair cp;
    if (idx < 32) {  // xor(pt[j], rk0[j])
        uint32_t j = idx >> 1, side = idx & 1;
        cp.dst = CellRef{side ? a1 : a0, base + 16 + j};
        cp.src = side ? round_key_cell(sh, 0, (int)j) : CellRef{a0, base + j};
        return cp;
    }
    idx -= 32;
    if (idx < 9 * 208) {
        uint32_t r = idx / 208 + 1, o = idx % 208;
        uint64_t R = base + 32 + 144 * (uint64_t)(r - 1);
        if (o < 16) {  // sbox(prev_round[j]) ; prev_round = previous xor outputs
            cp.dst = CellRef{a0, R + o};
            cp.src = CellRef{a2, r == 1 ? base + 16 + o : R - 16 + o};
            return cp;
        }
        o -= 16;
        if (o < 160) {  // 16 lcons x 10 copies
            uint32_t l = o / 10, e = o % 10, w = l >> 2, c = l & 3;
            uint64_t L = R + 16 + 7 * (uint64_t)l;
            if (e < 4) {  // term t: a0[L+t] <- subbed[(w+t)%4][t]
                cp.dst = CellRef{a0, L + e};
                cp.src = CellRef{a1, R + 4 * ((w + e) & 3) + e};
            } else if (e < 8) {  // xor(tmp0,tmp1) at L+4, xor(tmp2,tmp3) at L+5
                uint32_t t = e - 4, side = t & 1;
                cp.dst = CellRef{side ? a1 : a0, L + 4 + (t >> 1)};
                // tmp[t] is the copied cell itself for coefficient 1, else the chip output in a1
                cp.src = CellRef{mix_coeff((int)c, (int)t) == 1 ? a0 : a1, L + t};
            } else {  // xor(inter_1, inter_2) at L+6
                uint32_t side = e - 8;
                cp.dst = CellRef{side ? a1 : a0, L + 6};
                cp.src = CellRef{a2, L + 4 + side};
            }
            return cp;
        }
        o -= 160;  // xor(mixed[i], rk_r[i])
        uint32_t i = o >> 1, side = o & 1;
        cp.dst = CellRef{side ? a1 : a0, R + 128 + i};
        cp.src = side ? round_key_cell(sh, (int)r, (int)i) : CellRef{a2, R + 16 + 7 * (uint64_t)i + 6};
        return cp;
    }
    idx -= 9 * 208;  // round 10
    uint64_t R = base + 1328;
    if (idx < 16) {
        cp.dst = CellRef{a0, R + idx};
        cp.src = CellRef{a2, R - 16 + idx};
        return cp;
    }
    idx -= 16;
    uint32_t i = idx >> 1, side = idx & 1, w = i >> 2, j = i & 3;
    cp.dst = CellRef{side ? a1 : a0, R + 16 + i};
    cp.src = side ? round_key_cell(sh, 10, (int)i) : CellRef{a1, R + 4 * ((w + j) & 3) + j};
    return cp;
}

// global copy index -> pair.  Order: the 640 key-schedule copies, then block 0, 1, ...
H2_HD bool copy_pair(const Shape &sh, uint64_t idx, CopyPair *out) {
    if (idx < COPIES_KEY_SCHEDULE) {
        *out = ks_copy(sh, (uint32_t)idx);
        return true;
    }
    idx -= COPIES_KEY_SCHEDULE;
    uint64_t b = idx / COPIES_PER_BLOCK;
    int set;
    uint64_t base;
    if (!block_position(sh, b, &set, &base)) return false;
    *out = block_copy(sh, set, base, (uint32_t)(idx - b * COPIES_PER_BLOCK));
    return true;
}


// ---------------------------------------------------------------------------------------------
// permutation argument: columns in the order of the enable_equality calls —
// key_schedule.rs:52-57 (advices[0][0..3], words_column, enable_constant(round_constants)), then
// aes128.rs:126-130 (every advice column)  [UPSTREAM halo2 permutation::Argument::add_column]
// ---------------------------------------------------------------------------------------------
constexpr int PERM_CHUNK = 3;  // cs.degree() - 2 with degree 5 (lookup: 2 + input 2 + table 1)
H2_HD int perm_n_columns(const Shape &sh) { return 3 * sh.n_sets() + 2; }
H2_HD int perm_n_sets(const Shape &sh) { return (perm_n_columns(sh) + PERM_CHUNK - 1) / PERM_CHUNK; }
// position p of the permutation -> (is_fixed, column index)
H2_HD void perm_column(const Shape &sh, int p, int *is_fixed, int *index) {
    *is_fixed = 0;
    if (p < 3) *index = p;
    else if (p == 3) *index = sh.words_column();
    else if (p == 4) { *is_fixed = 1; *index = 4; }
    else *index = p - 2;  // advice 3, 4, ... 3N-1
}
// advice column -> position in the permutation
H2_HD int perm_pos_of_advice(const Shape &sh, int a) {
    if (a < 3) return a;
    if (a == sh.words_column()) return 3;
    return a + 2;
}

}  // namespace h2
