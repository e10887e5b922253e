// permmap.h — the permutation argument's cycle structure in closed form (host + device).
//
// halo2_proofs v0.3.0 [UPSTREAM] plonk/permutation/keygen.rs Assembly::copy merges the cycles of
// the two cells of every copy constraint, in the order `synthesize` makes them; build_pk turns the
// resulting mapping (cell -> next cell of its cycle) into the sigma columns.  Every copy the chips
// make (`copy_advice`: src/chips/*.rs, key_schedule.rs:138-221, aes128.rs:176-261) constrains a
// FRESHLY assigned cell to an existing one, and Assembly::copy then splices the fresh cell right
// behind the existing one: mapping[dst] = mapping[src], mapping[src] = dst.  So a cycle is its
// root followed by its copies, latest first, each followed by its own copies — and for this
// circuit that has a closed form:
//   * copies between cells of one AES block stay inside the block: one 3 x 1360 successor
//     template (built once on the host by replaying ONE block's 1952 copies), shifted to every
//     block;
//   * the 176 round-key cells are the only cells copied from outside: block b copies round-key
//     byte (r, i) into exactly one cell ("leaf", column a1 of the block), so the cycle of that
//     round-key cell is  R -> leaf(block B-1) -> ... -> leaf(block 0) -> (R's key-schedule copies);
//   * the key schedule's own 640 copies: a 1296-entry table from replaying them once.
// Nothing walks the 2^k-row columns sequentially any more (round 1 replayed 7.5 M copies over 17 M
// cells on one host thread at the first h2aes_perm_* call).
#pragma once
#include <utility>
#include <vector>

#include "layout.h"

namespace h2 {

constexpr uint32_t PT_BLOCK_CELLS = 3 * (uint32_t)AES_ROWS;            // template: (j, q) -> j*1360 + q
constexpr uint32_t PT_KS_CELLS = 3 * (uint32_t)KEY_SCHEDULE_ROWS + WORDS_ROWS;  // (j,row<400) | words row<96
constexpr uint32_t PT_KS_WORDS = 3 * (uint32_t)KEY_SCHEDULE_ROWS;

struct PermTemplate {
    uint16_t blk[PT_BLOCK_CELLS];  // successor inside the block: j' * 1360 + q'
    uint16_t ks[PT_KS_CELLS];      // successor among the key-schedule cells (compact ids)
};

// compact id of the key-schedule cell holding byte i of round key r (round_key_cell)
H2_HD uint32_t pt_rk_cell(int r, int i) {
    return r == 0 ? PT_KS_WORDS + (uint32_t)i : 2u * (uint32_t)KEY_SCHEDULE_ROWS + 40u * (uint32_t)(r - 1) + 8u + (uint32_t)i;
}
// is compact key-schedule cell `cid` a round-key cell?  -> 16 r + i
H2_HD bool pt_is_rk_cell(uint32_t cid, uint32_t *rk) {
    if (cid >= PT_KS_WORDS) {
        if (cid < PT_KS_WORDS + 16) { *rk = cid - PT_KS_WORDS; return true; }
        return false;
    }
    if (cid < 2u * (uint32_t)KEY_SCHEDULE_ROWS) return false;  // columns a0, a1
    const uint32_t row = cid - 2u * (uint32_t)KEY_SCHEDULE_ROWS, o = row % 40;
    if (o < 8 || o >= 24) return false;
    *rk = 16 * (row / 40 + 1) + (o - 8);
    return true;
}
// block-relative row of the a1 cell that copies round-key byte rk = 16 r + i (aes128.rs:194-198,
// :250-261): the xor rows' y input
H2_HD uint32_t pt_leaf_row(uint32_t rk) {
    const uint32_t r = rk >> 4, i = rk & 15;
    return r == 0 ? 16 + i : (r == 10 ? 1344 + i : 32 + 144 * (r - 1) + 128 + i);
}
H2_HD bool pt_is_leaf_row(uint32_t q, uint32_t *rk) {
    if (q >= 16 && q < 32) { *rk = q - 16; return true; }
    if (q >= 1344) { *rk = 160 + (q - 1344); return true; }
    if (q < 32) return false;
    const uint32_t r = (q - 32) / 144 + 1, rr = (q - 32) % 144;
    if (r > 9 || rr < 128) return false;
    *rk = 16 * r + (rr - 128);
    return true;
}
// does copy `idx` (0..1951) of a block read a round-key cell?  (block_copy, layout.h)
H2_HD bool pt_block_copy_from_rk(uint32_t idx) {
    if (idx < 32) return idx & 1;
    idx -= 32;
    if (idx < 9 * 208) {
        const uint32_t o = idx % 208;
        return o >= 176 && ((o - 176) & 1);
    }
    idx -= 9 * 208;
    return idx >= 16 && ((idx - 16) & 1);
}

// compact key-schedule cell -> (permutation position, row)
H2_HD void pt_ks_global(const Shape &sh, uint32_t cid, uint32_t *pos, uint64_t *row) {
    if (cid >= PT_KS_WORDS) { *pos = 3; *row = cid - PT_KS_WORDS; return; }
    *pos = cid / (uint32_t)KEY_SCHEDULE_ROWS;  // advice 0..2 of set 0 sit at positions 0..2
    *row = cid % (uint32_t)KEY_SCHEDULE_ROWS;
}

// mapping[(pos, row)] for a circuit holding n_blocks AES blocks
H2_HD void perm_successor(const Shape &sh, uint64_t n_blocks, const PermTemplate &t, uint32_t pos,
                          uint64_t row, uint32_t *opos, uint64_t *orow) {
    *opos = pos;
    *orow = row;
    int is_fixed, a;
    perm_column(sh, (int)pos, &is_fixed, &a);
    if (is_fixed) return;  // round_constants: never copied
    uint32_t cid = ~0u;    // compact key-schedule cell, if (pos,row) is one
    if (a == sh.words_column()) {
        if (row >= (uint64_t)WORDS_ROWS) return;
        cid = PT_KS_WORDS + (uint32_t)row;
    } else {
        const int s = a / 3, j = a % 3;
        if (s == 0 && row < KEY_SCHEDULE_ROWS) {
            cid = (uint32_t)j * (uint32_t)KEY_SCHEDULE_ROWS + (uint32_t)row;
        } else {
            if (sh.kind != KIND_AES) return;
            const uint64_t base = sh.set_base(s);
            if (row < base) return;
            const uint64_t rel = row - base, b = rel / AES_ROWS;
            if (b >= sh.blocks_in_set(s, n_blocks)) return;
            const uint32_t q = (uint32_t)(rel - b * AES_ROWS);
            uint32_t rk;
            if (j == 1 && pt_is_leaf_row(q, &rk)) {
                // the chain of the copies of round-key byte rk: towards the previous block
                const uint64_t bg = sh.first_block_of_set(s) + b;
                if (bg) {
                    int ps;
                    uint64_t pbase;
                    block_position(sh, bg - 1, &ps, &pbase);
                    *opos = (uint32_t)perm_pos_of_advice(sh, 3 * ps + 1);
                    *orow = pbase + pt_leaf_row(rk);
                } else {  // block 0 closes into the round-key cell's key-schedule cycle
                    pt_ks_global(sh, t.ks[pt_rk_cell((int)(rk >> 4), (int)(rk & 15))], opos, orow);
                }
                return;
            }
            const uint32_t nx = t.blk[(uint32_t)j * (uint32_t)AES_ROWS + q];
            *opos = (uint32_t)perm_pos_of_advice(sh, 3 * s + (int)(nx / (uint32_t)AES_ROWS));
            *orow = base + b * AES_ROWS + nx % (uint32_t)AES_ROWS;
            return;
        }
    }
    // a key-schedule cell
    uint32_t rk;
    if (n_blocks && sh.kind == KIND_AES && pt_is_rk_cell(cid, &rk)) {  // its latest copy: the last block's
        int ls;
        uint64_t lbase;
        block_position(sh, n_blocks - 1, &ls, &lbase);
        *opos = (uint32_t)perm_pos_of_advice(sh, 3 * ls + 1);
        *orow = lbase + pt_leaf_row(rk);
        return;
    }
    pt_ks_global(sh, t.ks[cid], opos, orow);
}

// Assembly::copy on a compact index space (host; a few thousand cells)
struct PtUnion {
    std::vector<uint32_t> mapping, aux, sizes;
    explicit PtUnion(uint32_t cells) : mapping(cells), aux(cells), sizes(cells, 1u) {
        for (uint32_t i = 0; i < cells; ++i) mapping[i] = aux[i] = i;
    }
    void copy(uint32_t left, uint32_t right) {
        uint32_t lc = aux[left], rc = aux[right];
        if (lc == rc) return;
        if (sizes[lc] < sizes[rc]) std::swap(lc, rc);
        sizes[lc] += sizes[rc];
        uint32_t i = rc;
        do {
            aux[i] = lc;
            i = mapping[i];
        } while (i != rc);
        std::swap(mapping[left], mapping[right]);
    }
};
inline void perm_template_build(const Shape &sh, PermTemplate *t) {
    {   // one AES block, its copies from round-key cells left out (they are the leaf chains)
        PtUnion u(PT_BLOCK_CELLS);
        for (uint32_t idx = 0; idx < COPIES_PER_BLOCK; ++idx) {
            if (pt_block_copy_from_rk(idx)) continue;
            const CopyPair cp = block_copy(sh, 0, 0, idx);
            u.copy(cp.dst.col * (uint32_t)AES_ROWS + (uint32_t)cp.dst.row,
                   cp.src.col * (uint32_t)AES_ROWS + (uint32_t)cp.src.row);
        }
        for (uint32_t i = 0; i < PT_BLOCK_CELLS; ++i) t->blk[i] = (uint16_t)u.mapping[i];
    }
    {   // the key schedule
        PtUnion u(PT_KS_CELLS);
        const uint32_t wcol = (uint32_t)sh.words_column();
        auto cid = [&](const CellRef &c) {
            return c.col == wcol ? PT_KS_WORDS + (uint32_t)c.row
                                 : c.col * (uint32_t)KEY_SCHEDULE_ROWS + (uint32_t)c.row;
        };
        for (uint32_t idx = 0; idx < COPIES_KEY_SCHEDULE; ++idx) {
            const CopyPair cp = ks_copy(sh, idx);
            u.copy(cid(cp.dst), cid(cp.src));
        }
        for (uint32_t i = 0; i < PT_KS_CELLS; ++i) t->ks[i] = (uint16_t)u.mapping[i];
    }
}

}  // namespace h2
