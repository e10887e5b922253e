// TEST INFRASTRUCTURE — CPU oracle for the halo2-aes witness + lookup-helper-column path.
//
// This file is a CPU restatement of what the reference computes on this path, written region by
// region the way the reference's `synthesize` drives halo2 (two passes per region under
// SimpleFloorPlanner: a shape pass that records the touched columns, then the assignment pass).
// It is deliberately NOT the closed-form layout the CUDA product uses: the two are independent
// derivations of every (column,row,value), and tests/ compare them.
//
// What it follows (paths relative to /root/reference):
//   * src/table.rs:18-192          load_enc_full_table
//   * src/key_schedule.rs:80-224   schedule_keys / assign_first_round / assign_round
//   * src/aes128.rs:143-356        schedule_key / encrypt / lcon / aes_callable + getters
//   * src/chips/*.rs               xor / substitute / mul / range_check region bodies
//   * src/utils.rs:8-33            xor_bytes / sub_byte / ROUND_CONSTANT
//   * src/constant.rs              S_BOX (including its deviation from FIPS-197 at index 255),
//                                  MUL_BY_2, MUL_BY_3, KEY_SCHEDULE_ROWS, AES_ROWS
// Third-party behaviour that is NOT in /root/reference (un-vendored git dependency
// halo2_proofs v0.3.0 @ 73408a140737d8336490452193b21f5a7a94e7de, Cargo.lock:741-743, and
// halo2curves 0.6.1) is restated from its published algorithm and marked [UPSTREAM]:
//   * circuit/floor_planner/single_pass.rs  region placement = max(next free row of touched cols)
//   * circuit/table_layouter.rs + keygen fill_from_row  table padding with the offset-0 value
//   * plonk/lookup/prover.rs  commit_permuted / permute_expression_pair / commit_product
//   * dev.rs (MockProver::verify)  gate / lookup / permutation checks
// PARITY STATUS: witness/fixed/selector values are pinned by the reference's own vectors
// (EXPANDED key_schedule.rs:337-345, MockProver-satisfied zero/zero circuits, xor 5^12=9) plus
// FIPS-197 App. B / C.1; the permuted/Z columns are "parity unpinned" — no reference test holds
// their values, only algebraic invariants are available.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "fr.hpp"
#include "g1.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------
// constants (src/constant.rs, src/utils.rs:28, src/table.rs:10-16)
// ---------------------------------------------------------------------------------------------
static const uint64_t KEY_SCHEDULE_ROWS = 1760;  // constant.rs:113
static const uint64_t AES_ROWS = 1360;           // constant.rs:114
static const uint64_t ROUND_CONSTANT[10] = {1, 2, 4, 8, 16, 32, 64, 128, 27, 54};  // utils.rs:28
enum Tag { TAG_U8 = 1, TAG_XOR = 2, TAG_SBOX = 3, TAG_GFMUL2 = 4, TAG_GFMUL3 = 5 };  // table.rs:10

struct ByteTables {
    uint8_t sbox[256], mul2[256], mul3[256];
};

static uint8_t gf_xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x & 0x80) ? 0x1b : 0)); }
static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; ++i) {
        if (b & 1) p ^= a;
        a = gf_xtime(a);
        b >>= 1;
    }
    return p;
}

// The tables are generated from first principles (GF(2^8) inverse + affine map; xtime), then —
// unless `fips` — entry 255 of the S-box is set to 23, which is what src/constant.rs:14 holds
// (FIPS-197 says 0x16 = 22).  tests/test_oracle_golden.py checks all three tables byte for byte
// against tests/golden/ref_tables.json, extracted from the reference's constant.rs.
static ByteTables make_tables(bool fips) {
    ByteTables t;
    for (int x = 0; x < 256; ++x) {
        uint8_t inv = 0;
        if (x)
            for (int y = 1; y < 256; ++y)
                if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv;
        for (int i = 1; i <= 4; ++i) s ^= (uint8_t)((inv << i) | (inv >> (8 - i)));
        t.sbox[x] = s ^ 0x63;
        t.mul2[x] = gf_xtime((uint8_t)x);
        t.mul3[x] = gf_xtime((uint8_t)x) ^ (uint8_t)x;
    }
    if (!fips) t.sbox[255] = 23;
    return t;
}

// ---------------------------------------------------------------------------------------------
// circuit shapes
// ---------------------------------------------------------------------------------------------
enum Kind { KIND_AES = 0, KIND_KEYSCHED = 1 };
enum ErrorCode {
    ORC_OK = 0,
    ORC_ERR_ARG = 1,
    ORC_ERR_CAPACITY = 2,         // aes128.rs:160-162 panic!("AES calls too many...")
    ORC_ERR_NOT_ENOUGH_ROWS = 3,  // [UPSTREAM] Error::NotEnoughRowsAvailable
    ORC_ERR_KEYS_NOT_SCHEDULED = 4,  // aes128.rs:170 expect("Keys should be scheduled")
    ORC_ERR_LOOKUP_INPUT_NOT_IN_TABLE = 5,  // [UPSTREAM] Error::ConstraintSystemFailure
    ORC_ERR_FROM_BYTES = 6,                 // utils.rs:18 from_bytes(..).unwrap()
};

struct OrcError : std::runtime_error {
    int code;
    OrcError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct Shape {
    int kind, k, N;
    size_t n() const { return (size_t)1 << k; }
    size_t usable() const { return n() - 6; }  // blinding_factors()=5, +1  [UPSTREAM]
    int n_sets() const { return kind == KIND_AES ? N : 1; }
    int n_advice() const { return 3 * n_sets() + 1; }
    int n_fixed() const { return 5; }  // 4 table columns + round_constants
    int n_complex_sel() const { return kind == KIND_AES ? 5 * N : 3; }
    int n_selectors() const { return n_complex_sel() + 1; }  // + q_eq_rcon
    int n_lookups() const { return n_complex_sel(); }
    // selector/lookup index of op t (0 range,1 xor,2 sbox,3 mul2,4 mul3) in column set i:
    //   aes128.rs:63-115 allocates them per set in this order; the key-schedule test/bench
    //   circuit (key_schedule.rs:267-283, benches/key_schedule.rs:47-58) only has the first 3.
    int sel(int set, int t) const { return kind == KIND_AES ? 5 * set + t : t; }
    int sel_rcon() const { return n_complex_sel(); }
    int words_column() const { return 3 * n_sets(); }  // key_schedule.rs:48, allocated last
    int rc_column() const { return 4; }                // key_schedule.rs:49, after 4 table cols
};

struct AssignedCell {
    int column;  // advice column index
    size_t row;
    Fr value;
    bool known;  // false during the shape pass (Value::unknown())
};

struct Copy {
    uint32_t col_a;  // advice column index (all copies on this path are advice<->advice)
    uint64_t row_a;
    uint32_t col_b;
    uint64_t row_b;
};

class Synth;

// One region as seen by an assignment closure.  [UPSTREAM] RegionShape (shape pass) and
// SingleChipLayouterRegion (assignment pass).
class Region {
  public:
    Synth *s;
    bool shape_pass;
    size_t start = 0;
    std::vector<int> cols;  // RegionColumn ids touched (shape pass)
    size_t row_count = 0;

    void touch(int region_col, size_t offset);
    void enable_selector(int sel, size_t offset);
    AssignedCell assign_advice(int col, size_t offset, const std::function<Fr()> &to);
    void assign_fixed(int col, size_t offset, const std::function<Fr()> &to);
    AssignedCell copy_advice(const AssignedCell &src, int col, size_t offset);
};

class Synth {
  public:
    Shape sh;
    ByteTables tb;
    size_t n, usable;
    std::vector<std::vector<Fr>> advice, fixed;
    std::vector<std::vector<uint8_t>> selectors;
    std::vector<Copy> copies;
    std::vector<size_t> next_free;  // per RegionColumn: advice | fixed | selector
    bool record_copies = true;
    uint64_t n_regions = 0;

    Synth(const Shape &shape, bool fips) : sh(shape), tb(make_tables(fips)) {
        n = sh.n();
        usable = sh.usable();
        advice.assign(sh.n_advice(), std::vector<Fr>(n, FR_ZERO));
        fixed.assign(sh.n_fixed(), std::vector<Fr>(n, FR_ZERO));
        selectors.assign(sh.n_selectors(), std::vector<uint8_t>(n, 0));
        next_free.assign(sh.n_advice() + sh.n_fixed() + sh.n_selectors(), 0);
    }
    int rc_advice(int c) const { return c; }
    int rc_fixed(int c) const { return sh.n_advice() + c; }
    int rc_selector(int c) const { return sh.n_advice() + sh.n_fixed() + c; }

    void check_row(size_t row) const {
        if (row >= usable)
            throw OrcError(ORC_ERR_NOT_ENOUGH_ROWS, "row outside usable rows");
    }

    // [UPSTREAM] SingleChipLayouter::assign_region: run the closure on a RegionShape, place the
    // region at the max next-free row of the columns it touched, then run it again for real.
    template <class F>
    auto assign_region(F &&f) -> decltype(f(std::declval<Region &>())) {
        Region shape;
        shape.s = this;
        shape.shape_pass = true;
        f(shape);
        size_t region_start = 0;
        for (int c : shape.cols) region_start = std::max(region_start, next_free[c]);
        for (int c : shape.cols) next_free[c] = region_start + shape.row_count;
        Region real;
        real.s = this;
        real.shape_pass = false;
        real.start = region_start;
        ++n_regions;
        return f(real);
    }
};

void Region::touch(int region_col, size_t offset) {
    if (shape_pass) {
        if (std::find(cols.begin(), cols.end(), region_col) == cols.end())
            cols.push_back(region_col);
        row_count = std::max(row_count, offset + 1);
    }
}
void Region::enable_selector(int sel, size_t offset) {
    if (shape_pass) return touch(s->rc_selector(sel), offset);
    s->check_row(start + offset);
    s->selectors[sel][start + offset] = 1;
}
AssignedCell Region::assign_advice(int col, size_t offset, const std::function<Fr()> &to) {
    if (shape_pass) {
        touch(s->rc_advice(col), offset);
        return AssignedCell{col, offset, FR_ZERO, false};
    }
    s->check_row(start + offset);
    Fr v = to();
    s->advice[col][start + offset] = v;
    return AssignedCell{col, start + offset, v, true};
}
void Region::assign_fixed(int col, size_t offset, const std::function<Fr()> &to) {
    if (shape_pass) return touch(s->rc_fixed(col), offset);
    s->check_row(start + offset);
    s->fixed[col][start + offset] = to();
}
// [UPSTREAM] AssignedCell::copy_advice = assign_advice(value) + region.constrain_equal(cells).
// constrain_equal does not affect the region shape.
AssignedCell Region::copy_advice(const AssignedCell &src, int col, size_t offset) {
    AssignedCell c = assign_advice(col, offset, [&] { return src.value; });
    if (!shape_pass && s->record_copies)
        s->copies.push_back(Copy{(uint32_t)c.column, c.row, (uint32_t)src.column, src.row});
    return c;
}

// ---------------------------------------------------------------------------------------------
// byte helpers (src/utils.rs)
// ---------------------------------------------------------------------------------------------
// utils.rs:8-19: XOR all 32 canonical bytes, then Fp::from_bytes(..).unwrap().
static Fr xor_bytes(const Fr &x, const Fr &y) {
    uint64_t a[4], b[4], c[4];
    fr_to_canonical(x, a);
    fr_to_canonical(y, b);
    for (int i = 0; i < 4; ++i) c[i] = a[i] ^ b[i];
    Fr out;
    if (!fr_from_canonical(c, &out)) throw OrcError(ORC_ERR_FROM_BYTES, "from_bytes None");
    return out;
}
static uint8_t first_byte(const Fr &x) {  // `*v.to_bytes().first().unwrap()`
    uint64_t a[4];
    fr_to_canonical(x, a);
    return (uint8_t)(a[0] & 0xff);
}
// utils.rs:22-24
static Fr sub_byte(const ByteTables &tb, const Fr &x) { return fr_from_u64(tb.sbox[first_byte(x)]); }

// ---------------------------------------------------------------------------------------------
// chips (src/chips/*.rs) — each call is one 1-row region
// ---------------------------------------------------------------------------------------------
struct ChipCols {
    int a0, a1, a2;  // advices[set][0..3]
    int q_range, q_xor, q_sbox, q_mul2, q_mul3;
};

// u8_xor_chip.rs:63-100
static AssignedCell chip_xor(Synth &s, const ChipCols &c, const AssignedCell &x,
                             const AssignedCell &y) {
    return s.assign_region([&](Region &r) {
        r.enable_selector(c.q_xor, 0);
        AssignedCell xc = r.copy_advice(x, c.a0, 0);
        AssignedCell yc = r.copy_advice(y, c.a1, 0);
        return r.assign_advice(c.a2, 0, [&] { return xor_bytes(xc.value, yc.value); });
    });
}
// sbox_chip.rs:57-83
static AssignedCell chip_sbox(Synth &s, const ChipCols &c, const AssignedCell &x) {
    return s.assign_region([&](Region &r) {
        r.enable_selector(c.q_sbox, 0);
        AssignedCell xc = r.copy_advice(x, c.a0, 0);
        return r.assign_advice(c.a1, 0, [&] { return sub_byte(s.tb, xc.value); });
    });
}
// gf_mul_chip.rs:59-89 (macro instantiated for MUL_BY_2 / MUL_BY_3)
static AssignedCell chip_mul(Synth &s, const ChipCols &c, int which, const AssignedCell &x) {
    return s.assign_region([&](Region &r) {
        r.enable_selector(which == 2 ? c.q_mul2 : c.q_mul3, 0);
        r.copy_advice(x, c.a0, 0);
        return r.assign_advice(c.a1, 0, [&] {
            const uint8_t *dict = which == 2 ? s.tb.mul2 : s.tb.mul3;
            return fr_from_u64(dict[first_byte(x.value)]);
        });
    });
}
// u8_range_check_chip.rs:51-70
static void chip_range(Synth &s, const ChipCols &c, const AssignedCell &x) {
    s.assign_region([&](Region &r) {
        r.enable_selector(c.q_range, 0);
        r.copy_advice(x, c.a0, 0);
        return 0;
    });
}

// ---------------------------------------------------------------------------------------------
// table (src/table.rs:18-192) + [UPSTREAM] SimpleTableLayouter default/fill_from_row
// ---------------------------------------------------------------------------------------------
static void load_enc_full_table(Synth &s) {
    auto &T = s.fixed;  // table columns are fixed columns 0..3
    size_t first_unused = 256 + 256 + 65536 + 256 + 256 + 1;
    if (first_unused > s.usable) throw OrcError(ORC_ERR_NOT_ENOUGH_ROWS, "table needs k>=17");
    auto put = [&](size_t pos, uint64_t t0, uint64_t t1, uint64_t t2, uint64_t t3) {
        T[0][pos] = fr_from_u64(t0);
        T[1][pos] = fr_from_u64(t1);
        T[2][pos] = fr_from_u64(t2);
        T[3][pos] = fr_from_u64(t3);
    };
    size_t offset = 0;
    for (int i = 0; i < 256; ++i) put(offset + i, TAG_U8, i, 0, 0);  // :23-50
    offset += 256;
    for (int i = 0; i < 256; ++i) put(offset + i, TAG_SBOX, i, s.tb.sbox[i], 0);  // :53-81
    offset += 256;
    size_t l = offset;
    for (int i = 0; i < 256; ++i)
        for (int j = 0; j < 256; ++j) put(l++, TAG_XOR, i, j, i ^ j);  // :84-115
    offset += 65536;
    for (int i = 0; i < 256; ++i) put(offset + i, TAG_GFMUL2, i, s.tb.mul2[i], 0);  // :118-145
    offset += 256;
    for (int i = 0; i < 256; ++i) put(offset + i, TAG_GFMUL3, i, s.tb.mul3[i], 0);  // :148-175
    offset += 256;
    put(offset, 0, 0, 0, 0);  // :178-188 "assign zero row"
    // [UPSTREAM] every table column is padded over the remaining usable rows with the value
    // assigned at offset 0 of that column.
    for (int c = 0; c < 4; ++c) {
        Fr def = T[c][0];
        for (size_t row = first_unused; row < s.usable; ++row) T[c][row] = def;
    }
}

// ---------------------------------------------------------------------------------------------
// key schedule (src/key_schedule.rs)
// ---------------------------------------------------------------------------------------------
typedef std::vector<AssignedCell> Cells;

static Cells ks_assign_first_round(Synth &s, const uint8_t key[16]) {  // :98-118
    int w = s.sh.words_column();
    return s.assign_region([&](Region &r) {
        Cells words;
        for (int i = 0; i < 16; ++i)
            words.push_back(r.assign_advice(w, i, [&] { return fr_from_u64(key[i]); }));
        return words;
    });
}

static Cells ks_assign_round(Synth &s, const ChipCols &c, int round, const Cells &prev) {  // :122-224
    int w = s.sh.words_column();
    Cells words;
    Cells shifted = s.assign_region([&](Region &r) {  // :138-154
        static const int order[4] = {13, 14, 15, 12};
        Cells out;
        for (int i = 0; i < 4; ++i) out.push_back(r.copy_advice(prev[order[i]], w, i));
        return out;
    });
    Cells subbed;
    for (auto &b : shifted) subbed.push_back(chip_sbox(s, c, b));  // :156-159
    Fr rc = fr_from_u64(ROUND_CONSTANT[round - 1]);                // :161, utils.rs:31-33
    Cells rc_assigned = s.assign_region([&](Region &r) {           // :162-188
        Cells res;
        r.enable_selector(s.sh.sel_rcon(), 0);
        r.assign_fixed(s.sh.rc_column(), 0, [&] { return rc; });
        res.push_back(r.assign_advice(w, 0, [&] { return rc; }));
        for (int i = 0; i < 3; ++i)
            res.push_back(r.assign_advice(w, i + 1, [&] { return fr_from_u64(0); }));
        return res;
    });
    Cells rconned;
    for (int i = 0; i < 4; ++i) rconned.push_back(chip_xor(s, c, subbed[i], rc_assigned[i]));  // :190-194
    Cells next_word;
    for (int i = 0; i < 4; ++i) next_word.push_back(chip_xor(s, c, prev[i], rconned[i]));  // :197-202
    words.insert(words.end(), next_word.begin(), next_word.end());
    for (int i = 1; i < 4; ++i) {  // :207-216
        Cells nw;
        for (int j = 0; j < 4; ++j) nw.push_back(chip_xor(s, c, prev[4 * i + j], next_word[j]));
        next_word = nw;
        words.insert(words.end(), next_word.begin(), next_word.end());
    }
    for (auto &b : words) chip_range(s, c, b);  // :218-221
    return words;
}

static std::vector<Cells> schedule_keys(Synth &s, const ChipCols &c, const uint8_t key[16]) {  // :80-96
    std::vector<Cells> words;
    Cells round = ks_assign_first_round(s, key);
    words.push_back(round);
    for (int i = 1; i <= 10; ++i) {
        round = ks_assign_round(s, c, i, round);
        words.push_back(round);
    }
    return words;
}

// ---------------------------------------------------------------------------------------------
// FixedAes128Config (src/aes128.rs)
// ---------------------------------------------------------------------------------------------
struct AesConfig {
    Synth *s;
    bool have_keys = false;
    std::vector<Cells> keys;
    size_t current = 0;  // :36-39
    uint64_t count = 0;  // :41-42

    ChipCols cols(size_t set) const {
        const Shape &sh = s->sh;
        int i = (int)set;
        return ChipCols{3 * i, 3 * i + 1, 3 * i + 2, sh.sel(i, 0), sh.sel(i, 1),
                        sh.sel(i, 2), sh.sel(i, 3), sh.sel(i, 4)};
    }

    void schedule_key(const uint8_t key[16]) {  // :143-152; key schedule uses set 0 (:118-124)
        keys = schedule_keys(*s, cols(0), key);
        have_keys = true;
    }

    bool aes_callable() {  // :303-325
        uint64_t max_row = (uint64_t)1 << s->sh.k;
        if (current == 0) max_row -= KEY_SCHEDULE_ROWS;
        if (max_row >= count * AES_ROWS + AES_ROWS) return true;
        if (current < (size_t)s->sh.N - 1) {
            current += 1;
            count = 0;
            return true;
        }
        return false;
    }

    AssignedCell lcon(const Cells &word, const int coeffs[4]) {  // :268-301
        ChipCols c = cols(current);
        Cells tmp;
        for (int t = 0; t < 4; ++t) {
            switch (coeffs[t]) {
            case 1:
                tmp.push_back(s->assign_region(
                    [&](Region &r) { return r.copy_advice(word[t], c.a0, 0); }));
                break;
            case 2: tmp.push_back(chip_mul(*s, c, 2, word[t])); break;
            case 3: tmp.push_back(chip_mul(*s, c, 3, word[t])); break;
            default: throw OrcError(ORC_ERR_ARG, "col should be 1, 2, or 3.");
            }
        }
        AssignedCell inter_1 = chip_xor(*s, c, tmp[0], tmp[1]);
        AssignedCell inter_2 = chip_xor(*s, c, tmp[2], tmp[3]);
        return chip_xor(*s, c, inter_1, inter_2);
    }

    Cells encrypt(const uint8_t plaintext[16]) {  // :154-265
        if (!aes_callable())
            throw OrcError(ORC_ERR_CAPACITY, "AES calls too many. doesn't fit in the rows");
        count += 1;
        if (!have_keys) throw OrcError(ORC_ERR_KEYS_NOT_SCHEDULED, "Keys should be scheduled");
        ChipCols c = cols(current);
        Cells assigned_plaintext = s->assign_region([&](Region &r) {  // :176-192
            Cells out;
            for (int i = 0; i < 16; ++i)
                out.push_back(r.assign_advice(c.a0, i, [&] { return fr_from_u64(plaintext[i]); }));
            return out;
        });
        Cells prev_round;
        for (int i = 0; i < 16; ++i)
            prev_round.push_back(chip_xor(*s, c, assigned_plaintext[i], keys[0][i]));  // :194-198
        static const int matrix[4][4] = {{2, 3, 1, 1}, {1, 2, 3, 1}, {1, 1, 2, 3}, {3, 1, 1, 2}};
        for (int no_round = 1; no_round < 11; ++no_round) {  // :201-262
            Cells subbed_flat;
            for (auto &b : prev_round) subbed_flat.push_back(chip_sbox(*s, c, b));  // :203-209
            std::vector<Cells> shifted(4);  // :216-223
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) shifted[i].push_back(subbed_flat[4 * ((i + j) % 4) + j]);
            std::vector<Cells> mixed;
            if (no_round == 10) {
                mixed = shifted;
            } else {
                for (auto &word : shifted) {  // :239-248
                    Cells out;
                    for (int col = 0; col < 4; ++col) out.push_back(lcon(word, matrix[col]));
                    mixed.push_back(out);
                }
            }
            Cells next;
            for (int i = 0; i < 4; ++i)  // :250-261
                for (int j = 0; j < 4; ++j)
                    next.push_back(chip_xor(*s, c, mixed[i][j], keys[no_round][i * 4 + j]));
            prev_round = next;
        }
        return prev_round;
    }
};

// ---------------------------------------------------------------------------------------------
// lookup arguments configured on this path, in `meta.lookup` order
// ---------------------------------------------------------------------------------------------
struct LookupDef {
    int selector;   // complex selector index q
    uint64_t tag;   // q * Fp::from(Tag)
    int m;          // number of (input, table) pairs
    int adv[3];     // advice columns of x, y, z (m-1 of them used)
};

static LookupDef lookup_def(const Shape &sh, int idx) {
    int set = sh.kind == KIND_AES ? idx / 5 : 0;
    int t = sh.kind == KIND_AES ? idx % 5 : idx;
    LookupDef d;
    d.selector = sh.sel(set, t);
    d.adv[0] = 3 * set;
    d.adv[1] = 3 * set + 1;
    d.adv[2] = 3 * set + 2;
    switch (t) {
    case 0: d.tag = TAG_U8; d.m = 2; break;      // u8_range_check_chip.rs:35-43
    case 1: d.tag = TAG_XOR; d.m = 4; break;     // u8_xor_chip.rs:41-53
    case 2: d.tag = TAG_SBOX; d.m = 3; break;    // sbox_chip.rs:38-48
    case 3: d.tag = TAG_GFMUL2; d.m = 3; break;  // gf_mul_chip.rs:40-50
    default: d.tag = TAG_GFMUL3; d.m = 3; break;
    }
    return d;
}

struct Columns {  // a view over n-row columns, owned by a Synth or supplied by the caller
    Shape sh;
    std::vector<const Fr *> advice, fixed;
    std::vector<const uint8_t *> selectors;
};

static Columns view_of(const Synth &s) {
    Columns c;
    c.sh = s.sh;
    for (auto &v : s.advice) c.advice.push_back(v.data());
    for (auto &v : s.fixed) c.fixed.push_back(v.data());
    for (auto &v : s.selectors) c.selectors.push_back(v.data());
    return c;
}

// canonical (non-Montgomery) value as a sortable key: [UPSTREAM] `Ord for Fr`
struct Key {
    uint64_t c[4];
    bool operator<(const Key &o) const {
        for (int i = 3; i >= 0; --i) {
            if (c[i] != o.c[i]) return c[i] < o.c[i];
        }
        return false;
    }
    bool operator==(const Key &o) const { return !memcmp(c, o.c, 32); }
};
static Key key_of(const Fr &a) {
    Key k;
    fr_to_canonical(a, k.c);
    return k;
}
static Fr fr_of(const Key &k) {
    Fr a = {{k.c[0], k.c[1], k.c[2], k.c[3]}};
    return fr_mul(a, FR_R2);
}

// [UPSTREAM] plonk/lookup/prover.rs Argument::commit_permuted (compress_expressions) :
//   compressed = fold(expressions, 0, |acc, e| acc * theta + e)  over all n rows.
// Input expressions here are q*tag, q*x[, q*y[, q*z]] with q the selector's 0/1 fixed column.
static void compress_lookup(const Columns &c, const LookupDef &d, const Fr &theta,
                            std::vector<Fr> &A, std::vector<Fr> &S) {
    size_t n = c.sh.n();
    A.assign(n, FR_ZERO);
    S.assign(n, FR_ZERO);
    Fr tag = fr_from_u64(d.tag);
    for (size_t row = 0; row < n; ++row) {
        Fr q = c.selectors[d.selector][row] ? FR_ONE : FR_ZERO;
        Fr acc = fr_mul(q, tag);  // Expression::Scaled(q, tag); first fold step: 0*theta + e0
        Fr tacc = c.fixed[0][row];
        for (int j = 0; j + 1 < d.m; ++j) {
            Fr e = fr_mul(q, c.advice[d.adv[j]][row]);  // Expression::Product(q, advice)
            acc = fr_add(fr_mul(acc, theta), e);
            tacc = fr_add(fr_mul(tacc, theta), c.fixed[1 + j][row]);
        }
        A[row] = acc;
        S[row] = tacc;
    }
}

// [UPSTREAM] plonk/lookup/prover.rs permute_expression_pair.  Rows >= usable (blinding rows,
// random in halo2) are left zero here.
static int permute_expression_pair(const Shape &sh, const std::vector<Fr> &A,
                                   const std::vector<Fr> &S, std::vector<Fr> &Ap,
                                   std::vector<Fr> &Sp) {
    size_t n = sh.n(), usable = sh.usable();
    std::vector<Key> in(usable);
    for (size_t i = 0; i < usable; ++i) in[i] = key_of(A[i]);
    std::sort(in.begin(), in.end());  // permuted_input_expression.sort()
    std::map<Key, uint32_t> leftover;  // BTreeMap<Scalar, u32>
    for (size_t i = 0; i < usable; ++i) leftover[key_of(S[i])] += 1;
    std::vector<Key> tab(usable);
    std::vector<uint8_t> filled(usable, 0);
    std::vector<size_t> repeated;
    for (size_t row = 0; row < usable; ++row) {
        if (row == 0 || !(in[row] == in[row - 1])) {
            tab[row] = in[row];
            filled[row] = 1;
            auto it = leftover.find(in[row]);
            if (it == leftover.end() || it->second == 0) return ORC_ERR_LOOKUP_INPUT_NOT_IN_TABLE;
            it->second -= 1;
        } else {
            repeated.push_back(row);
        }
    }
    // "Populate permuted table at unfilled rows with leftover table elements": BTreeMap iterated
    // ascending, rows popped from the END of repeated_input_rows.
    for (auto &kv : leftover)
        for (uint32_t t = 0; t < kv.second; ++t) {
            tab[repeated.back()] = kv.first;
            repeated.pop_back();
        }
    if (!repeated.empty()) return ORC_ERR_ARG;
    Ap.assign(n, FR_ZERO);
    Sp.assign(n, FR_ZERO);
    for (size_t i = 0; i < usable; ++i) {
        Ap[i] = fr_of(in[i]);
        Sp[i] = fr_of(tab[i]);
    }
    return ORC_OK;
}

// [UPSTREAM] plonk/lookup/prover.rs Permuted::commit_product.  z has n-5 deterministic entries
// (z[0]=1 ... z[n-6]); the 5 blinding entries are left zero here.
static void commit_product(const Shape &sh, const std::vector<Fr> &A, const std::vector<Fr> &S,
                           const std::vector<Fr> &Ap, const std::vector<Fr> &Sp, const Fr &beta,
                           const Fr &gamma, std::vector<Fr> &Z) {
    size_t n = sh.n();
    std::vector<Fr> prod(n);
    for (size_t i = 0; i < n; ++i) prod[i] = fr_mul(fr_add(beta, Ap[i]), fr_add(gamma, Sp[i]));
    fr_batch_invert(prod);
    for (size_t i = 0; i < n; ++i) {
        prod[i] = fr_mul(prod[i], fr_add(A[i], beta));
        prod[i] = fr_mul(prod[i], fr_add(S[i], gamma));
    }
    Z.assign(n, FR_ZERO);
    Fr state = FR_ONE;
    Z[0] = state;
    for (size_t i = 1; i < n - 5; ++i) {
        state = fr_mul(state, prod[i - 1]);
        Z[i] = state;
    }
}

// ---------------------------------------------------------------------------------------------
// [UPSTREAM] permutation argument (plonk/permutation/keygen.rs Assembly + prover.rs commit),
// halo2_proofs v0.3.0 without the "thread-safe-region" feature (the reference enables only
// derive_serde, dev-graph, cost-estimator: Cargo.toml:14-18).  SURVEY 8(f) item 1.
// ---------------------------------------------------------------------------------------------
// halo2curves 0.6.1 bn256::Fr constants [UPSTREAM]: MULTIPLICATIVE_GENERATOR = 7, S = 28,
// ROOT_OF_UNITY = 7^((r-1)/2^28), DELTA = 7^(2^28)  (both re-derived from 7 in
// tests/test_oracle_golden.py).
static const uint64_t ROOT_OF_UNITY_CANON[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull,
                                                0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
static const uint64_t DELTA_CANON[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull,
                                        0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull};

static Fr fr_of_canon(const uint64_t c[4]) {
    Fr o;
    fr_from_canonical(c, &o);
    return o;
}
// EvaluationDomain::new: omega = ROOT_OF_UNITY squared (S - k) times
static Fr domain_omega(int k) {
    Fr w = fr_of_canon(ROOT_OF_UNITY_CANON);
    for (int i = k; i < 28; ++i) w = fr_mul(w, w);
    return w;
}

struct PermColumn {
    int type;   // 0 advice, 1 fixed
    int index;
};
// order of the enable_equality calls: key_schedule.rs:52-57 (advices[0][0..3], words_column,
// enable_constant(round_constants)), then aes128.rs:126-130 (every advice column)
static std::vector<PermColumn> perm_columns(const Shape &sh) {
    std::vector<PermColumn> c;
    for (int j = 0; j < 3; ++j) c.push_back(PermColumn{0, j});
    c.push_back(PermColumn{0, sh.words_column()});
    c.push_back(PermColumn{1, sh.rc_column()});
    if (sh.kind == KIND_AES)
        for (int a = 3; a < 3 * sh.N; ++a) c.push_back(PermColumn{0, a});
    return c;
}
// cs.degree() = max(permutation 3, lookups max(4, 2 + 2 + 1) = 5, gate 2..3) = 5 -> chunk_len 3
static const int PERM_CHUNK = 3;

// [UPSTREAM] permutation::keygen::Assembly
struct Assembly {
    size_t n;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> mapping, aux;
    std::vector<std::vector<uint32_t>> sizes;
    Assembly(size_t n_, size_t ncols) : n(n_) {
        mapping.resize(ncols);
        for (size_t i = 0; i < ncols; ++i) {
            mapping[i].resize(n);
            for (size_t j = 0; j < n; ++j) mapping[i][j] = {(uint32_t)i, (uint32_t)j};
        }
        aux = mapping;
        sizes.assign(ncols, std::vector<uint32_t>(n, 1));
    }
    void copy(uint32_t lc, uint32_t lr, uint32_t rc, uint32_t rr) {
        auto left_cycle = aux[lc][lr], right_cycle = aux[rc][rr];
        if (left_cycle == right_cycle) return;
        if (sizes[left_cycle.first][left_cycle.second] < sizes[right_cycle.first][right_cycle.second])
            std::swap(left_cycle, right_cycle);
        sizes[left_cycle.first][left_cycle.second] += sizes[right_cycle.first][right_cycle.second];
        auto i = right_cycle;
        while (true) {
            aux[i.first][i.second] = left_cycle;
            i = mapping[i.first][i.second];
            if (i == right_cycle) break;
        }
        std::swap(mapping[lc][lr], mapping[rc][rr]);
    }
};

// mapping after all of the circuit's copy constraints, in recording order
static Assembly build_assembly(const Synth &syn, const std::vector<PermColumn> &cols) {
    const Shape &sh = syn.sh;
    std::vector<int> pos_of_advice(sh.n_advice(), -1);
    for (size_t i = 0; i < cols.size(); ++i)
        if (cols[i].type == 0) pos_of_advice[cols[i].index] = (int)i;
    Assembly as(syn.n, cols.size());
    for (const Copy &cp : syn.copies)  // constrain_equal(left = new cell, right = source cell)
        as.copy((uint32_t)pos_of_advice[cp.col_a], (uint32_t)cp.row_a,
                (uint32_t)pos_of_advice[cp.col_b], (uint32_t)cp.row_b);
    return as;
}

// [UPSTREAM] Assembly::build_pk: sigma_i(omega^j) = delta^{i'} omega^{j'}, (i', j') = mapping[i][j]
static void build_sigma(const Synth &syn, const std::vector<PermColumn> &cols,
                        std::vector<std::vector<Fr>> &sigma, std::vector<uint32_t> *map_out) {
    Assembly as = build_assembly(syn, cols);
    const size_t n = syn.n;
    const Fr omega = domain_omega(syn.sh.k), delta = fr_of_canon(DELTA_CANON);
    std::vector<Fr> omega_pow(n), delta_pow(cols.size());
    Fr cur = FR_ONE;
    for (size_t j = 0; j < n; ++j) { omega_pow[j] = cur; cur = fr_mul(cur, omega); }
    cur = FR_ONE;
    for (size_t i = 0; i < cols.size(); ++i) { delta_pow[i] = cur; cur = fr_mul(cur, delta); }
    sigma.assign(cols.size(), std::vector<Fr>(n));
    if (map_out) map_out->assign(2 * cols.size() * n, 0);
    for (size_t i = 0; i < cols.size(); ++i)
        for (size_t j = 0; j < n; ++j) {
            auto m = as.mapping[i][j];
            sigma[i][j] = fr_mul(delta_pow[m.first], omega_pow[m.second]);
            if (map_out) {
                (*map_out)[2 * (i * n + j)] = m.first;
                (*map_out)[2 * (i * n + j) + 1] = m.second;
            }
        }
}

// [UPSTREAM] permutation::Argument::commit: one grand product per chunk of PERM_CHUNK columns,
// chained through last_z = z[n - 6].  The 5 blinding entries of every z are left zero here.
static void permutation_product(const Columns &c, const std::vector<PermColumn> &cols,
                                const std::vector<std::vector<Fr>> &sigma, const Fr &beta,
                                const Fr &gamma, std::vector<std::vector<Fr>> &Z) {
    const size_t n = c.sh.n();
    const Fr omega = domain_omega(c.sh.k), delta = fr_of_canon(DELTA_CANON);
    Fr deltaomega0 = FR_ONE, last_z = FR_ONE;
    Z.clear();
    for (size_t base = 0; base < cols.size(); base += PERM_CHUNK) {
        const size_t cnt = std::min((size_t)PERM_CHUNK, cols.size() - base);
        std::vector<Fr> mv(n, FR_ONE);
        for (size_t t = 0; t < cnt; ++t) {
            const PermColumn &pc = cols[base + t];
            const Fr *v = pc.type == 0 ? c.advice[pc.index] : c.fixed[pc.index];
            for (size_t i = 0; i < n; ++i)
                mv[i] = fr_mul(mv[i], fr_add(fr_add(fr_mul(beta, sigma[base + t][i]), gamma), v[i]));
        }
        fr_batch_invert(mv);
        for (size_t t = 0; t < cnt; ++t) {
            const PermColumn &pc = cols[base + t];
            const Fr *v = pc.type == 0 ? c.advice[pc.index] : c.fixed[pc.index];
            Fr dw = deltaomega0;
            for (size_t i = 0; i < n; ++i) {
                mv[i] = fr_mul(mv[i], fr_add(fr_add(fr_mul(dw, beta), gamma), v[i]));
                dw = fr_mul(dw, omega);
            }
            deltaomega0 = fr_mul(deltaomega0, delta);
        }
        std::vector<Fr> z(n, FR_ZERO);
        z[0] = last_z;
        for (size_t row = 1; row < n - 5; ++row) z[row] = fr_mul(z[row - 1], mv[row - 1]);
        last_z = z[n - 6];
        Z.push_back(std::move(z));
    }
}

// ---------------------------------------------------------------------------------------------
// [UPSTREAM] MockProver::verify restated for this constraint system.
// Returns 0 when satisfied; otherwise a failure class (100 gate, 200 lookup, 300 copy,
// 400 assignment in unusable rows) and fills `where` with (index,row).
// ---------------------------------------------------------------------------------------------
static int mock_verify(const Columns &c, const Copy *copies, size_t n_copies, uint64_t where[2]) {
    const Shape &sh = c.sh;
    size_t n = sh.n(), usable = sh.usable();
    where[0] = where[1] = 0;
    // (iv) nothing may be assigned outside the usable rows
    for (size_t row = usable; row < n; ++row) {
        for (size_t a = 0; a < c.advice.size(); ++a)
            if (!fr_is_zero(c.advice[a][row])) { where[0] = a; where[1] = row; return 400; }
        for (size_t f = 0; f < c.fixed.size(); ++f)
            if (!fr_is_zero(c.fixed[f][row])) { where[0] = f; where[1] = row; return 401; }
        for (size_t q = 0; q < c.selectors.size(); ++q)
            if (c.selectors[q][row]) { where[0] = q; where[1] = row; return 402; }
    }
    // (i) gate "Equality RC": q_eq_rcon * (words - round_constants) = 0  (key_schedule.rs:59-64)
    for (size_t row = 0; row < usable; ++row) {
        if (!c.selectors[sh.sel_rcon()][row]) continue;
        if (c.advice[sh.words_column()][row] != c.fixed[sh.rc_column()][row]) {
            where[0] = 0; where[1] = row; return 100;
        }
    }
    // (ii) lookups: every usable row's input tuple must be a row of the table tuple set
    for (int li = 0; li < sh.n_lookups(); ++li) {
        LookupDef d = lookup_def(sh, li);
        std::vector<std::array<Key, 4>> table(usable);
        for (size_t row = 0; row < usable; ++row)
            for (int j = 0; j < 4; ++j)
                table[row][j] = j < d.m ? key_of(c.fixed[j][row]) : Key{{0, 0, 0, 0}};
        std::sort(table.begin(), table.end());
        table.erase(std::unique(table.begin(), table.end()), table.end());
        Key tagk = key_of(fr_from_u64(d.tag));
        std::array<Key, 4> zero_t{};
        bool zero_ok = std::binary_search(table.begin(), table.end(), zero_t);
        for (size_t row = 0; row < usable; ++row) {
            if (!c.selectors[d.selector][row]) {
                if (!zero_ok) { where[0] = li; where[1] = row; return 200; }
                continue;
            }
            std::array<Key, 4> t{};
            t[0] = tagk;
            for (int j = 0; j + 1 < d.m; ++j) t[1 + j] = key_of(c.advice[d.adv[j]][row]);
            if (!std::binary_search(table.begin(), table.end(), t)) {
                where[0] = li; where[1] = row; return 200;
            }
        }
    }
    // (iii) every copy constraint holds
    for (size_t i = 0; i < n_copies; ++i) {
        const Copy &cp = copies[i];
        if (cp.col_a >= c.advice.size() || cp.col_b >= c.advice.size() || cp.row_a >= usable ||
            cp.row_b >= usable) { where[0] = i; where[1] = cp.row_a; return 301; }
        if (c.advice[cp.col_a][cp.row_a] != c.advice[cp.col_b][cp.row_b]) {
            where[0] = i; where[1] = cp.row_a; return 300;
        }
    }
    return 0;
}

struct Circuit {
    Synth s;
    AesConfig cfg;
    std::vector<uint8_t> ciphertexts;
    uint8_t round_keys[176];
    Circuit(const Shape &sh, bool fips) : s(sh, fips) { cfg.s = &s; }
};

// The `synthesize` bodies: aes128.rs:389-402 / main.rs:42-54 / benches/aes128.rs:44-56 (AES) and
// key_schedule.rs:296-315 / benches/key_schedule.rs:72-84 (key schedule only).
static void synthesize(Circuit &c, const uint8_t key[16], const uint8_t *pts, uint64_t n_blocks) {
    load_enc_full_table(c.s);
    if (c.s.sh.kind == KIND_KEYSCHED) {
        AesConfig &cfg = c.cfg;
        cfg.keys = schedule_keys(c.s, cfg.cols(0), key);
        cfg.have_keys = true;
    } else {
        c.cfg.schedule_key(key);
    }
    for (int r = 0; r < 11; ++r)
        for (int i = 0; i < 16; ++i) c.round_keys[16 * r + i] = first_byte(c.cfg.keys[r][i].value);
    c.ciphertexts.resize(16 * n_blocks);
    for (uint64_t b = 0; b < n_blocks; ++b) {
        Cells ct = c.cfg.encrypt(pts + 16 * b);
        for (int i = 0; i < 16; ++i) c.ciphertexts[16 * b + i] = first_byte(ct[i].value);
    }
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests/, bench.py cpu legs)
// ---------------------------------------------------------------------------------------------
using namespace orc;

static Fr fr_load(const uint64_t *p) { return Fr{{p[0], p[1], p[2], p[3]}}; }

extern "C" {

int orc_shape(int kind, int k, int N, int32_t out[5]) {
    if (k < 1 || k > 28 || N < 1 || (kind != KIND_AES && kind != KIND_KEYSCHED)) return ORC_ERR_ARG;
    Shape sh{kind, k, N};
    out[0] = sh.n_advice();
    out[1] = sh.n_fixed();
    out[2] = sh.n_selectors();
    out[3] = sh.n_lookups();
    out[4] = sh.n_sets();
    return ORC_OK;
}

void *orc_new(int kind, int k, int N, int fips_sbox) {
    if (k < 4 || k > 26 || N < 1 || (kind != KIND_AES && kind != KIND_KEYSCHED)) return nullptr;
    try {
        return new Circuit(Shape{kind, k, N}, fips_sbox != 0);
    } catch (...) { return nullptr; }
}
void orc_free(void *h) { delete (Circuit *)h; }

void orc_set_record_copies(void *h, int on) { ((Circuit *)h)->s.record_copies = on != 0; }

int orc_synthesize(void *h, const uint8_t *key, const uint8_t *pts, uint64_t n_blocks) {
    Circuit *c = (Circuit *)h;
    try {
        synthesize(*c, key, pts, n_blocks);
    } catch (const OrcError &e) { return e.code; }
    return ORC_OK;
}

uint64_t orc_num_regions(void *h) { return ((Circuit *)h)->s.n_regions; }
uint64_t orc_num_copies(void *h) { return ((Circuit *)h)->s.copies.size(); }
int orc_get_copies(void *h, uint64_t *out) {  // 4 u64 per copy: col_a,row_a,col_b,row_b
    Circuit *c = (Circuit *)h;
    for (size_t i = 0; i < c->s.copies.size(); ++i) {
        out[4 * i] = c->s.copies[i].col_a;
        out[4 * i + 1] = c->s.copies[i].row_a;
        out[4 * i + 2] = c->s.copies[i].col_b;
        out[4 * i + 3] = c->s.copies[i].row_b;
    }
    return ORC_OK;
}
int orc_get_advice(void *h, int col, uint64_t *out) {
    Circuit *c = (Circuit *)h;
    if (col < 0 || col >= (int)c->s.advice.size()) return ORC_ERR_ARG;
    memcpy(out, c->s.advice[col].data(), c->s.n * 32);
    return ORC_OK;
}
int orc_get_fixed(void *h, int col, uint64_t *out) {
    Circuit *c = (Circuit *)h;
    if (col < 0 || col >= (int)c->s.fixed.size()) return ORC_ERR_ARG;
    memcpy(out, c->s.fixed[col].data(), c->s.n * 32);
    return ORC_OK;
}
int orc_get_selector(void *h, int sel, uint8_t *out) {
    Circuit *c = (Circuit *)h;
    if (sel < 0 || sel >= (int)c->s.selectors.size()) return ORC_ERR_ARG;
    memcpy(out, c->s.selectors[sel].data(), c->s.n);
    return ORC_OK;
}
int orc_get_ciphertexts(void *h, uint8_t *out) {
    Circuit *c = (Circuit *)h;
    memcpy(out, c->ciphertexts.data(), c->ciphertexts.size());
    return ORC_OK;
}
int orc_get_round_keys(void *h, uint8_t *out) {
    memcpy(out, ((Circuit *)h)->round_keys, 176);
    return ORC_OK;
}

// A, S, A', S', Z for lookup `idx` of the circuit's own columns; any output may be NULL.
int orc_lookup(void *h, int idx, const uint64_t *theta, const uint64_t *beta,
               const uint64_t *gamma, uint64_t *A, uint64_t *S, uint64_t *Ap, uint64_t *Sp,
               uint64_t *Z) {
    Circuit *c = (Circuit *)h;
    if (idx < 0 || idx >= c->s.sh.n_lookups()) return ORC_ERR_ARG;
    Columns cols = view_of(c->s);
    LookupDef d = lookup_def(c->s.sh, idx);
    std::vector<Fr> a, s, ap, sp, z;
    compress_lookup(cols, d, fr_load(theta), a, s);
    int rc = permute_expression_pair(c->s.sh, a, s, ap, sp);
    if (rc) return rc;
    size_t bytes = c->s.n * 32;
    if (Z) {
        commit_product(c->s.sh, a, s, ap, sp, fr_load(beta), fr_load(gamma), z);
        memcpy(Z, z.data(), bytes);
    }
    if (A) memcpy(A, a.data(), bytes);
    if (S) memcpy(S, s.data(), bytes);
    if (Ap) memcpy(Ap, ap.data(), bytes);
    if (Sp) memcpy(Sp, sp.data(), bytes);
    return ORC_OK;
}

// Restated MockProver on the circuit's own columns.
int orc_mock_verify(void *h, uint64_t where[2]) {
    Circuit *c = (Circuit *)h;
    Columns cols = view_of(c->s);
    return mock_verify(cols, c->s.copies.data(), c->s.copies.size(), where);
}

// Restated MockProver on caller-supplied columns (e.g. copied back from the GPU):
// advice  = n_advice   columns of n Fr, contiguous column-major
// fixed   = 5          columns of n Fr (4 table columns, round_constants)
// selectors = n_selectors columns of n bytes (0/1)
// copies  = 4 u64 per copy constraint
int orc_mock_verify_ext(int kind, int k, int N, const uint64_t *advice, const uint64_t *fixed,
                        const uint8_t *selectors, const uint64_t *copies, uint64_t n_copies,
                        uint64_t where[2]) {
    Shape sh{kind, k, N};
    size_t n = sh.n();
    Columns cols;
    cols.sh = sh;
    for (int a = 0; a < sh.n_advice(); ++a) cols.advice.push_back((const Fr *)advice + (size_t)a * n);
    for (int f = 0; f < sh.n_fixed(); ++f) cols.fixed.push_back((const Fr *)fixed + (size_t)f * n);
    for (int q = 0; q < sh.n_selectors(); ++q) cols.selectors.push_back(selectors + (size_t)q * n);
    std::vector<Copy> cp(n_copies);
    for (size_t i = 0; i < n_copies; ++i)
        cp[i] = Copy{(uint32_t)copies[4 * i], copies[4 * i + 1], (uint32_t)copies[4 * i + 2],
                     copies[4 * i + 3]};
    return mock_verify(cols, cp.data(), cp.size(), where);
}

// permutation argument of the circuit's own columns: ncols/nsets via orc_perm_shape;
// sigma [ncols][n][4], mapping [ncols][n][2] (u32 column position, row), Z [nsets][n][4]; any
// output may be NULL.
int orc_perm_shape(int kind, int k, int N, int32_t out[3], int32_t *col_types, int32_t *col_indices) {
    Shape sh{kind, k, N};
    std::vector<PermColumn> cols = perm_columns(sh);
    out[0] = (int32_t)cols.size();
    out[1] = PERM_CHUNK;
    out[2] = (int32_t)((cols.size() + PERM_CHUNK - 1) / PERM_CHUNK);
    for (size_t i = 0; i < cols.size(); ++i) {
        if (col_types) col_types[i] = cols[i].type;
        if (col_indices) col_indices[i] = cols[i].index;
    }
    return ORC_OK;
}
int orc_permutation(void *h, const uint64_t *beta, const uint64_t *gamma, uint64_t *sigma_out,
                    uint32_t *mapping_out, uint64_t *Z_out) {
    Circuit *c = (Circuit *)h;
    std::vector<PermColumn> cols = perm_columns(c->s.sh);
    std::vector<std::vector<Fr>> sigma, Z;
    std::vector<uint32_t> map;
    build_sigma(c->s, cols, sigma, mapping_out ? &map : nullptr);
    const size_t n = c->s.n;
    if (sigma_out)
        for (size_t i = 0; i < cols.size(); ++i) memcpy(sigma_out + 4 * i * n, sigma[i].data(), n * 32);
    if (mapping_out) memcpy(mapping_out, map.data(), map.size() * 4);
    if (Z_out) {
        Columns cv = view_of(c->s);
        permutation_product(cv, cols, sigma, fr_load(beta), fr_load(gamma), Z);
        for (size_t sidx = 0; sidx < Z.size(); ++sidx)
            memcpy(Z_out + 4 * sidx * n, Z[sidx].data(), n * 32);
    }
    return ORC_OK;
}
void orc_domain_constants(int k, uint64_t omega[4], uint64_t delta[4]) {
    Fr w = domain_omega(k), d = fr_of_canon(DELTA_CANON);
    memcpy(omega, w.l, 32);
    memcpy(delta, d.l, 32);
}

// ---------------------------------------------------------------------------------------------
// [UPSTREAM] evaluation domain: halo2_proofs v0.3.0 poly/domain.rs (EvaluationDomain::new,
// lagrange_to_coeff, coeff_to_extended, extended_to_coeff, distribute_powers_zeta) and
// arithmetic.rs (best_fft, serial form).  SURVEY 8(f) items 3/4: the transforms keygen_pk applies
// to the fixed/sigma columns and create_proof to advice, A', S', Z (call sites main.rs:78-101).
// halo2curves 0.6.1 bn256::Fr::ZETA = 7^((r-1)/3) (re-derived in tests/test_oracle_golden.py).
// ---------------------------------------------------------------------------------------------
static const uint64_t ZETA_CANON[4] = {0x8b17ea66b99c90ddull, 0x5bfc41088d8daaa7ull,
                                       0xb3c4d79d41a91758ull, 0x0ull};
struct Domain {
    int k, extended_k;
    uint64_t n, quotient_poly_degree;
    Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv;
    Fr ifft_divisor, extended_ifft_divisor;
};
// EvaluationDomain::new(j, k): quotient_poly_degree = j - 1; extended_k = smallest with
// 2^extended_k >= n * (j - 1); extended_omega = ROOT_OF_UNITY squared (S - extended_k) times;
// omega = extended_omega squared (extended_k - k) times
static Domain domain_new(int j, int k) {
    Domain d;
    d.k = k;
    d.n = 1ull << k;
    d.quotient_poly_degree = (uint64_t)(j - 1);
    d.extended_k = k;
    while ((1ull << d.extended_k) < d.n * d.quotient_poly_degree) d.extended_k += 1;
    Fr w = fr_of_canon(ROOT_OF_UNITY_CANON);
    for (int i = d.extended_k; i < 28; ++i) w = fr_mul(w, w);
    d.extended_omega = w;
    for (int i = k; i < d.extended_k; ++i) w = fr_mul(w, w);
    d.omega = w;
    d.omega_inv = fr_inv(d.omega);
    d.extended_omega_inv = fr_inv(d.extended_omega);
    d.g_coset = fr_of_canon(ZETA_CANON);
    d.g_coset_inv = fr_mul(d.g_coset, d.g_coset);
    d.ifft_divisor = fr_inv(fr_from_u64(d.n));
    d.extended_ifft_divisor = fr_inv(fr_from_u64(1ull << d.extended_k));
    return d;
}
static uint64_t bitreverse(uint64_t v, int bits) {
    uint64_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}
// arithmetic.rs best_fft: bit-reversal swap, then log_n butterfly layers; layer m pairs
// (k+j, k+j+m) with twiddle omega^(j n / 2m).  halo2 walks each block with a running twiddle (or
// recurses per thread); here the twiddles come from one table of omega^t, t < n/2, and the
// butterflies of a layer run under OpenMP — field arithmetic is exact, so the values are the same.
static void serial_fft(Fr *a, const Fr &omega, int log_n) {
    const uint64_t n = 1ull << log_n;
    for (uint64_t k = 0; k < n; ++k) {
        const uint64_t rk = bitreverse(k, log_n);
        if (k < rk) std::swap(a[rk], a[k]);
    }
    std::vector<Fr> tw(n > 1 ? n / 2 : 1);
    tw[0] = FR_ONE;
    for (uint64_t t = 1; t < n / 2; ++t) tw[t] = fr_mul(tw[t - 1], omega);
    for (int layer = 0; layer < log_n; ++layer) {
        const uint64_t m = 1ull << layer, step = n >> (layer + 1);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (int64_t b = 0; b < (int64_t)(n / 2); ++b) {
            const uint64_t j = (uint64_t)b & (m - 1), k = ((uint64_t)b >> layer) << (layer + 1);
            const Fr t = fr_mul(a[k + j + m], tw[j * step]);
            a[k + j + m] = fr_sub(a[k + j], t);
            a[k + j] = fr_add(a[k + j], t);
        }
    }
}
// EvaluationDomain::ifft
static void domain_ifft(Fr *a, const Fr &omega_inv, int log_n, const Fr &divisor) {
    serial_fft(a, omega_inv, log_n);
    const uint64_t n = 1ull << log_n;
    for (uint64_t i = 0; i < n; ++i) a[i] = fr_mul(a[i], divisor);
}
// distribute_powers_zeta: a[i] *= [1, c0, c1][i % 3] with (c0, c1) = (zeta, zeta^2) into the
// coset and (zeta^2, zeta) out of it
static void distribute_powers_zeta(const Domain &d, Fr *a, uint64_t len, bool into_coset) {
    const Fr c0 = into_coset ? d.g_coset : d.g_coset_inv, c1 = into_coset ? d.g_coset_inv : d.g_coset;
    for (uint64_t i = 0; i < len; ++i) {
        const uint64_t r = i % 3;
        if (r == 1) a[i] = fr_mul(a[i], c0);
        if (r == 2) a[i] = fr_mul(a[i], c1);
    }
}

// op: 0 best_fft(a, omega, k)            (coefficients -> evaluations over the 2^k domain)
//     1 lagrange_to_coeff                  (in place, 2^k elements)
//     2 coeff_to_extended                  (in: 2^k coefficients, out: 2^extended_k evaluations
//                                           over the coset zeta * <extended_omega>)
//     3 extended_to_coeff                  (in: 2^extended_k, out: n * (j-1) coefficients)
// returns the number of output elements (or a negative error)
int64_t orc_domain_transform(int j, int k, int op, const uint64_t *in, uint64_t *out) {
    if (k < 1 || k > 26 || j < 2) return -ORC_ERR_ARG;
    const Domain d = domain_new(j, k);
    const uint64_t en = 1ull << d.extended_k;
    if (op == 0 || op == 1) {
        std::vector<Fr> a(d.n);
        memcpy(a.data(), in, d.n * 32);
        if (op == 0) serial_fft(a.data(), d.omega, k);
        else domain_ifft(a.data(), d.omega_inv, k, d.ifft_divisor);
        memcpy(out, a.data(), d.n * 32);
        return (int64_t)d.n;
    }
    if (op == 2) {
        std::vector<Fr> a(en, fr_from_u64(0));
        memcpy(a.data(), in, d.n * 32);
        distribute_powers_zeta(d, a.data(), d.n, true);
        serial_fft(a.data(), d.extended_omega, d.extended_k);
        memcpy(out, a.data(), en * 32);
        return (int64_t)en;
    }
    if (op == 3) {
        std::vector<Fr> a(en);
        memcpy(a.data(), in, en * 32);
        domain_ifft(a.data(), d.extended_omega_inv, d.extended_k, d.extended_ifft_divisor);
        distribute_powers_zeta(d, a.data(), en, false);
        const uint64_t keep = d.n * d.quotient_poly_degree;
        memcpy(out, a.data(), keep * 32);
        return (int64_t)keep;
    }
    return -ORC_ERR_ARG;
}
// omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor,
// extended_ifft_divisor (8 x 4 limbs); returns extended_k
int orc_domain_query(int j, int k, uint64_t *out32) {
    const Domain d = domain_new(j, k);
    const Fr *f[8] = {&d.omega, &d.omega_inv, &d.extended_omega, &d.extended_omega_inv,
                      &d.g_coset, &d.g_coset_inv, &d.ifft_divisor, &d.extended_ifft_divisor};
    for (int i = 0; i < 8; ++i) memcpy(out32 + 4 * i, f[i]->l, 32);
    return d.extended_k;
}

// ---------------------------------------------------------------------------------------------
// [UPSTREAM] plonk/evaluation.rs Evaluator::evaluate_h, the lookup part, and
// vanishing/prover.rs + poly/domain.rs divide_by_vanishing_poly (halo2_proofs v0.3.0).
// SURVEY 8(f) item 4.  For every lookup, in meta.lookup order, with `values` the running
// extended-domain accumulator and idx an extended row (r_next / r_prev = idx +- 2^(ek-k)):
//   values = values*y + (1 - z) * l0
//   values = values*y + (z^2 - z) * l_last
//   values = values*y + (z[r_next] (a'+beta)(s'+gamma) - z (A+beta)(S+gamma)) * l_active_row
//   values = values*y + (a' - s') * l0
//   values = values*y + (a' - s')(a' - a'[r_prev]) * l_active_row
// A, S = the theta-compressions of the input / table expressions evaluated on the extended
// coset (q*tag, q*x, ... with q, x the cosets of the selector and advice polynomials).
// l0, l_last, l_blind as keygen.rs builds them: one-hot Lagrange columns at row 0, row n-6 and
// rows n-5.., taken to the extended coset; l_active_row = 1 - (l_last + l_blind).
// ---------------------------------------------------------------------------------------------
static const int BLINDING_FACTORS = 5;  // cs.blinding_factors() of both circuits (usable = n - 6)

static std::vector<Fr> to_extended(const Domain &d, const Fr *lagrange) {
    std::vector<Fr> a(1ull << d.extended_k, FR_ZERO);
    memcpy(a.data(), lagrange, d.n * 32);
    domain_ifft(a.data(), d.omega_inv, d.k, d.ifft_divisor);
    distribute_powers_zeta(d, a.data(), d.n, true);
    serial_fft(a.data(), d.extended_omega, d.extended_k);
    return a;
}

struct ExtLagrange {
    std::vector<Fr> l0, l_last, l_active_row;
};
static ExtLagrange ext_lagrange(const Domain &d) {
    std::vector<Fr> v(d.n, FR_ZERO);
    ExtLagrange e;
    v[0] = FR_ONE;
    e.l0 = to_extended(d, v.data());
    v[0] = FR_ZERO;
    v[d.n - BLINDING_FACTORS - 1] = FR_ONE;
    e.l_last = to_extended(d, v.data());
    v[d.n - BLINDING_FACTORS - 1] = FR_ZERO;
    for (int i = 0; i < BLINDING_FACTORS; ++i) v[d.n - 1 - i] = FR_ONE;
    std::vector<Fr> l_blind = to_extended(d, v.data());
    const uint64_t en = 1ull << d.extended_k;
    e.l_active_row.resize(en);
    for (uint64_t i = 0; i < en; ++i)
        e.l_active_row[i] = fr_sub(FR_ONE, fr_add(e.l_last[i], l_blind[i]));
    return e;
}

// values (2^extended_k elements, in/out) accumulates lookups [first, first+count) of the
// circuit's own columns; blinding rows of every column are zero, as everywhere in this oracle.
int orc_lookup_quotient(void *h, int first, int count, const uint64_t *theta_,
                        const uint64_t *beta_, const uint64_t *gamma_, const uint64_t *y_,
                        uint64_t *values) {
    Circuit *c = (Circuit *)h;
    const Shape &sh = c->s.sh;
    if (first < 0 || count < 0 || first + count > sh.n_lookups()) return ORC_ERR_ARG;
    const Domain d = domain_new(5, sh.k);
    const uint64_t en = 1ull << d.extended_k, n = d.n;
    const int64_t rot = (int64_t)1 << (d.extended_k - d.k);
    const Fr theta = fr_load(theta_), beta = fr_load(beta_), gamma = fr_load(gamma_), y = fr_load(y_);
    const ExtLagrange L = ext_lagrange(d);
    Columns cols = view_of(c->s);
    std::vector<Fr> acc(en);
    memcpy(acc.data(), values, en * 32);
    std::vector<std::vector<Fr>> table_ext;
    for (int t = 0; t < 4; ++t) table_ext.push_back(to_extended(d, cols.fixed[t]));
    std::vector<Fr> tmp(n);
    for (int li = first; li < first + count; ++li) {
        const LookupDef def = lookup_def(sh, li);
        std::vector<Fr> a, s, ap, sp, z;
        compress_lookup(cols, def, theta, a, s);
        int rc = permute_expression_pair(sh, a, s, ap, sp);
        if (rc) return rc;
        commit_product(sh, a, s, ap, sp, beta, gamma, z);
        const std::vector<Fr> ap_e = to_extended(d, ap.data()), sp_e = to_extended(d, sp.data()),
                              z_e = to_extended(d, z.data());
        for (uint64_t i = 0; i < n; ++i) tmp[i] = cols.selectors[def.selector][i] ? FR_ONE : FR_ZERO;
        const std::vector<Fr> q_e = to_extended(d, tmp.data());
        std::vector<std::vector<Fr>> adv_e;
        for (int jx = 0; jx + 1 < def.m; ++jx) adv_e.push_back(to_extended(d, cols.advice[def.adv[jx]]));
        const Fr tag = fr_from_u64(def.tag);
        for (uint64_t idx = 0; idx < en; ++idx) {
            // the lookup's expressions on the extended coset, folded with theta
            Fr in = fr_mul(q_e[idx], tag), tb = table_ext[0][idx];
            for (int jx = 0; jx + 1 < def.m; ++jx) {
                in = fr_add(fr_mul(in, theta), fr_mul(q_e[idx], adv_e[jx][idx]));
                tb = fr_add(fr_mul(tb, theta), table_ext[1 + jx][idx]);
            }
            const Fr table_value = fr_mul(fr_add(in, beta), fr_add(tb, gamma));
            const uint64_t r_next = (uint64_t)(((int64_t)idx + rot) % (int64_t)en);
            const uint64_t r_prev = (uint64_t)(((int64_t)idx - rot + (int64_t)en) % (int64_t)en);
            const Fr a_minus_s = fr_sub(ap_e[idx], sp_e[idx]);
            Fr v = acc[idx];
            v = fr_add(fr_mul(v, y), fr_mul(fr_sub(FR_ONE, z_e[idx]), L.l0[idx]));
            v = fr_add(fr_mul(v, y), fr_mul(fr_sub(fr_mul(z_e[idx], z_e[idx]), z_e[idx]), L.l_last[idx]));
            const Fr lhs = fr_mul(fr_mul(z_e[r_next], fr_add(ap_e[idx], beta)), fr_add(sp_e[idx], gamma));
            v = fr_add(fr_mul(v, y), fr_mul(fr_sub(lhs, fr_mul(z_e[idx], table_value)), L.l_active_row[idx]));
            v = fr_add(fr_mul(v, y), fr_mul(a_minus_s, L.l0[idx]));
            v = fr_add(fr_mul(v, y), fr_mul(fr_mul(a_minus_s, fr_sub(ap_e[idx], ap_e[r_prev])), L.l_active_row[idx]));
            acc[idx] = v;
        }
    }
    memcpy(values, acc.data(), en * 32);
    return ORC_OK;
}

// [UPSTREAM] plonk/evaluation.rs evaluate_h, what comes BEFORE the lookups in halo2's order:
//   custom gates  values = values*y + q_eq_rcon (words - round_constants)   (key_schedule.rs:59-64;
//                 the simple selector is a 0/1 fixed column after compress_selectors)
//   permutation   values = values*y + (1 - z_0) l0
//                 values = values*y + (z_last^2 - z_last) l_last
//                 values = values*y + (z_i - z_{i-1}[r_last]) l0                    for sets i >= 1
//                 values = values*y + (z_i[r_next] prod_j (v_j + beta sigma_j + gamma)
//                                      - z_i prod_j (v_j + delta^p beta X + gamma)) l_active_row
//   with X = zeta * extended_omega^idx, r_last = idx - 6 * 2^(ek-k), p running over ALL columns.
int orc_gate_perm_quotient(void *h, const uint64_t *beta_, const uint64_t *gamma_,
                           const uint64_t *y_, uint64_t *values) {
    Circuit *c = (Circuit *)h;
    const Shape &sh = c->s.sh;
    const Domain d = domain_new(5, sh.k);
    const uint64_t en = 1ull << d.extended_k, n = d.n;
    const int64_t rot = (int64_t)1 << (d.extended_k - d.k);
    const Fr beta = fr_load(beta_), gamma = fr_load(gamma_), y = fr_load(y_);
    const Fr delta = fr_of_canon(DELTA_CANON);
    const ExtLagrange L = ext_lagrange(d);
    Columns cols = view_of(c->s);
    std::vector<Fr> acc(en);
    memcpy(acc.data(), values, en * 32);
    // gate
    {
        std::vector<Fr> tmp(n);
        for (uint64_t i = 0; i < n; ++i) tmp[i] = cols.selectors[sh.sel_rcon()][i] ? FR_ONE : FR_ZERO;
        const std::vector<Fr> q_e = to_extended(d, tmp.data()),
                              w_e = to_extended(d, cols.advice[sh.words_column()]),
                              rc_e = to_extended(d, cols.fixed[sh.rc_column()]);
        for (uint64_t idx = 0; idx < en; ++idx)
            acc[idx] = fr_add(fr_mul(acc[idx], y), fr_mul(q_e[idx], fr_sub(w_e[idx], rc_e[idx])));
    }
    // permutation
    std::vector<PermColumn> pcols = perm_columns(sh);
    std::vector<std::vector<Fr>> sigma, Z;
    build_sigma(c->s, pcols, sigma, nullptr);
    permutation_product(cols, pcols, sigma, beta, gamma, Z);
    std::vector<std::vector<Fr>> z_e, sig_e, val_e;
    for (auto &z : Z) z_e.push_back(to_extended(d, z.data()));
    for (size_t p = 0; p < pcols.size(); ++p) {
        sig_e.push_back(to_extended(d, sigma[p].data()));
        const Fr *v = pcols[p].type == 0 ? cols.advice[pcols[p].index] : cols.fixed[pcols[p].index];
        val_e.push_back(to_extended(d, v));
    }
    const size_t n_sets = z_e.size();
    Fr beta_term = FR_ONE;  // extended_omega^idx
    for (uint64_t idx = 0; idx < en; ++idx) {
        const uint64_t r_next = (uint64_t)(((int64_t)idx + rot) % (int64_t)en);
        const uint64_t r_last =
            (uint64_t)((((int64_t)idx - (BLINDING_FACTORS + 1) * rot) % (int64_t)en + (int64_t)en) % (int64_t)en);
        Fr v = acc[idx];
        v = fr_add(fr_mul(v, y), fr_mul(fr_sub(FR_ONE, z_e[0][idx]), L.l0[idx]));
        const Fr zl = z_e[n_sets - 1][idx];
        v = fr_add(fr_mul(v, y), fr_mul(fr_sub(fr_mul(zl, zl), zl), L.l_last[idx]));
        for (size_t si = 1; si < n_sets; ++si)
            v = fr_add(fr_mul(v, y), fr_mul(fr_sub(z_e[si][idx], z_e[si - 1][r_last]), L.l0[idx]));
        Fr current_delta = fr_mul(fr_mul(beta, d.g_coset), beta_term);  // delta_start * beta_term
        for (size_t si = 0; si < n_sets; ++si) {
            const size_t base = si * PERM_CHUNK, cnt = std::min((size_t)PERM_CHUNK, pcols.size() - base);
            Fr left = z_e[si][r_next], right = z_e[si][idx];
            for (size_t t = 0; t < cnt; ++t)
                left = fr_mul(left, fr_add(fr_add(val_e[base + t][idx], fr_mul(beta, sig_e[base + t][idx])), gamma));
            for (size_t t = 0; t < cnt; ++t) {
                right = fr_mul(right, fr_add(fr_add(val_e[base + t][idx], current_delta), gamma));
                current_delta = fr_mul(current_delta, delta);
            }
            v = fr_add(fr_mul(v, y), fr_mul(fr_sub(left, right), L.l_active_row[idx]));
        }
        acc[idx] = v;
        beta_term = fr_mul(beta_term, d.extended_omega);
    }
    memcpy(values, acc.data(), en * 32);
    return ORC_OK;
}

// EvaluationDomain::divide_by_vanishing_poly: a[i] *= t_evaluations[i % 2^(ek-k)], the inverses of
// (zeta * extended_omega^i)^n - 1  (in place, 2^extended_k elements)
int orc_divide_by_vanishing(int j, int k, uint64_t *a_) {
    const Domain d = domain_new(j, k);
    const uint64_t en = 1ull << d.extended_k, period = 1ull << (d.extended_k - d.k);
    std::vector<Fr> t(period);
    uint64_t e[4] = {d.n, 0, 0, 0};
    const Fr zn = fr_pow(d.g_coset, e), wn = fr_pow(d.extended_omega, e);
    Fr cur = zn;
    for (uint64_t i = 0; i < period; ++i) { t[i] = fr_inv(fr_sub(cur, FR_ONE)); cur = fr_mul(cur, wn); }
    Fr *a = (Fr *)a_;
    for (uint64_t i = 0; i < en; ++i) a[i] = fr_mul(a[i], t[i % period]);
    return ORC_OK;
}

// Horner evaluation of a coefficient vector at x (eval_polynomial of arithmetic.rs)
void orc_poly_eval(const uint64_t *coeffs, uint64_t len, const uint64_t *x_, uint64_t out[4]) {
    const Fr x = fr_load(x_);
    Fr acc = FR_ZERO;
    for (uint64_t i = len; i-- > 0;) acc = fr_add(fr_mul(acc, x), fr_load(coeffs + 4 * i));
    memcpy(out, acc.l, 32);
}

// ---------------------------------------------------------------------------------------------
// [UPSTREAM] commitments: halo2curves bn256 G1 + halo2_proofs arithmetic.rs best_multiexp
// (SURVEY 8(f) item 2).  Points travel as G1Affine (x, y: 4 + 4 u64 Montgomery limbs of Fq,
// identity = all zero).
// ---------------------------------------------------------------------------------------------
// out = sum_i scalars[i] * bases[i]
void orc_g1_msm(const uint64_t *scalars, const uint64_t *bases, uint64_t n, uint64_t out[8]) {
    const G1Affine r = g1_to_affine(g1_msm((const Fr *)scalars, (const G1Affine *)bases, n));
    memcpy(out, &r, 64);
}
// Test SRS with known discrete logarithms: bases[i] = (1 + i*d) * G, G = (1, 2): one addition of
// D = d*G per point, then one batch inversion back to affine.  d is an Fr in Montgomery form.
void orc_g1_test_bases(const uint64_t *d_, uint64_t n, uint64_t *bases) {
    uint64_t dc[4];
    fr_to_canonical(fr_load(d_), dc);
    const G1 g = g1_from_affine(g1_generator());
    const G1 dpt = g1_mul_canonical(g, dc);
    std::vector<G1> pts(n);
    G1 cur = g;
    for (uint64_t i = 0; i < n; ++i) { pts[i] = cur; cur = g1_add(cur, dpt); }
    // Montgomery-trick batch inversion of the z coordinates
    std::vector<Fq> pre(n);
    Fq acc = FQ_ONE;
    for (uint64_t i = 0; i < n; ++i) { pre[i] = acc; acc = fq_mul(acc, pts[i].z); }
    Fq inv = fq_inv(acc);
    G1Affine *out = (G1Affine *)bases;
    for (uint64_t i = n; i-- > 0;) {
        const Fq zi = fq_mul(inv, pre[i]);
        inv = fq_mul(inv, pts[i].z);
        const Fq zi2 = fq_mul(zi, zi);
        out[i] = G1Affine{fq_mul(pts[i].x, zi2), fq_mul(fq_mul(pts[i].y, zi2), zi)};
    }
}
// k * G for an Fr scalar in Montgomery form (affine out); and the curve equation
void orc_g1_mul_generator(const uint64_t *k_, uint64_t out[8]) {
    uint64_t kc[4];
    fr_to_canonical(fr_load(k_), kc);
    const G1Affine r = g1_to_affine(g1_mul_canonical(g1_from_affine(g1_generator()), kc));
    memcpy(out, &r, 64);
}
int orc_g1_on_curve(const uint64_t *p) { return g1a_on_curve(*(const G1Affine *)p) ? 1 : 0; }
void orc_g1_add_affine(const uint64_t *a, const uint64_t *b, uint64_t out[8]) {
    const G1Affine r = g1_to_affine(g1_add(g1_from_affine(*(const G1Affine *)a), g1_from_affine(*(const G1Affine *)b)));
    memcpy(out, &r, 64);
}

// ---- small Fr helpers so Python tests can do field arithmetic through the same code ----------
void orc_fr_from_u64(uint64_t v, uint64_t out[4]) { Fr a = fr_from_u64(v); memcpy(out, a.l, 32); }
void orc_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t out[4]) {
    Fr r = fr_mul(fr_load(a), fr_load(b)); memcpy(out, r.l, 32);
}
void orc_fr_add(const uint64_t *a, const uint64_t *b, uint64_t out[4]) {
    Fr r = fr_add(fr_load(a), fr_load(b)); memcpy(out, r.l, 32);
}
void orc_fr_inv(const uint64_t *a, uint64_t out[4]) { Fr r = fr_inv(fr_load(a)); memcpy(out, r.l, 32); }
void orc_fr_to_canonical(const uint64_t *a, uint64_t out[4]) { fr_to_canonical(fr_load(a), out); }
int orc_fr_from_canonical(const uint64_t *a, uint64_t out[4]) {
    Fr r;
    if (!fr_from_canonical(a, &r)) return ORC_ERR_FROM_BYTES;
    memcpy(out, r.l, 32);
    return ORC_OK;
}
// utils.rs:8-19 on two Fr values (golden: 5 ^ 12 == 9, utils.rs:41-47)
int orc_xor_bytes(const uint64_t *x, const uint64_t *y, uint64_t out[4]) {
    try {
        Fr r = xor_bytes(fr_load(x), fr_load(y));
        memcpy(out, r.l, 32);
    } catch (const OrcError &e) { return e.code; }
    return ORC_OK;
}
void orc_tables(int fips, uint8_t sbox[256], uint8_t mul2[256], uint8_t mul3[256]) {
    ByteTables t = make_tables(fips != 0);
    memcpy(sbox, t.sbox, 256);
    memcpy(mul2, t.mul2, 256);
    memcpy(mul3, t.mul3, 256);
}

// ---- CPU baseline: one whole circuit (synthesize + every lookup's A,S,A',S',Z), timed ----------
// The reference's synthesize is strictly sequential (one thread).  The lookup arguments are
// independent of one another, so `threads` > 1 runs them in parallel with OpenMP — an upper
// bound on what halo2's rayon `parallelize` could extract.  times[0] = synthesize seconds,
// times[1] = lookup seconds.  digest = XOR of all Z[n-6] limbs (forces the work; must be Fr(1)).
int orc_bench_circuit(int kind, int k, int N, const uint8_t *key, const uint8_t *pts,
                      uint64_t n_blocks, const uint64_t *theta, const uint64_t *beta,
                      const uint64_t *gamma, int threads, int n_lookups_to_run, double times[2],
                      uint64_t digest[4]) {
    Shape sh{kind, k, N};
    auto t0 = std::chrono::steady_clock::now();
    Circuit *c = new Circuit(sh, false);
    c->s.record_copies = false;  // the proving-time WitnessCollection ignores copy constraints
    int rc = ORC_OK;
    try {
        synthesize(*c, key, pts, n_blocks);
    } catch (const OrcError &e) { rc = e.code; }
    auto t1 = std::chrono::steady_clock::now();
    if (rc) { delete c; return rc; }
    Columns cols = view_of(c->s);
    int nl = n_lookups_to_run < 0 ? sh.n_lookups() : std::min(n_lookups_to_run, sh.n_lookups());
    std::vector<Fr> last(nl, FR_ZERO);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (int li = 0; li < nl; ++li) {
        LookupDef d = lookup_def(sh, li);
        std::vector<Fr> a, s, ap, sp, z;
        compress_lookup(cols, d, fr_load(theta), a, s);
        if (permute_expression_pair(sh, a, s, ap, sp)) {
#pragma omp atomic
            bad += 1;
            continue;
        }
        commit_product(sh, a, s, ap, sp, fr_load(beta), fr_load(gamma), z);
        last[li] = z[sh.n() - 6];
    }
    auto t2 = std::chrono::steady_clock::now();
    times[0] = std::chrono::duration<double>(t1 - t0).count();
    times[1] = std::chrono::duration<double>(t2 - t1).count();
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    for (auto &f : last)
        for (int i = 0; i < 4; ++i) digest[i] ^= f.l[i];
    delete c;
    return bad ? ORC_ERR_LOOKUP_INPUT_NOT_IN_TABLE : ORC_OK;
}


// ---- CPU baseline on all host threads: `n_circuits` independent circuits (the multi-circuit
// workload of BASELINE config 4), one circuit per OpenMP thread, each circuit run exactly like
// orc_bench_circuit with threads=1.  keys = n_circuits*16 bytes, pts = n_circuits*n_blocks*16.
// seconds_out = wall time of the whole batch.
int orc_bench_batch(int kind, int k, int N, int n_circuits, const uint8_t *keys,
                    const uint8_t *pts, uint64_t n_blocks, const uint64_t *theta,
                    const uint64_t *beta, const uint64_t *gamma, int threads, double *seconds_out,
                    uint64_t digest[4]) {
    auto t0 = std::chrono::steady_clock::now();
    int bad = 0;
    uint64_t dg[4] = {0, 0, 0, 0};
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (int c = 0; c < n_circuits; ++c) {
        double times[2];
        uint64_t d[4];
        int rc = orc_bench_circuit(kind, k, N, keys + 16 * (size_t)c,
                                   pts + 16 * n_blocks * (size_t)c, n_blocks, theta, beta, gamma,
                                   1, -1, times, d);
#pragma omp critical
        {
            if (rc) bad = rc;
            for (int i = 0; i < 4; ++i) dg[i] ^= d[i];
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    for (int i = 0; i < 4; ++i) digest[i] = dg[i];
    return bad;
}

}  // extern "C"
