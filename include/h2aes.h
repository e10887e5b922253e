/*
 * h2aes.h — C ABI of the B200 witness-generation backend for tkmct/halo2-aes.
 *
 * This is the drop-in boundary: everything a Rust `build.rs`/`extern "C"` block (or cgo, JNI,
 * ctypes ...) needs to replace the CPU assignment done by the reference's chips.  The reference
 * has no FFI of its own; each entry point below names the reference interface it stands in for
 * (paths relative to the reference repository).  See INTEGRATION.md for the Rust-side binding.
 *
 * Conventions
 *   - every function returns an int32 status (H2AES_OK == 0); nothing unwinds across the ABI;
 *   - field elements are bn256 `Fr` exactly as halo2curves 0.6.1 lays them out in memory:
 *     4 little-endian u64 limbs of x*R mod r (Montgomery form) — a column is `n` such elements,
 *     bit-compatible with `Vec<Fr>` / `Polynomial<Fr, LagrangeCoeff>`;
 *   - multi-column outputs are column-major `[column][row][4]`;
 *   - `_dev` functions take DEVICE pointers for bulk data and are asynchronous on the context's
 *     stream; `_host` functions take HOST pointers, do the copies themselves and return after
 *     the result is in host memory.  Small inputs (keys, challenges) are always host pointers;
 *   - buffers are caller-owned; the library never frees caller memory;
 *   - a context (and its circuits) may be used from one host thread at a time;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     H2AES_ERR_CUDA.
 */
#ifndef H2AES_H
#define H2AES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2AES_ABI_VERSION 1

/* status codes */
#define H2AES_OK 0
#define H2AES_ERR_ARG 1            /* null pointer, bad kind/k/N, out-of-range index             */
#define H2AES_ERR_CAPACITY 2       /* src/aes128.rs:160-162 panic!("AES calls too many...")      */
#define H2AES_ERR_ROWS 3           /* halo2 Error::NotEnoughRowsAvailable (k < 17, or a block     */
                                   /* would touch the last 6 rows)                                */
#define H2AES_ERR_NOT_SCHEDULED 4  /* src/aes128.rs:170 expect("Keys should be scheduled")       */
#define H2AES_ERR_LOOKUP 5         /* halo2 Error::ConstraintSystemFailure in                     */
                                   /* permute_expression_pair (input value not in table)          */
#define H2AES_ERR_CUDA 6
#define H2AES_ERR_NCCL 7
#define H2AES_ERR_NOMEM 8
#define H2AES_ERR_STATE 9          /* call order: lookups before witness, product before permuted */
#define H2AES_ERR_ZERO_DENOM 10    /* (A'+beta)(S'+gamma) == 0 for some row (prob. ~2^-230);      */
                                   /* halo2's batch_invert would silently leave a zero there      */

/* circuit kinds */
#define H2AES_KIND_AES 0      /* FixedAes128Config<K,N>: src/aes128.rs:27-141                     */
#define H2AES_KIND_KEYSCHED 1 /* key-schedule-only circuit: src/key_schedule.rs:242-320,          */
                              /* benches/key_schedule.rs:26-89 (N is ignored, one column set)     */

/* S-box data */
#define H2AES_SBOX_REFERENCE 0 /* src/constant.rs:1-14 as committed: S_BOX[255] == 23             */
#define H2AES_SBOX_FIPS197 1   /* FIPS-197 value 0x16 at index 255                                */

/* per-row operation codes (which complex selector is enabled on a row of a column set) */
#define H2AES_OP_NONE 0
#define H2AES_OP_RANGE 1 /* == table.rs Tag::U8     */
#define H2AES_OP_XOR 2   /* == table.rs Tag::Xor    */
#define H2AES_OP_SBOX 3  /* == table.rs Tag::Sbox   */
#define H2AES_OP_MUL2 4  /* == table.rs Tag::GfMul2 */
#define H2AES_OP_MUL3 5  /* == table.rs Tag::GfMul3 */

typedef struct h2aes_ctx h2aes_ctx;
typedef struct h2aes_circuit h2aes_circuit;

/* Result of h2aes_layout_query: what `configure` allocates and what the floor planner yields. */
typedef struct h2aes_layout {
    int32_t kind, k, n_sets;
    int32_t n_advice;    /* 3*n_sets + 1 : advices[i][0..3] = 3i..3i+2, words_column last       */
    int32_t n_fixed;     /* 5 : 4 table columns, then round_constants                            */
    int32_t n_selectors; /* 5*N (+1 q_eq_rcon) for AES; 3 (+1) for KEYSCHED; index 5i+op-1       */
    int32_t n_lookups;   /* == number of complex selectors, same order                           */
    int32_t reserved;
    uint64_t n_rows;             /* 2^k                                                           */
    uint64_t usable_rows;        /* 2^k - 6 (blinding_factors() + 1)                              */
    uint64_t table_rows;         /* 66561 assigned rows of load_enc_full_table                    */
    uint64_t key_schedule_rows;  /* 400 rows of advices[0][0] actually used by schedule_key       */
    uint64_t rows_per_block;     /* 1360 == constant.rs AES_ROWS                                  */
    uint64_t capacity_set0;      /* floor((2^k - 1760)/1360)  (aes128.rs:303-325)                 */
    uint64_t capacity_set;       /* floor(2^k/1360) for every later column set                    */
    uint64_t capacity_blocks;    /* capacity_set0 + (N-1)*capacity_set                            */
    uint64_t copies_key_schedule;/* 640                                                           */
    uint64_t copies_per_block;   /* 1952                                                          */
} h2aes_layout;

/* ---------------------------------------------------------------------------------------------
 * Host-only queries (no GPU needed).
 * ------------------------------------------------------------------------------------------- */
int32_t h2aes_abi_version(void);
const char *h2aes_status_string(int32_t status);

/* replaces: FixedAes128Config::configure (aes128.rs:46-141) + Aes128KeyScheduleConfig::configure
 * (key_schedule.rs:40-77) as far as column/selector/lookup counts and order go. */
int32_t h2aes_layout_query(int32_t kind, int32_t k, int32_t N, h2aes_layout *out);

/* replaces: aes_callable + get_advices (aes128.rs:303-356): the column set and the first row of
 * advices[set][0] used by the `block`-th call of `encrypt`.  H2AES_ERR_CAPACITY past the end. */
int32_t h2aes_block_position(int32_t k, int32_t N, uint64_t block, int32_t *set,
                             uint64_t *start_row);

/* which selector (H2AES_OP_*) the chips enable on `row` of column set `set` when `n_blocks`
 * blocks have been encrypted; negative status on error.  (q.enable calls in the chips, src/chips/) */
int32_t h2aes_row_op(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks, int32_t set,
                     uint64_t row);

/* rows on which the selector of lookup `lookup` is enabled after schedule_key + n_blocks encrypts
 * (the Selector::enable calls of the chips; input independent). */
int32_t h2aes_active_rows(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks, int32_t lookup,
                          uint64_t *out);

/* number of copy constraints recorded by schedule_key + n_blocks encrypts (640 + 1952*n). */
int32_t h2aes_copy_count(int32_t kind, uint64_t n_blocks, uint64_t *out);

/* copy constraints [first, first+count) as 4 u64 each: (advice col a, row a, advice col b,
 * row b); replaces the constrain_equal calls made by copy_advice in the chips. Host buffer. */
int32_t h2aes_copy_constraints_host(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks,
                                    uint64_t first, uint64_t count, uint64_t *out);

/* ---- permutation (copy-constraint) argument — SURVEY 8(f) item 1 ---------------------------
 * replaces (as far as the columns go): halo2_proofs v0.3.0 plonk/permutation/keygen.rs
 * (Assembly::copy / build_pk) and plonk/permutation/prover.rs (Argument::commit), call sites
 * keygen_vk/keygen_pk and create_proof (main.rs:78-101).
 * Columns of the argument, in enable_equality order (key_schedule.rs:52-57, aes128.rs:126-130):
 * advices[0][0..3], words_column, round_constants (fixed), then the remaining advice columns;
 * chunk_len = cs.degree() - 2 = 3 columns per grand product.  Arrays: n_columns entries. */
int32_t h2aes_perm_query(int32_t kind, int32_t k, int32_t N, int32_t *n_columns, int32_t *chunk_len,
                         int32_t *n_sets, int32_t *is_fixed, int32_t *column_index);
/* the cycle structure halo2's keygen builds from the copy constraints of schedule_key + n_blocks
 * encrypts: mapping[(p*n + row)*2 + {0,1}] = (position, row) that cell (p,row) maps to.  Host
 * out, n_columns * 2^k * 2 u32.  Input independent (shape + n_blocks only). */
int32_t h2aes_perm_mapping_host(int32_t kind, int32_t k, int32_t N, uint64_t n_blocks,
                                uint32_t *mapping_out);

/* ---------------------------------------------------------------------------------------------
 * Context: one per (process, GPU).  Uploads the S-box / GF tables and the 256-entry
 * byte -> Fr(Montgomery) table (replaces src/constant.rs + every `Fp::from(x as u64)`).
 * ------------------------------------------------------------------------------------------- */
int32_t h2aes_ctx_create(int32_t device, int32_t sbox_mode, h2aes_ctx **out);
int32_t h2aes_ctx_destroy(h2aes_ctx *ctx);
/* `stream` is a cudaStream_t; NULL = the legacy default stream. */
int32_t h2aes_ctx_set_stream(h2aes_ctx *ctx, void *stream);
int32_t h2aes_ctx_synchronize(h2aes_ctx *ctx);
const char *h2aes_ctx_last_error(h2aes_ctx *ctx);
/* number of this library's kernels launched through `ctx` so far. */
uint64_t h2aes_ctx_kernel_launches(h2aes_ctx *ctx);
/* per-kernel timing: with enable != 0 every kernel launched through `ctx` is bracketed by CUDA
 * events on the context's stream; h2aes_ctx_profile_read synchronizes and writes one line
 * "kernel_name launches total_ms" per kernel (first-launch order) into buf, then clears. */
int32_t h2aes_ctx_profile(h2aes_ctx *ctx, int32_t enable);
int32_t h2aes_ctx_profile_read(h2aes_ctx *ctx, char *buf, uint64_t cap);
/* pinned host memory helpers for the _host entry points (optional; any host memory works). */
int32_t h2aes_host_alloc(uint64_t bytes, void **out);
int32_t h2aes_host_free(void *p);

/* ---------------------------------------------------------------------------------------------
 * Circuit: one halo2 circuit instance of shape (kind, k, N).  Holds device workspace (the
 * compact per-row trace the lookup phases read, the per-lookup bin structures, round keys).
 * ------------------------------------------------------------------------------------------- */
int32_t h2aes_circuit_create(h2aes_ctx *ctx, int32_t kind, int32_t k, int32_t N,
                             h2aes_circuit **out);
int32_t h2aes_circuit_destroy(h2aes_circuit *c);

/* replaces: FixedAes128Config::schedule_key (aes128.rs:143-152 -> key_schedule.rs:80-224)
 * followed by n_blocks calls of FixedAes128Config::encrypt (aes128.rs:154-265, lcon :268-301).
 *   key            host, 16 bytes
 *   plaintexts     n_blocks*16 bytes (may be NULL when n_blocks == 0)
 *   advice         n_advice columns * 2^k rows * 4 u64: every cell the chips assign, zero
 *                  elsewhere (rows >= 2^k-6 are left zero: halo2 fills them with blinding)
 *   ciphertexts    n_blocks*16 bytes = the values of the 16 cells `encrypt` returns; nullable
 * For H2AES_KIND_KEYSCHED n_blocks must be 0. */
int32_t h2aes_witness_dev(h2aes_circuit *c, const uint8_t *key, const uint8_t *plaintexts_dev,
                          uint64_t n_blocks, uint64_t *advice_dev, uint8_t *ciphertexts_dev);
int32_t h2aes_witness_host(h2aes_circuit *c, const uint8_t *key, const uint8_t *plaintexts,
                           uint64_t n_blocks, uint64_t *advice, uint8_t *ciphertexts);

/* replaces: repeated FixedAes128Config::encrypt calls (aes128.rs:154, `&mut self`: one block per
 * call, position kept in `current`/`count`, aes128.rs:303-325).  Appends n more blocks after the
 * ones this circuit already holds, under the key of the last h2aes_witness_* call (which may have
 * had n_blocks == 0: that is schedule_key alone) — so a shim behind `encrypt` can flush
 * incrementally instead of collecting every plaintext first.
 *   H2AES_ERR_NOT_SCHEDULED  no key scheduled yet (the reference's expect, aes128.rs:170)
 *   H2AES_ERR_CAPACITY       held + n exceeds the capacity (the reference's panic, :160-162);
 *                            nothing is written
 *   advice_dev               the SAME buffer as the h2aes_witness_dev call (it is completed in
 *                            place; rows of blocks not yet appended stay zero)
 *   ciphertexts_dev          n*16 bytes for the appended blocks, nullable
 *   *first_block             out, nullable: index the first appended block received
 * Unsharded circuits only. */
int32_t h2aes_encrypt_append_dev(h2aes_circuit *c, const uint8_t *plaintexts_dev, uint64_t n,
                                 uint64_t *advice_dev, uint8_t *ciphertexts_dev,
                                 uint64_t *first_block);

/* replaces: the chips' own per-cell assign methods called outside schedule_key / encrypt --
 * U8RangeCheckChip::range_check (u8_range_check_chip.rs:51-70, tag 1), U8XorChip::xor
 * (u8_xor_chip.rs:63-100, tag 2; value rule utils.rs:8-19), SboxChip::substitute
 * (sbox_chip.rs:57-83, tag 3; the context's S-box, i.e. the reference's table with its
 * S_BOX[255] by default), MulBy2Chip / MulBy3Chip::mul (gf_mul_chip.rs:59-89, tags 4 / 5).
 * Each of those calls assigns a ONE-ROW region; this call produces n such rows at once:
 *   x_dev, y_dev   n input bytes each (y_dev only for tag 2)
 *   cells_dev      nullable; [3][n][4] u64: the region's cells of row i in the chip's columns
 *                  (x | x, y, z | x, y=out), Montgomery limbs, unused columns zero
 *   out_dev        nullable; n bytes: the value of the cell the method returns (range: x)
 * The rows are NOT placed in a circuit: only the two circuit shapes of the reference have a
 * layout here (their chip rows come from the fused witness kernels), so a circuit that places
 * chip regions of its own supplies selectors and copy constraints for them itself. */
int32_t h2aes_chip_rows_dev(h2aes_ctx *ctx, int32_t tag, const uint8_t *x_dev, const uint8_t *y_dev,
                            uint64_t n, uint64_t *cells_dev, uint8_t *out_dev);

/* the 176 expanded key bytes (values of the 11x16 cells schedule_keys returns). Host out. */
int32_t h2aes_round_keys(h2aes_circuit *c, uint8_t *out176);

/* replaces: load_enc_full_table (table.rs:18-192, + halo2's padding of table columns), the
 * assign_fixed of round_constants (key_schedule.rs:169) and every Selector::enable.
 *   fixed      5 columns (4 table columns, round_constants)
 *   selectors  n_selectors columns of Fr 0/1, order = h2aes_layout.n_selectors
 * Input-independent: depends only on the shape and n_blocks. Either pointer may be NULL. */
int32_t h2aes_fixed_columns_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed_dev,
                                uint64_t *selectors_dev);
int32_t h2aes_fixed_columns_host(h2aes_circuit *c, uint64_t n_blocks, uint64_t *fixed,
                                 uint64_t *selectors);

/* device variant of h2aes_copy_constraints_host: all 640 + 1952*n_blocks copies. */
int32_t h2aes_copy_constraints_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *copies_dev);

/* replaces: halo2_proofs v0.3.0 plonk/lookup/prover.rs Argument::commit_permuted up to (not
 * including) the commitments, for lookups [first, first+count) of the circuit's witness
 * (call sites: main.rs:94-101, benches/aes128.rs:100-107, benches/key_schedule.rs:117-124).
 *   theta      host, 4 u64 Montgomery limbs
 *   A, S       compressed input / table expressions over all 2^k rows (nullable)
 *   Ap, Sp     permuted input / table columns; rows >= 2^k-6 are left zero (halo2: blinding)
 * each output is [count][2^k][4]. Requires a prior h2aes_witness_* on this circuit. */
int32_t h2aes_lookup_permuted_dev(h2aes_circuit *c, const uint64_t *theta, int32_t first,
                                  int32_t count, uint64_t *A_dev, uint64_t *S_dev,
                                  uint64_t *Ap_dev, uint64_t *Sp_dev);
int32_t h2aes_lookup_permuted_host(h2aes_circuit *c, const uint64_t *theta, int32_t first,
                                   int32_t count, uint64_t *A, uint64_t *S, uint64_t *Ap,
                                   uint64_t *Sp);

/* replaces: plonk/lookup/prover.rs Permuted::commit_product up to the commitment.
 *   Z  [count][2^k][4]; entries 0..2^k-6 are the grand product (Z[0]=1, Z[2^k-6]=1 for a
 *      satisfied lookup), the last 5 (halo2: blinding) are left zero.
 * Requires h2aes_lookup_permuted_* with the same theta to have run for these lookups. */
int32_t h2aes_lookup_product_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                 int32_t first, int32_t count, uint64_t *Z_dev);
int32_t h2aes_lookup_product_host(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                  int32_t first, int32_t count, uint64_t *Z);

/* the _dev lookup entry points are asynchronous; conditions the kernels detect on the device
 * (H2AES_ERR_LOOKUP, H2AES_ERR_ZERO_DENOM) are reported by this call, which synchronizes the
 * stream.  The _host entry points call it themselves. */
int32_t h2aes_lookup_check(h2aes_circuit *c);

/* statistics of the last grand product, per lookup: rows (of this shard) whose factor
 * (A+beta)(S+gamma)/((A'+beta)(S'+gamma)) is not identically one — the rows that cost a multiply
 * in each of the chunk-product and propagate kernels.  Host out, n_lookups entries; synchronizes. */
int32_t h2aes_lookup_stats(h2aes_circuit *c, uint64_t *nonconst_rows);

/* permutation argument on the device (unsharded circuits, k <= 24).
 *   sigma_dev  [n_columns][2^k][4]: sigma_p(omega^i) = delta^p' * omega^i' (keygen side;
 *              replaces Assembly::build_pk's `permutations`)
 *   Zp_dev     [n_sets][2^k][4]: the grand products of Argument::commit, one per chunk of 3
 *              columns, set s starting at set s-1's Z[2^k-6]; entries 0..2^k-6 deterministic,
 *              the last 5 (halo2: blinding) left zero
 *   advice_dev the [n_advice][2^k][4] buffer h2aes_witness_dev filled (rows >= 2^k-6 are not read) */
int32_t h2aes_perm_sigma_dev(h2aes_circuit *c, uint64_t n_blocks, uint64_t *sigma_dev);
int32_t h2aes_perm_product_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                               const uint64_t *advice_dev, uint64_t *Zp_dev);

/* ---------------------------------------------------------------------------------------------
 * Evaluation domain — SURVEY 8(f) items 3/4: the transforms on either side of the witness path.
 * replaces: halo2_proofs v0.3.0 poly/domain.rs EvaluationDomain::{new, lagrange_to_coeff,
 * coeff_to_extended, extended_to_coeff} and arithmetic.rs best_fft, as keygen_pk applies them to
 * the fixed / sigma columns and create_proof to the advice, A', S', Z columns
 * (call sites main.rs:78-101, benches/aes128.rs:91-107, benches/key_schedule.rs:108-124).
 *   j   cs.degree() (5 for both circuits: the lookups are degree 2 + 2 + 1); quotient_poly_degree
 *       = j - 1; extended_k = smallest with 2^extended_k >= 2^k * (j - 1)  (k + 2 here)
 * Columns are [ncols][len][4] as everywhere else.  The extended domain is the coset
 * zeta * <extended_omega>, zeta = Fr::ZETA = 7^((r-1)/3) (g_coset), exactly as halo2 evaluates it.
 * Limits: 2 <= k, extended_k <= 24 — i.e. k <= 22 for these circuits (extended_k = k + 2): a
 * transform is two passes of at most 2^11-point sub-transforms in shared memory up to 2^22
 * points, three passes for 2^23 and 2^24.  h2aes_domain_create returns H2AES_ERR_ARG
 * ("domain: 2 <= k, extended_k <= 24") beyond that.  The quotient entry points below are tested
 * up to k = 20 (a K=22 circuit's extended columns are 512 MiB each).
 * ------------------------------------------------------------------------------------------- */
typedef struct h2aes_domain h2aes_domain;
#define H2AES_DOMAIN_FFT 0               /* best_fft(a, omega, k): a'[i] = sum_t a[t] omega^(it)  */
#define H2AES_DOMAIN_LAGRANGE_TO_COEFF 1 /* EvaluationDomain::lagrange_to_coeff                   */
#define H2AES_DOMAIN_COEFF_TO_EXTENDED 2 /* EvaluationDomain::coeff_to_extended                   */
#define H2AES_DOMAIN_EXTENDED_TO_COEFF 3 /* EvaluationDomain::extended_to_coeff                   */
/* host-only: EvaluationDomain::new's extended_k and the length extended_to_coeff truncates to. */
int32_t h2aes_domain_query(int32_t j, int32_t k, int32_t *extended_k, uint64_t *quotient_len);
int32_t h2aes_domain_create(h2aes_ctx *ctx, int32_t j, int32_t k, h2aes_domain **out);
int32_t h2aes_domain_destroy(h2aes_domain *d);
/* in place on ncols columns of 2^k elements; inverse != 0 uses omega^-1 (no 1/n scaling). */
int32_t h2aes_fft_dev(h2aes_domain *d, uint64_t *cols_dev, int32_t ncols, int32_t inverse);
/* in place: evaluations over <omega> -> coefficients (inverse transform and the 1/n divisor). */
int32_t h2aes_lagrange_to_coeff_dev(h2aes_domain *d, uint64_t *cols_dev, int32_t ncols);
/* coeff_dev [ncols][2^k][4] -> ext_dev [ncols][2^extended_k][4]: coefficient t times zeta^t, zero
 * padding, forward transform with extended_omega. */
int32_t h2aes_coeff_to_extended_dev(h2aes_domain *d, const uint64_t *coeff_dev, int32_t ncols,
                                    uint64_t *ext_dev);
/* in place on [ncols][2^extended_k][4]: inverse transform, 1/2^extended_k, zeta^-t; the
 * coefficients past quotient_len (halo2 truncates the vector there) are written as zero. */
int32_t h2aes_extended_to_coeff_dev(h2aes_domain *d, uint64_t *ext_dev, int32_t ncols);
/* any of the four with HOST buffers (copies inside): in = ncols columns of the op's input length,
 * out = ncols columns of its output length (2^k, 2^k, 2^extended_k, quotient_len). */
int32_t h2aes_domain_transform_host(h2aes_domain *d, int32_t op, const uint64_t *in, int32_t ncols,
                                    uint64_t *out);

/* ---- lookup terms of the quotient polynomial — SURVEY 8(f) item 4 ---------------------------
 * replaces: halo2_proofs v0.3.0 plonk/evaluation.rs Evaluator::evaluate_h, lookup part, for lookups
 * [first, first+count) in meta.lookup order (call sites: create_proof, main.rs:94-101).  For each
 * lookup the running accumulator takes five steps values = values*y + expr with the expressions
 *   (1 - z) l0;  (z^2 - z) l_last;  (z(wX)(a'+beta)(s'+gamma) - z (A+beta)(S+gamma)) l_active_row;
 *   (a' - s') l0;  (a' - s')(a' - a'(X/w)) l_active_row
 * evaluated on the extended coset; A, S = theta-folds of the lookup's input expressions
 * (q*tag, q*x, ...: the configure functions under src/chips) and of its table columns.
 *   d           domain (j = 5, k = the circuit's); l0 / l_last / l_active_row are built on first use
 *               as keygen.rs does (blinding_factors = 5)
 *   advice_ext  [n_advice][2^extended_k][4]  extended cosets of the advice polynomials
 *   table_ext   [4][2^extended_k][4]         ... of the 4 table columns
 *   sel_ext     [count][2^extended_k][4]     ... of the selectors of lookups first..first+count-1
 *   Ap/Sp/Z_ext [count][2^extended_k][4]     ... of A', S', Z of the same lookups
 *   theta, beta, gamma, y   host, 4 u64 Montgomery limbs each
 *   values_ext  [2^extended_k][4] in/out: the accumulator (gates and permutation terms come first
 *               in halo2's order; start from zero for the lookup terms alone) */
int32_t h2aes_lookup_quotient_dev(h2aes_circuit *c, h2aes_domain *d, const uint64_t *theta,
                                  const uint64_t *beta, const uint64_t *gamma, const uint64_t *y,
                                  int32_t first, int32_t count, const uint64_t *advice_ext,
                                  const uint64_t *table_ext, const uint64_t *sel_ext,
                                  const uint64_t *Ap_ext, const uint64_t *Sp_ext,
                                  const uint64_t *Z_ext, uint64_t *values_ext);
/* replaces: the part of Evaluator::evaluate_h that comes BEFORE the lookups in halo2's order:
 *   custom gates  values = values*y + q_eq_rcon (words_column - round_constants)
 *                 (create_gate "Equality RC", key_schedule.rs:59-64; the simple selector is a 0/1
 *                 fixed column after keygen's compress_selectors)
 *   permutation   values = values*y + (1 - z_0) l0
 *                 values = values*y + (z_last^2 - z_last) l_last
 *                 values = values*y + (z_i - z_{i-1}(w^-6 X)) l0                 for sets i >= 1
 *                 values = values*y + (z_i(wX) prod_j (v_j + beta sigma_j + gamma)
 *                                      - z_i(X) prod_j (v_j + delta^p beta X + gamma)) l_active_row
 *   advice_ext  [n_advice][2^extended_k][4];  rc_ext, qrc_ext  [2^extended_k][4]: extended cosets of
 *               round_constants and of the q_eq_rcon selector column;  sigma_ext [n_columns][..]
 *               and Zp_ext [n_sets][..]: cosets of h2aes_perm_sigma_dev / h2aes_perm_product_dev
 *               columns (h2aes_perm_query order).  values_ext in/out as above. */
int32_t h2aes_gate_perm_quotient_dev(h2aes_circuit *c, h2aes_domain *d, const uint64_t *beta,
                                     const uint64_t *gamma, const uint64_t *y,
                                     const uint64_t *advice_ext, const uint64_t *rc_ext,
                                     const uint64_t *qrc_ext, const uint64_t *sigma_ext,
                                     const uint64_t *Zp_ext, uint64_t *values_ext);
/* replaces: EvaluationDomain::divide_by_vanishing_poly: element i of every column times
 * 1 / ((zeta * extended_omega^i)^n - 1).  In place on [ncols][2^extended_k][4]. */
int32_t h2aes_divide_by_vanishing_dev(h2aes_domain *d, uint64_t *ext_dev, int32_t ncols);

/* ---- commitments: multi-scalar multiplication over bn256 G1 — SURVEY 8(f) item 2 -------------
 * replaces: halo2_proofs v0.3.0 arithmetic.rs best_multiexp as reached through
 * ParamsKZG::commit_lagrange / commit in keygen and create_proof (call sites main.rs:76-101,
 * benches/aes128.rs:88-107, benches/key_schedule.rs:105-124): out = sum_i scalars[i] * bases[i].
 *   bases    n points as halo2curves `G1Affine`: x then y, 4 u64 Montgomery limbs of Fq each
 *            (64 bytes per point, identity = all zero); device memory, caller-owned, must outlive
 *            the handle (e.g. params.g_lagrange uploaded once)
 *   scalars  [ncols][n][4]: columns of Fr exactly as every other entry point produces them
 *   out      [ncols][2][4]: one G1Affine per column
 * The blinding term of a halo2 commitment (blind * params.w) is the caller's: append it as one more
 * (scalar, base) pair or add it to the result. */
/* Synthetic bases for tests and benchmarks (the stand-in for ParamsKZG::setup's unsafe SRS, whose
 * points halo2 derives from a random scalar): bases_dev[i] = (1 + i*d) * G with G = (1, 2) and d
 * an Fr in Montgomery form (host).  n points of 64 bytes, device memory. */
int32_t h2aes_g1_synthetic_bases_dev(h2aes_ctx *ctx, const uint64_t *d, uint64_t n, uint64_t *bases_dev);
typedef struct h2aes_msm h2aes_msm;
int32_t h2aes_msm_create(h2aes_ctx *ctx, const uint64_t *bases_dev, uint64_t n, h2aes_msm **out);
int32_t h2aes_msm_destroy(h2aes_msm *m);
/* Optional, once per handle: builds the running sums P_i = bases[0] + ... + bases[i] (64 bytes per
 * point more).  Afterwards every column is first differenced (s_i - s_{i+1}); when that leaves
 * clearly fewer non-zero scalars than the column has — the permuted lookup columns, Z, Z_perm are
 * constant over long runs — the bucket method runs on sum_i (s_i - s_{i+1}) P_i, the same group
 * element by Abel summation, hence the same output bits.  Synchronizes the stream. */
int32_t h2aes_msm_enable_runs(h2aes_msm *m);
/* scalars and out in device memory; asynchronous on the context's stream. */
int32_t h2aes_msm_dev(h2aes_msm *m, const uint64_t *scalars_dev, int32_t ncols, uint64_t *out_dev);
/* scalars and out in host memory (the bases stay on the device). */
int32_t h2aes_msm_host(h2aes_msm *m, const uint64_t *scalars, int32_t ncols, uint64_t *out);

/* ---------------------------------------------------------------------------------------------
 * Row-sharded single circuit (one process per GPU, shard g of G owns rows
 * [g*2^k/G, (g+1)*2^k/G) of every column).  The only cross-GPU data are the per-lookup bin
 * histograms (summed) and one 32-byte partial product per shard per lookup (gathered); the
 * library exposes them as plain device buffers so the caller can move them with NCCL
 * (ncclAllReduce / ncclAllGather) — see INTEGRATION.md and halo2_aes_b200/sharded.py.
 * Outputs hold only the shard's rows: [column][rows_per_shard][4].
 * ------------------------------------------------------------------------------------------- */
int32_t h2aes_circuit_set_shard(h2aes_circuit *c, int32_t shard, int32_t n_shards);
/* The same with a balanced ownership: the rows are cut into tiles of H2AES_SHARD_TILE_ROWS
 * consecutive rows (32 AES blocks) dealt round-robin — shard g holds tiles g, g+G, g+2G, ... back
 * to back.  With contiguous shards the end of every column (where A' is not zero and every row of
 * Z costs two multiplies) falls on the last shards; dealt tiles give every shard the same mix.
 * All outputs then hold the shard's local rows, [column][local_rows][4] (the shard holding the
 * last tile of the column holds fewer rows: h2aes_circuit_shard_query), and one partial product
 * per (lookup, tile of this shard) is exchanged instead of one per lookup. */
#define H2AES_SHARD_TILE_ROWS 43520
int32_t h2aes_circuit_set_shard_tiled(h2aes_circuit *c, int32_t shard, int32_t n_shards);
/* rows of every column this (shard of the) circuit holds, and the number of u64 words of its
 * partial products (h2aes_lookup_partial_dev; the all-gather's send count).  Either may be NULL. */
int32_t h2aes_circuit_shard_query(h2aes_circuit *c, uint64_t *rows, uint64_t *partial_words);
/* phase 1a: local histogram of table-bin hits; *hist_dev/ *hist_words give the buffer to
 * all-reduce (uint32 sum) across shards before phase 1b. */
int32_t h2aes_lookup_hist_dev(h2aes_circuit *c, const uint64_t *theta, uint32_t **hist_dev,
                              uint64_t *hist_words);
/* phase 1b: after the all-reduce: expand this shard's rows of A, S, A', S' (all lookups). */
int32_t h2aes_lookup_permuted_shard_dev(h2aes_circuit *c, uint64_t *A_dev, uint64_t *S_dev,
                                        uint64_t *Ap_dev, uint64_t *Sp_dev);
/* phase 2a: per-row factors of this shard's rows are written to Z_dev ([n_lookups][rows][4],
 * intermediate data), and their product, one Fr per lookup, to *partial_dev ([n_lookups][4]);
 * all-gather the partials into [n_shards][n_lookups][4] before phase 2b (tiled shards:
 * [n_shards][n_lookups][tiles per shard][4], see h2aes_circuit_shard_query). */
int32_t h2aes_lookup_partial_dev(h2aes_circuit *c, const uint64_t *beta, const uint64_t *gamma,
                                 uint64_t *Z_dev, uint64_t **partial_dev);
/* phase 2b: turns the factors in Z_dev (the same buffer as in phase 2a) into Z for this
 * shard's rows, given every shard's partial products. */
int32_t h2aes_lookup_product_shard_dev(h2aes_circuit *c, const uint64_t *gathered_dev,
                                       uint64_t *Z_dev);

/* ---------------------------------------------------------------------------------------------
 * The same with the two exchanges done INSIDE the library over NCCL (the
 * reference has no communication layer at all, this is what a multi-GPU prover built on it would
 * bind).  libnccl.so.2 is resolved at run time (dlopen) on the first h2aes_comm_* call, so the
 * library still loads on a box without NCCL; a missing library or a failing NCCL call returns
 * H2AES_ERR_NCCL with the NCCL error text in h2aes_ctx_last_error.
 *   rank 0:      h2aes_comm_unique_id(id)          -> ship the 128 bytes to every rank (any channel)
 *   every rank:  h2aes_comm_init(ctx, id, rank, world)   (collective; binds the ctx's device)
 *                h2aes_circuit_set_shard(c, rank, world); h2aes_witness_dev(...)
 *                h2aes_lookup_columns_sharded_dev(c, theta, beta, gamma, Ap, Sp, Z)
 * h2aes_lookup_columns_sharded_dev = hist -> ncclAllReduce(sum, all lookups in ONE call) ->
 * permuted (this shard's rows of A', S') -> partial products -> ncclAllGather (one 32-byte Fr per
 * lookup per shard, ONE call) -> Z of this shard's rows; everything is enqueued on the context's
 * stream, nothing synchronizes.  beta/gamma NULL stops after A', S' (the prover's order: the
 * challenges beta, gamma only exist after A', S' are committed); a later call with Ap_dev ==
 * Sp_dev == NULL and beta/gamma given finishes Z from the same theta.
 * ------------------------------------------------------------------------------------------- */
#define H2AES_COMM_ID_BYTES 128
int32_t h2aes_comm_unique_id(uint8_t *id /* H2AES_COMM_ID_BYTES */);
int32_t h2aes_comm_init(h2aes_ctx *ctx, const uint8_t *id /* H2AES_COMM_ID_BYTES */, int32_t rank,
                        int32_t world);
int32_t h2aes_comm_destroy(h2aes_ctx *ctx);
int32_t h2aes_lookup_columns_sharded_dev(h2aes_circuit *c, const uint64_t *theta,
                                         const uint64_t *beta, const uint64_t *gamma,
                                         uint64_t *Ap_dev, uint64_t *Sp_dev, uint64_t *Z_dev);

/* ---------------------------------------------------------------------------------------------
 * Debug / test hooks: the same host+device source the kernels use, run on the host, so
 * CPU-only tests can check the field arithmetic and closed-form layout without a GPU.
 * ------------------------------------------------------------------------------------------- */
int32_t h2aes_dbg_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *out);
int32_t h2aes_dbg_fr_add(const uint64_t *a, const uint64_t *b, uint64_t *out);
int32_t h2aes_dbg_fr_sub(const uint64_t *a, const uint64_t *b, uint64_t *out);
int32_t h2aes_dbg_fr_inv(const uint64_t *a, uint64_t *out);        /* binary extended Euclid */
int32_t h2aes_dbg_fr_inv_fermat(const uint64_t *a, uint64_t *out); /* a^(r-2), cross-check   */
int32_t h2aes_dbg_fr_from_u64(uint64_t v, uint64_t *out);
int32_t h2aes_dbg_fr_to_canonical(const uint64_t *a, uint64_t *out);
/* n Montgomery products out[i] = a[i]*b[i] computed ON THE DEVICE by the kernels' multiply
 * (host pointers; pins the generated PTX multiply against host arithmetic in the GPU tests). */
int32_t h2aes_dbg_fr_mul_dev(h2aes_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t n,
                             uint64_t *out);
int32_t h2aes_dbg_tables(int32_t sbox_mode, uint8_t *sbox256, uint8_t *mul2_256,
                         uint8_t *mul3_256);
/* host evaluation of the per-row trace word (a0 | a1<<8 | a2<<16 | op<<24) the witness
 * kernels produce for one AES block, 1360 words; used to pin the device byte pipeline. */
int32_t h2aes_dbg_block_trace(int32_t sbox_mode, const uint8_t *round_keys176,
                              const uint8_t *plaintext16, uint32_t *trace1360);
int32_t h2aes_dbg_key_schedule_trace(int32_t sbox_mode, const uint8_t *key16,
                                     uint8_t *round_keys176, uint32_t *trace400,
                                     uint8_t *words96);

#ifdef __cplusplus
}
#endif
#endif /* H2AES_H */
