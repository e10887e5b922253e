"""CPU-only checks of the product library: it loads, exports everything include/h2aes.h
declares, and its host+device shared source (closed-form layout, Fr arithmetic, byte pipeline)
agrees with the oracle's region-by-region replay.  No GPU compute is called here."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT, SEED, splitmix_bytes

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def _p(a, t):
    return a.ctypes.data_as(t)


def test_exports_every_declared_symbol(ffi):
    hdr = open(os.path.join(ROOT, "include", "h2aes.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(h2aes_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    L = C.CDLL(ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/h2aes.h but not exported"
    assert declared == set(ffi.SIGNATURES), declared ^ set(ffi.SIGNATURES)
    assert ffi.lib().h2aes_abi_version() == 1


def test_layout_query_matches_configure(ffi, orc, golden):
    # src/aes128.rs:46-141 / key_schedule.rs:40-77 column, selector and lookup counts
    for kind, k, N in ((ffi.KIND_AES, 20, 5), (ffi.KIND_AES, 20, 3), (ffi.KIND_AES, 17, 1),
                       (ffi.KIND_KEYSCHED, 17, 1), (ffi.KIND_AES, 22, 5)):
        lay = ffi.layout_query(kind, k, N)
        ref = orc.shape(kind, k, N)
        for f in ("n_advice", "n_fixed", "n_selectors", "n_lookups", "n_sets"):
            assert lay[f] == ref[f], f
        assert lay["n_rows"] == 1 << k and lay["usable_rows"] == (1 << k) - 6
        assert lay["rows_per_block"] == golden["AES_ROWS"]
        if kind == ffi.KIND_AES:
            c0 = ((1 << k) - golden["KEY_SCHEDULE_ROWS"]) // golden["AES_ROWS"]
            assert lay["capacity_set0"] == c0
            assert lay["capacity_blocks"] == c0 + (N - 1) * ((1 << k) // golden["AES_ROWS"])
    # K=20,N=5 holds 3853 blocks; main.rs shape 3082; unit-test shape 2311
    assert ffi.layout_query(ffi.KIND_AES, 20, 5)["capacity_blocks"] == 3853
    assert ffi.layout_query(ffi.KIND_AES, 20, 4)["capacity_blocks"] == 3082
    assert ffi.layout_query(ffi.KIND_AES, 20, 3)["capacity_blocks"] == 2311


def test_error_codes_without_gpu_work(ffi):
    L = ffi.lib()
    lay = ffi.Layout()
    assert L.h2aes_layout_query(ffi.KIND_AES, 16, 1, C.byref(lay)) == ffi.ERR_ROWS
    assert L.h2aes_layout_query(7, 17, 1, C.byref(lay)) == ffi.ERR_ARG
    assert L.h2aes_layout_query(ffi.KIND_AES, 17, 0, C.byref(lay)) == ffi.ERR_ARG
    assert L.h2aes_layout_query(ffi.KIND_AES, 17, 1, None) == ffi.ERR_ARG
    with pytest.raises(ffi.CapacityError):
        ffi.block_position(20, 5, 3853)
    assert ffi.block_position(20, 5, 3852) == (4, 770 * 1360)
    assert ffi.block_position(20, 5, 0) == (0, 400)
    assert ffi.block_position(20, 5, 769) == (1, 0)
    assert L.h2aes_status_string(ffi.ERR_CAPACITY).decode().startswith("AES calls too many")
    assert L.h2aes_status_string(ffi.ERR_NOT_SCHEDULED).decode() == "Keys should be scheduled"
    assert L.h2aes_circuit_create(None, 0, 17, 1, None) == ffi.ERR_ARG
    assert L.h2aes_witness_dev(None, None, None, 0, None, None) == ffi.ERR_ARG


def test_no_cpu_fallback(ffi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert ffi.lib().h2aes_ctx_create(0, 0, C.byref(h)) == ffi.ERR_CUDA
    assert not h.value
    from halo2_aes_b200.backend import Context
    with pytest.raises(RuntimeError):
        Context(0)


def test_fr_arithmetic_host_build(ffi, orc):
    """fr.cuh compiled for the host (the same source the kernels use) against Python bigints."""
    L = ffi.lib()
    rng = np.random.default_rng(SEED)
    out = np.zeros(4, np.uint64)
    cases = [(0, 0), (1, 1), (R_MOD - 1, R_MOD - 1), (R_MOD - 1, 1), (2, (R_MOD + 1) // 2)]
    cases += [(int.from_bytes(rng.bytes(32), "little") % R_MOD,
               int.from_bytes(rng.bytes(32), "little") % R_MOD) for _ in range(300)]
    for a, b in cases:
        A, B = orc.fr_from_int(a), orc.fr_from_int(b)
        assert L.h2aes_dbg_fr_mul(_p(A, u64p), _p(B, u64p), _p(out, u64p)) == 0
        assert orc.fr_to_int(out) == a * b % R_MOD
        L.h2aes_dbg_fr_add(_p(A, u64p), _p(B, u64p), _p(out, u64p))
        assert orc.fr_to_int(out) == (a + b) % R_MOD
        L.h2aes_dbg_fr_sub(_p(A, u64p), _p(B, u64p), _p(out, u64p))
        assert orc.fr_to_int(out) == (a - b) % R_MOD
        L.h2aes_dbg_fr_to_canonical(_p(A, u64p), _p(out, u64p))
        assert orc.limbs_to_int(out) == a
    for a, _ in cases[:40]:
        A = orc.fr_from_int(a)
        L.h2aes_dbg_fr_inv(_p(A, u64p), _p(out, u64p))
        assert orc.fr_to_int(out) == (pow(a, -1, R_MOD) if a else 0)
    for v in (0, 1, 255, 1 << 40, (1 << 64) - 1):
        L.h2aes_dbg_fr_from_u64(v, _p(out, u64p))
        assert (out == orc.fr_from_int(v)).all()


def test_tables_match_reference_constants(ffi, golden):
    s, m2, m3 = (np.zeros(256, np.uint8) for _ in range(3))
    assert ffi.lib().h2aes_dbg_tables(ffi.SBOX_REFERENCE, _p(s, u8p), _p(m2, u8p), _p(m3, u8p)) == 0
    assert s.tolist() == golden["S_BOX"]
    assert m2.tolist() == golden["MUL_BY_2"] and m3.tolist() == golden["MUL_BY_3"]
    ffi.lib().h2aes_dbg_tables(ffi.SBOX_FIPS197, _p(s, u8p), None, None)
    assert s[255] == 0x16


def _trace_cols(words):
    w = np.asarray(words, dtype=np.uint32)
    return w & 255, (w >> 8) & 255, (w >> 16) & 255, w >> 24


def _col_bytes(orc, col):
    """Montgomery column -> small ints (all values on this path are bytes)."""
    inv = pow((1 << 256) % R_MOD, -1, R_MOD)
    uniq, idx = np.unique(col, axis=0, return_inverse=True)
    vals = np.array([orc.limbs_to_int(u) * inv % R_MOD for u in uniq], dtype=np.int64)
    return vals[idx.ravel()]


@pytest.mark.parametrize("case", ["zero", "random", "ff"])
def test_closed_form_block_trace_equals_region_replay(ffi, orc, case):
    """layout.h (what the kernels evaluate per row) vs the oracle's floor-planner replay of
    aes128.rs:154-301 — two independent derivations of every (column,row,value)."""
    if case == "zero":
        key, pts = bytes(16), np.zeros((2, 16), np.uint8)
    elif case == "ff":
        key, pts = bytes([0xFF] * 16), np.full((2, 16), 0xFF, np.uint8)
    else:
        key = splitmix_bytes(SEED + 11, 16).tobytes()
        pts = splitmix_bytes(SEED + 12, 16 * 5).reshape(-1, 16)
    ref = orc.Circuit(orc.KIND_AES, 17, 1).synthesize(key, pts)
    L = ffi.lib()
    rk = np.zeros(176, np.uint8)
    tr = np.zeros(400, np.uint32)
    words = np.zeros(96, np.uint8)
    keya = np.frombuffer(key, np.uint8).copy()
    assert L.h2aes_dbg_key_schedule_trace(0, _p(keya, u8p), _p(rk, u8p), _p(tr, u32p), _p(words, u8p)) == 0
    assert (rk == ref.round_keys()).all()
    cols = [_col_bytes(orc, ref.advice(j)) for j in range(4)]
    sel = ref.selectors()
    a0, a1, a2, op = _trace_cols(tr)
    assert (cols[0][:400] == a0).all() and (cols[1][:400] == a1).all() and (cols[2][:400] == a2).all()
    assert (cols[3][:96] == words).all() and not cols[3][96:].any()
    for t in range(5):
        assert (sel[t][:400] == (op == t + 1)).all()
    for b in range(pts.shape[0]):
        bt = np.zeros(1360, np.uint32)
        assert L.h2aes_dbg_block_trace(0, _p(rk, u8p), _p(pts[b].copy(), u8p), _p(bt, u32p)) == 0
        a0, a1, a2, op = _trace_cols(bt)
        s, row = ffi.block_position(17, 1, b)
        assert (s, row) == (0, 400 + 1360 * b)
        sl = slice(row, row + 1360)
        assert (cols[0][sl] == a0).all() and (cols[1][sl] == a1).all() and (cols[2][sl] == a2).all()
        for t in range(5):
            assert (sel[t][sl] == (op == t + 1)).all()
        assert (a2[1344:1360] == ref.ciphertexts()[b]).all()
    end = 400 + 1360 * pts.shape[0]
    assert not cols[0][end:].any() and not cols[1][end:].any() and not cols[2][end:].any()


def test_row_op_equals_selector_columns(ffi, orc):
    k, N = 17, 2
    lay = ffi.layout_query(ffi.KIND_AES, k, N)
    nb = lay["capacity_set0"] + 3
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(bytes(16), np.zeros((nb, 16), np.uint8))
    sel = ref.selectors()
    L = ffi.lib()
    rows = np.unique(np.concatenate([np.arange(0, 3000), np.arange(129000, 1 << k),
                                     np.random.default_rng(1).integers(0, 1 << k, 4000)]))
    for s in range(N):
        for row in rows:
            op = L.h2aes_row_op(ffi.KIND_AES, k, N, nb, s, int(row))
            got = [1 if op == t + 1 else 0 for t in range(5)]
            assert got == [int(sel[5 * s + t][row]) for t in range(5)], (s, row)
    # q_eq_rcon (simple selector, last) at rows 20+8(r-1) of words_column
    rc = np.flatnonzero(sel[5 * N])
    assert rc.tolist() == [20 + 8 * r for r in range(10)]
    assert L.h2aes_row_op(ffi.KIND_AES, k, N, lay["capacity_blocks"] + 1, 0, 0) == -ffi.ERR_CAPACITY


@pytest.mark.parametrize("kind,k,N,nb", [(1, 17, 1, 0), (0, 17, 1, 4), (0, 17, 2, 97)])
def test_copy_constraints_equal_region_replay(ffi, orc, kind, k, N, nb):
    pts = splitmix_bytes(SEED + 3, 16 * max(nb, 1)).reshape(-1, 16)[:nb]
    ref = orc.Circuit(kind, k, N).synthesize(splitmix_bytes(SEED + 4, 16), pts if nb else None)
    want = ref.copies()
    total = C.c_uint64()
    assert ffi.lib().h2aes_copy_count(kind, nb, C.byref(total)) == 0
    assert total.value == len(want) == 640 + 1952 * nb
    got = np.zeros((total.value, 4), np.uint64)
    assert ffi.lib().h2aes_copy_constraints_host(kind, k, N, nb, 0, total.value, _p(got, u64p)) == 0
    assert (got == want).all()
    part = np.zeros((10, 4), np.uint64)
    assert ffi.lib().h2aes_copy_constraints_host(kind, k, N, nb, 635, 10 if nb else 5, _p(part, u64p)) == 0
    assert (part[:5] == want[635:640]).all()
    assert ffi.lib().h2aes_copy_constraints_host(kind, k, N, nb, total.value, 1, _p(part, u64p)) == ffi.ERR_ARG


def test_active_rows_equal_selector_sums(ffi, orc):
    """h2aes_active_rows (closed form) == number of enabled rows of each complex selector."""
    for kind, k, N, nb in ((1, 17, 1, 0), (0, 17, 1, 0), (0, 17, 2, 97), (0, 17, 2, 191)):
        pts = np.zeros((nb, 16), np.uint8)
        ref = orc.Circuit(kind, k, N).synthesize(bytes(16), pts if nb else None)
        sel = ref.selectors()
        for li in range(ref.info["n_lookups"]):
            out = C.c_uint64()
            assert ffi.lib().h2aes_active_rows(kind, k, N, nb, li, C.byref(out)) == 0
            assert out.value == int(sel[li].sum()), (kind, N, nb, li)


def test_fr_inv_binary_gcd_vs_fermat_and_bigint(ffi, orc):
    """fr_inv (binary extended Euclid, what k_inv_tables runs) == a^(r-2) == pow(a, -1, r)."""
    L = ffi.lib()
    rng = np.random.default_rng(SEED + 5)
    vals = [0, 1, 2, 3, R_MOD - 1, R_MOD - 2, (R_MOD + 1) // 2, 1 << 253, (1 << 256) % R_MOD,
            0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    vals += [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(400)]
    vals += [int(x) for x in rng.integers(1, 1 << 20, 50)]
    out, out2 = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    for i, a in enumerate(vals):
        A = orc.fr_from_int(a)
        assert L.h2aes_dbg_fr_inv(_p(A, u64p), _p(out, u64p)) == 0
        assert orc.fr_to_int(out) == (pow(a, -1, R_MOD) if a else 0), hex(a)
        if i < 60:
            assert L.h2aes_dbg_fr_inv_fermat(_p(A, u64p), _p(out2, u64p)) == 0
            assert (out == out2).all()


@pytest.mark.parametrize("kind,k,N,nb", [(1, 17, 1, 0), (0, 17, 1, 3), (0, 17, 2, 120)])
def test_permutation_columns_and_mapping_equal_keygen_replay(ffi, orc, kind, k, N, nb):
    """Permutation argument bookkeeping: column order and the cycle structure built
    from the closed-form copy list == halo2's keygen Assembly replayed by the oracle over the
    copies recorded region by region."""
    L = ffi.lib()
    nc, cl, ns = C.c_int32(), C.c_int32(), C.c_int32()
    fx = (C.c_int32 * 256)()
    ix = (C.c_int32 * 256)()
    assert L.h2aes_perm_query(kind, k, N, C.byref(nc), C.byref(cl), C.byref(ns), fx, ix) == 0
    ps = orc.perm_shape(kind, k, N)
    assert (nc.value, cl.value, ns.value) == (ps["n_columns"], ps["chunk_len"], ps["n_sets"])
    assert [("fixed" if fx[i] else "advice", ix[i]) for i in range(nc.value)] == ps["columns"]
    pts = splitmix_bytes(SEED + 13, 16 * max(nb, 1)).reshape(-1, 16)[:nb]
    ref = orc.Circuit(kind, k, N).synthesize(bytes(16), pts if nb else None)
    want = ref.permutation(np.ones(4, np.uint64), np.ones(4, np.uint64), want=("mapping",))["mapping"]
    got = np.zeros_like(want)
    assert L.h2aes_perm_mapping_host(kind, k, N, nb, got.ctypes.data_as(u32p)) == 0
    assert (got == want).all()


def _random_shapes():
    """seeded (k, N, n_blocks) triples: boundaries of the column sets and random fills"""
    rng = np.random.default_rng(SEED + 99)
    shapes = []
    for k, N in ((17, 3), (18, 2), (18, 4)):
        n = 1 << k
        cap0, cap = (n - 1760) // 1360, n // 1360
        total = cap0 + (N - 1) * cap
        shapes += [(k, N, cap0), (k, N, cap0 + 1), (k, N, total),
                   (k, N, int(rng.integers(1, total)))]
    return shapes


@pytest.mark.parametrize("k,N,nb", _random_shapes())
def test_closed_form_layout_on_more_shapes(ffi, orc, k, N, nb):
    """block positions, enabled selectors per row, copy constraints and the permutation cycle
    mapping from the product's closed forms == the oracle's region-by-region replay, at the
    column-set boundaries (set 0 full, first block of the next set, circuit at capacity) and at a
    random fill, for k = 17, 18 and N = 2..4."""
    L = ffi.lib()
    pts = splitmix_bytes(SEED + 7 * nb + N, 16 * nb).reshape(nb, 16)
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(splitmix_bytes(SEED + nb, 16).tobytes(), pts)
    sel = ref.selectors()
    # selectors: every row of every set at a sample of rows, incl. the ends of the columns
    rng = np.random.default_rng(nb)
    rows = np.unique(np.concatenate([np.arange(0, 500), np.arange((1 << k) - 1500, 1 << k),
                                     rng.integers(0, 1 << k, 1500)]))
    for s in range(N):
        ops = np.array([L.h2aes_row_op(ffi.KIND_AES, k, N, nb, s, int(r)) for r in rows])
        for t in range(5):
            assert ((ops == t + 1) == (sel[5 * s + t][rows] != 0)).all(), (s, t)
    for li in range(5 * N):
        out = C.c_uint64()
        assert L.h2aes_active_rows(ffi.KIND_AES, k, N, nb, li, C.byref(out)) == 0
        assert out.value == int(sel[li].sum()), li
    # copies
    want = ref.copies()
    got = np.zeros((len(want), 4), np.uint64)
    assert L.h2aes_copy_constraints_host(ffi.KIND_AES, k, N, nb, 0, len(want), _p(got, u64p)) == 0
    assert (got == want).all()
    # permutation mapping
    wm = ref.permutation(np.ones(4, np.uint64), np.ones(4, np.uint64), want=("mapping",))["mapping"]
    gm = np.zeros_like(wm)
    assert L.h2aes_perm_mapping_host(ffi.KIND_AES, k, N, nb, gm.ctypes.data_as(u32p)) == 0
    assert (gm == wm).all()
    # one block past capacity is the reference's panic
    lay = ffi.layout_query(ffi.KIND_AES, k, N)
    s_, r_ = C.c_int32(), C.c_uint64()
    assert L.h2aes_block_position(k, N, lay["capacity_blocks"], C.byref(s_), C.byref(r_)) == ffi.ERR_CAPACITY


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_generated_ptx_multiply_is_in_sync_and_self_tests(field, tmp_path):
    """csrc/gen_fr_mul.py builds the device Montgomery multiply once, runs it on a python model of
    the PTX carry flag against big-integer arithmetic (20 k operand pairs incl. the extremes) and
    only then emits it: the committed fr_mul_ptx.inc / fq_mul_ptx.inc must be that output."""
    import importlib.util
    import shutil
    src = os.path.join(ROOT, "halo2_aes_b200", "csrc")
    work = tmp_path / "gen"
    work.mkdir()
    shutil.copy(os.path.join(src, "gen_fr_mul.py"), work / "gen_fr_mul.py")
    argv = sys.argv
    try:
        sys.argv = ["gen_fr_mul.py", field]
        spec = importlib.util.spec_from_file_location(f"gen_{field}", work / "gen_fr_mul.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ops = mod.build()
        assert mod.selftest(ops, iters=3000) > 3000
        text = mod.emit(ops)
    finally:
        sys.argv = argv
    assert text == open(os.path.join(src, f"{field}_mul_ptx.inc")).read()


def test_generated_rust_ffi_block_is_in_sync_with_the_header(ffi):
    """integration/rust/ffi.rs (the extern "C" block a maintainer of the reference would add) is
    the generator's output for the current include/h2aes.h and declares every exported symbol."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gen_rust_ffi", os.path.join(ROOT, "integration", "gen_rust_ffi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = mod.generate()
    assert text == open(os.path.join(ROOT, "integration", "rust", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (h2aes_\w+)\(", text))
    assert declared == set(ffi.SIGNATURES)
    # argument counts agree with the ctypes signatures
    for name, (_, args) in ffi.SIGNATURES.items():
        m = re.search(r"pub fn %s\((.*?)\) ->" % name, text)
        n_args = 0 if not m.group(1).strip() else m.group(1).count(":")
        assert n_args == len(args), name


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU:
    one step of the K=20, N=5 circuit through the oracle, one JSON line with the keys the
    contract names and the same `config` the GPU arm reports."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "blocks/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert cfg["k"] == 20 and cfg["column_sets"] == 5 and cfg["blocks_per_circuit"] == 3853
    assert "model" not in cfg
