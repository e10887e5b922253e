"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs — every advice, fixed, selector, copy, A, S, A', S' and Z entry compared
limb for limb (integer work: bit-exact is the bar) — plus the restated MockProver on the
GPU-produced columns.  Run on the B200 box: `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest

from conftest import SEED, splitmix_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(ffi):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from halo2_aes_b200.backend import Context
    ctx = Context(0)
    yield ctx
    ctx.close()


def _host(t):
    from halo2_aes_b200.backend import to_host_u64
    return to_host_u64(t)


def _sel_bits(orc, sel_fr):
    """Fr 0/1 selector columns -> uint8"""
    one = orc.fr_from_int(1)
    is_one = (sel_fr == one).all(axis=-1)
    is_zero = ~sel_fr.any(axis=-1)
    assert (is_one | is_zero).all()
    return is_one.astype(np.uint8)


def _run_case(gpu, orc, ffi, kind, k, N, key, pts, challenges, lookups="all", fips=False):
    import torch
    from halo2_aes_b200.backend import Circuit, Context
    ctx = gpu
    if fips:
        ctx = Context(0, ffi.SBOX_FIPS197)
    cir = Circuit(ctx, kind, k, N)
    nb = 0 if pts is None else len(pts)
    d_pts = torch.from_numpy(np.ascontiguousarray(pts)).cuda() if nb else None
    advice, cts = cir.witness(key, d_pts)
    fixed, sel = cir.fixed_columns()
    copies = cir.copy_constraints()
    ref = orc.Circuit(kind, k, N, fips_sbox=fips).synthesize(key, pts)

    g_adv = _host(advice)
    assert (g_adv == ref.advice()).all(), "advice"
    if nb:
        assert (cts.cpu().numpy() == ref.ciphertexts()).all(), "ciphertexts"
    assert (cir.round_keys() == ref.round_keys()).all()
    g_fixed = _host(fixed)
    assert (g_fixed == ref.fixed()).all(), "fixed"
    g_sel = _sel_bits(orc, _host(sel))
    assert (g_sel == ref.selectors()).all(), "selectors"
    g_copies = _host(copies)
    assert (g_copies == ref.copies()).all(), "copies"

    # restated MockProver on what the GPU produced (north_star: MockProver passing on the GPU witness)
    rc, where = orc.mock_verify_ext(kind, k, N, g_adv, g_fixed, g_sel, g_copies)
    assert rc == 0, (rc, where)

    theta, beta, gamma = challenges
    L = cir.layout["n_lookups"]
    idxs = range(L) if lookups == "all" else lookups
    perm = cir.lookup_permuted(theta)
    Z = cir.lookup_product(beta, gamma)
    cir.lookup_check()
    hp = {nm: _host(t) for nm, t in perm.items()}
    hZ = _host(Z)
    one = orc.fr_from_int(1)
    u = (1 << k) - 6
    for li in range(L):
        assert (hZ[li, 0] == one).all() and (hZ[li, u] == one).all(), f"Z ends, lookup {li}"
        assert not hZ[li, u + 1:].any()
    for li in idxs:
        r = ref.lookup(li, theta, beta, gamma)
        for nm in ("A", "S", "Ap", "Sp"):
            bad = np.flatnonzero((hp[nm][li] != r[nm]).any(axis=1))
            assert bad.size == 0, f"{nm} lookup {li}: {bad.size} rows differ, first {bad[:5]}"
        bad = np.flatnonzero((hZ[li] != r["Z"]).any(axis=1))
        assert bad.size == 0, f"Z lookup {li}: {bad.size} rows differ, first {bad[:5]}"
    # partial ranges give the same columns
    if L >= 3:
        p2 = cir.lookup_permuted(theta, first=1, count=2, want_inputs=False)
        z2 = cir.lookup_product(beta, gamma, first=2, count=1)
        assert (_host(p2["Ap"]) == hp["Ap"][1:3]).all() and (_host(p2["Sp"]) == hp["Sp"][1:3]).all()
        assert (_host(z2)[0] == hZ[2]).all()
    cir.close()
    if fips:
        ctx.close()
    return ref


def test_key_schedule_circuit(gpu, orc, ffi, challenges):
    """BASELINE config 2: key schedule only, k=17, 3 lookups (benches/key_schedule.rs)."""
    _run_case(gpu, orc, ffi, ffi.KIND_KEYSCHED, 17, 1, bytes(16), None, challenges)
    _run_case(gpu, orc, ffi, ffi.KIND_KEYSCHED, 17, 1, splitmix_bytes(SEED, 16).tobytes(), None,
              challenges)


def test_single_block_fips_vectors(gpu, orc, ffi, golden, challenges):
    """BASELINE config 1: one block + key schedule (FIPS-197 vectors), small shape."""
    for v in golden["fips197"]:
        ref = _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 1, bytes.fromhex(v["key"]),
                        np.frombuffer(bytes.fromhex(v["pt"]), np.uint8).reshape(1, 16), challenges)
        assert ref.ciphertexts()[0].tobytes().hex() == v["ct"]


def test_empty_aes_circuit(gpu, orc, ffi, challenges):
    _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 2, splitmix_bytes(SEED + 1, 16).tobytes(), None,
              challenges, lookups=[0, 1, 2, 3, 4, 6])


def test_random_blocks_two_sets_ragged(gpu, orc, ffi, challenges):
    """Random blocks crossing the column-set switch (aes128.rs:303-325): 95 in set 0, 5 in set 1."""
    pts = splitmix_bytes(SEED + 2, 16 * 100).reshape(-1, 16)
    pts[7, :] = 0xFF
    _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 2, splitmix_bytes(SEED + 3, 16).tobytes(), pts,
              challenges)


def test_full_capacity(gpu, orc, ffi, challenges):
    lay = ffi.layout_query(ffi.KIND_AES, 17, 2)
    nb = lay["capacity_blocks"]
    pts = splitmix_bytes(SEED + 4, 16 * nb).reshape(-1, 16)
    _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 2, splitmix_bytes(SEED + 5, 16).tobytes(), pts,
              challenges, lookups=[1, 2, 3, 4, 5, 6])


def test_fips_sbox_switch(gpu, orc, ffi, challenges):
    pts = np.full((2, 16), 0xFF, np.uint8)
    _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 1, bytes([0xFF] * 16), pts, challenges,
              lookups=[2], fips=True)


@pytest.mark.parametrize("fips", [False, True])
def test_chip_rows_over_their_whole_domain(gpu, orc, ffi, fips):
    """The chips' stand-alone assign methods (u8_xor_chip.rs:63-100, sbox_chip.rs:57-83,
    gf_mul_chip.rs:59-89, u8_range_check_chip.rs:51-70) as one-row regions: every input of every
    chip against the oracle's tables (constant.rs, incl. the reference's S_BOX[255]) and
    xor_bytes (utils.rs:8-19, test :41-47), values and Montgomery cells."""
    import torch
    from halo2_aes_b200.backend import Context
    ctx = Context(0, ffi.SBOX_FIPS197) if fips else gpu
    sbox, m2, m3 = orc.tables(fips)
    assert (sbox[255] == 0x16) == fips
    frtab = np.stack([orc.fr_from_int(v) for v in range(256)])
    zero = np.zeros(4, np.uint64)
    x = torch.arange(256, dtype=torch.uint8, device="cuda")
    for tag, tab in ((1, np.arange(256, dtype=np.uint8)), (3, sbox), (4, m2), (5, m3)):
        out, cells = ctx.chip_rows(tag, x)
        assert (out.cpu().numpy() == tab).all(), tag
        c = _host(cells)
        assert (c[0] == frtab).all() and (c[2] == zero).all()
        assert (c[1] == (frtab[tab] if tag != 1 else zero)).all()
    xs = x.repeat_interleave(256).contiguous()
    ys = x.repeat(256).contiguous()
    out, cells = ctx.chip_rows(2, xs, ys)
    want = (xs.cpu().numpy() ^ ys.cpu().numpy())
    assert (out.cpu().numpy() == want).all()
    c = _host(cells)
    assert (c[0] == frtab[xs.cpu().numpy()]).all() and (c[1] == frtab[ys.cpu().numpy()]).all()
    assert (c[2] == frtab[want]).all()
    out, _ = ctx.chip_rows(2, xs[:2304:256][5:6].contiguous(), ys[12:13].contiguous())
    assert int(out.item()) == 9                                          # utils.rs:41-47
    for a, b in ((5, 12), (255, 255), (0, 0)):                           # the oracle's xor_bytes on Fr
        o = np.zeros(4, np.uint64)
        assert orc.lib().orc_xor_bytes(orc._u64(orc.fr_from_int(a)), orc._u64(orc.fr_from_int(b)),
                                       orc._u64(o)) == 0
        assert orc.fr_to_int(o) == a ^ b
    with pytest.raises(ffi.H2AesError):
        ctx.chip_rows(6, x)
    assert ctx.chip_rows(2, x[:0], x[:0])[0].numel() == 0
    if fips:
        ctx.close()


def test_chip_methods_of_the_host_mirror(gpu, orc, ffi):
    """U8XorChip::xor etc. called the way a user's synthesize would, chained, on cells of a
    flushed AES circuit: ct[0] ^ ct[1], then Sbox, mul2, mul3, range check."""
    from halo2_aes_b200 import aes128 as m
    sbox, m2, m3 = orc.tables(False)
    cfg = m.FixedAes128Config.configure(m.ConstraintSystem(), 17, 1)
    lay = m.B200Layouter(17, gpu)
    cfg.schedule_key(lay, bytes(range(16)))
    ct = cfg.encrypt(lay, bytes.fromhex("00112233445566778899aabbccddeeff"))
    want_ct = bytes.fromhex("69c4e0d86a7b0430d8cdb78070b4c55a")                 # FIPS-197 C.1
    z = m.U8XorChip.construct(cfg.configs[1][0]).xor(lay, ct[0], ct[1])
    assert z.value() == want_ct[0] ^ want_ct[1]
    s = m.SboxChip.construct(cfg.configs[2][0]).substitute(lay, z)
    assert s.value() == sbox[z.value()]
    d = m.MulBy2Chip.construct(cfg.configs[3][0]).mul(lay, s)
    t3 = m.MulBy3Chip.construct(cfg.configs[4][0]).mul(lay, s)
    assert d.value() == m2[s.value()] and t3.value() == m3[s.value()]
    r = m.U8RangeCheckChip.construct(cfg.configs[0][0]).range_check(lay, d)
    assert r.value() == d.value() and len(lay.chip_regions) == 5
    assert (_host(z.cells)[2] == orc.fr_from_int(z.value())).all()
    with pytest.raises(ValueError):
        m.SboxChip.construct(cfg.configs[2][0]).substitute(lay, 256)


def test_error_behaviour(gpu, ffi):
    import torch
    from halo2_aes_b200.backend import Circuit
    cir = Circuit(gpu, ffi.KIND_AES, 17, 1)
    cap = cir.layout["capacity_blocks"]
    pts = torch.zeros((cap + 1, 16), dtype=torch.uint8, device="cuda")
    with pytest.raises(ffi.CapacityError):          # aes128.rs:160-162
        cir.witness(bytes(16), pts)
    with pytest.raises(ffi.KeysNotScheduled):        # aes128.rs:170
        cir.round_keys()
    with pytest.raises(ffi.H2AesError) as e:         # lookups before any witness
        cir.lookup_permuted(np.ones(4, np.uint64))
    assert e.value.code == ffi.ERR_STATE
    cir.witness(bytes(16), pts[:cap])
    with pytest.raises(ffi.H2AesError) as e:         # product before permuted
        cir.lookup_product(np.ones(4, np.uint64), np.ones(4, np.uint64))
    assert e.value.code == ffi.ERR_STATE
    with pytest.raises(ffi.NotEnoughRows):
        Circuit(gpu, ffi.KIND_AES, 16, 1)
    ks = Circuit(gpu, ffi.KIND_KEYSCHED, 17, 1)
    with pytest.raises(ffi.H2AesError):
        ks.witness(bytes(16), pts[:1])
    cir.close()
    ks.close()


def test_host_entry_points(gpu, orc, ffi, challenges):
    """The _host variants (host pointers in, host pointers out) give the same bits."""
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    k, N = 17, 1
    pts = splitmix_bytes(SEED + 8, 16 * 9).reshape(-1, 16)
    key = splitmix_bytes(SEED + 9, 16).tobytes()
    cir = Circuit(gpu, ffi.KIND_AES, k, N)
    n = 1 << k
    adv = np.zeros((4, n, 4), np.uint64)
    cts = np.zeros((9, 16), np.uint8)
    cir.witness_host(key, pts, adv, cts)
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    assert (adv == ref.advice()).all() and (cts == ref.ciphertexts()).all()
    L = ffi.lib()
    u64p = C.POINTER(C.c_uint64)
    A, S, Ap, Sp, Z = (np.zeros((5, n, 4), np.uint64) for _ in range(5))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    ffi.check(L.h2aes_lookup_permuted_host(cir._h, theta.ctypes.data_as(u64p), 0, 5, vp(A), vp(S),
                                           vp(Ap), vp(Sp)), gpu._h)
    ffi.check(L.h2aes_lookup_product_host(cir._h, beta.ctypes.data_as(u64p),
                                          gamma.ctypes.data_as(u64p), 0, 5, vp(Z)), gpu._h)
    fixed = np.zeros((5, n, 4), np.uint64)
    sel = np.zeros((6, n, 4), np.uint64)
    ffi.check(L.h2aes_fixed_columns_host(cir._h, 9, vp(fixed), vp(sel)), gpu._h)
    assert (fixed == ref.fixed()).all()
    assert (_sel_bits(orc, sel) == ref.selectors()).all()
    for li in range(5):
        r = ref.lookup(li, theta, beta, gamma)
        for nm, g in (("A", A), ("S", S), ("Ap", Ap), ("Sp", Sp), ("Z", Z)):
            assert (g[li] == r[nm]).all(), (nm, li)
    cir.close()


def _assemble(shards, parts):
    """[cols, local rows, 4] per shard -> [cols, n, 4] in column order (host)"""
    n = shards[0].n
    out = np.zeros((parts[0].shape[0], n, 4), np.uint64)
    for c, p in zip(shards, parts):
        out[:, c.global_rows()] = _host(p)
    return out


@pytest.mark.parametrize("G,tiled", [(2, False), (8, False), (2, True), (4, True)])
def test_row_sharded_circuit_emulated_on_one_gpu(gpu, orc, ffi, challenges, G, tiled):
    """BASELINE config 5 (one circuit row-sharded over G GPUs) with the G shards run one after
    another on this GPU and the two exchanges (histogram sum, partial-product gather) done by
    hand — the same entry points sharded.RowShardedCircuit drives over NCCL."""
    import torch
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    k, N = 17, 2
    pts = splitmix_bytes(SEED + 50, 16 * 120).reshape(-1, 16)
    key = splitmix_bytes(SEED + 51, 16).tobytes()
    d_pts = torch.from_numpy(pts).cuda()
    shards = [Circuit(gpu, ffi.KIND_AES, k, N) for _ in range(G)]
    for g, c in enumerate(shards):
        c.set_shard(g, G, tiled=tiled)
    assert sum(c.rows for c in shards) == 1 << k
    adv = [c.witness(key, d_pts)[0] for c in shards]
    fx = [c.fixed_columns() for c in shards]
    hists = [c.shard_hist(theta) for c in shards]
    total = torch.stack(hists).sum(dim=0, dtype=torch.int32)
    for h in hists:
        h.copy_(total)                                    # all-reduce(sum)
    perm = [c.shard_permuted(want_inputs=True) for c in shards]
    zp = [c.shard_partial(beta, gamma) for c in shards]
    gathered = torch.stack([p.clone() for _, p in zp]).contiguous()   # all-gather
    Zs = [c.shard_product(gathered, z) for c, (z, _) in zip(shards, zp)]
    for c in shards:
        c.lookup_check()
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    cat = lambda ts: _assemble(shards, ts)
    assert (cat(adv) == ref.advice()).all()
    assert (cat([f[0] for f in fx]) == ref.fixed()).all()
    assert (_sel_bits(orc, cat([f[1] for f in fx])) == ref.selectors()).all()
    got = {nm: cat([p[nm] for p in perm]) for nm in ("A", "S", "Ap", "Sp")}
    gZ = cat(Zs)
    for li in (0, 1, 2, 3, 4, 6, 7):
        r = ref.lookup(li, theta, beta, gamma)
        for nm in ("A", "S", "Ap", "Sp"):
            assert (got[nm][li] == r[nm]).all(), (nm, li)
        assert (gZ[li] == r["Z"]).all(), ("Z", li)
    for c in shards:
        c.close()


def test_fr_mul_device_vs_bigint(gpu, orc, ffi):
    """The generated PTX Montgomery multiply (csrc/fr_mul_ptx.inc) on the device against Python
    big integers, including extreme operands."""
    R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    rng = np.random.default_rng(SEED)
    vals = [0, 1, 2, R_MOD - 1, R_MOD - 2, (1 << 254) % R_MOD, (1 << 256) % R_MOD, R_MOD >> 1]
    a = vals * len(vals) + [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(4000)]
    b = [y for y in vals for _ in vals] + [int.from_bytes(rng.bytes(32), "little") % R_MOD
                                            for _ in range(4000)]
    A = np.stack([orc.int_to_limbs(x) for x in a])
    B = np.stack([orc.int_to_limbs(x) for x in b])
    out = np.zeros_like(A)
    u64p = C.POINTER(C.c_uint64)
    ffi.check(ffi.lib().h2aes_dbg_fr_mul_dev(gpu._h, A.ctypes.data_as(u64p), B.ctypes.data_as(u64p),
                                            len(a), out.ctypes.data_as(u64p)), gpu._h)
    rinv = pow(1 << 256, -1, R_MOD)
    for i in range(len(a)):
        assert orc.limbs_to_int(out[i]) == a[i] * b[i] * rinv % R_MOD, i


def _oracle_lookups(ref, idxs, theta, beta, gamma, want=("Ap", "Sp", "Z"), workers=8):
    """oracle columns of several lookups at once (the C calls release the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return dict(zip(idxs, ex.map(lambda li: ref.lookup(li, theta, beta, gamma, want=want), idxs)))


def _same_on_device(dev_col, ref_col) -> bool:
    import torch
    return torch.equal(dev_col, torch.from_numpy(ref_col.view(np.int64)).to(dev_col.device))


def _compare_all_lookups(cir, ref, perm, Z, theta, beta, gamma, idxs, chunk=5):
    for c0 in range(0, len(idxs), chunk):
        part = list(idxs[c0:c0 + chunk])
        got = _oracle_lookups(ref, part, theta, beta, gamma)
        for li in part:
            for nm, t in (("Ap", perm["Ap"]), ("Sp", perm["Sp"]), ("Z", Z)):
                assert _same_on_device(t[li], got[li][nm]), (nm, li)


def test_bench_shape_full_capacity(gpu, orc, ffi, challenges):
    """The benchmarked circuit itself (BASELINE config 4 unit: K=20, N=5, 3853 random blocks):
    every advice column and ALL 25 lookups (A', S', Z) column for column against the oracle
    (compared on the device; the oracle columns are uploaded once)."""
    import torch
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    k, N = 20, 5
    nb = ffi.layout_query(ffi.KIND_AES, k, N)["capacity_blocks"]
    assert nb == 3853
    pts = splitmix_bytes(SEED + 70, 16 * nb).reshape(-1, 16)
    key = splitmix_bytes(SEED + 71, 16).tobytes()
    cir = Circuit(gpu, ffi.KIND_AES, k, N)
    advice, cts = cir.witness(key, torch.from_numpy(pts).cuda())
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    assert (cts.cpu().numpy() == ref.ciphertexts()).all()
    for col in range(16):
        assert _same_on_device(advice[col], ref.advice(col)), col
    del advice
    perm = cir.lookup_permuted(theta, want_inputs=False)
    Z = cir.lookup_product(beta, gamma)
    cir.lookup_check()
    one = orc.fr_from_int(1)
    n = 1 << k
    assert (_host(Z[:, n - 6]) == one).all() and (_host(Z[:, 0]) == one).all()
    _compare_all_lookups(cir, ref, perm, Z, theta, beta, gamma, list(range(25)))
    cir.close()


def test_single_block_at_the_bench_shape(gpu, orc, ffi, golden, challenges):
    """BASELINE config 1 at the shape benches/aes128.rs uses (K=20, N=5): ONE block (FIPS-197
    App. C.1) + key schedule; every advice column, all 25 lookups (20 of them over empty sets)."""
    import torch
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    v = golden["fips197"][0]
    key = bytes.fromhex(v["key"])
    pts = np.frombuffer(bytes.fromhex(v["pt"]), np.uint8).reshape(1, 16).copy()
    cir = Circuit(gpu, ffi.KIND_AES, 20, 5)
    advice, cts = cir.witness(key, torch.from_numpy(pts).cuda())
    ref = orc.Circuit(orc.KIND_AES, 20, 5).synthesize(key, pts)
    assert cts.cpu().numpy()[0].tobytes().hex() == v["ct"] == ref.ciphertexts()[0].tobytes().hex()
    for col in range(16):
        assert _same_on_device(advice[col], ref.advice(col)), col
    perm = cir.lookup_permuted(theta, want_inputs=False)
    Z = cir.lookup_product(beta, gamma)
    cir.lookup_check()
    _compare_all_lookups(cir, ref, perm, Z, theta, beta, gamma, list(range(25)))
    cir.close()


@pytest.mark.parametrize("k,N", [(20, 2), (21, 1)])
def test_1024_blocks_all_lookups(gpu, orc, ffi, challenges, k, N):
    """BASELINE config 3 (1024 blocks under one key in one circuit), both shapes of SURVEY 8(d):
    the reference-shaped K=20, N=2 (769 blocks in set 0, 255 in set 1) and K=21, N=1; every advice
    column and ALL lookups column for column."""
    import torch
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    pts = splitmix_bytes(SEED + 6, 16 * 1024).reshape(-1, 16)
    key = splitmix_bytes(SEED + 7, 16).tobytes()
    cir = Circuit(gpu, ffi.KIND_AES, k, N)
    advice, cts = cir.witness(key, torch.from_numpy(pts).cuda())
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    assert (cts.cpu().numpy() == ref.ciphertexts()).all()
    for col in range(3 * N + 1):
        assert _same_on_device(advice[col], ref.advice(col)), col
    perm = cir.lookup_permuted(theta, want_inputs=False)
    Z = cir.lookup_product(beta, gamma)
    cir.lookup_check()
    _compare_all_lookups(cir, ref, perm, Z, theta, beta, gamma, list(range(5 * N)))
    cir.close()


@pytest.mark.parametrize("tiled", [False, True])
def test_k22_row_sharded_against_the_oracle(gpu, orc, ffi, challenges, tiled):
    """BASELINE config 5's shape, oracle-compared: a K=22, N=1 circuit at capacity (3082 blocks)
    through the row-sharded entry points (4 shards run one after another on this GPU, the two
    exchanges done by hand): every advice column and the xor and sbox lookups."""
    import torch
    from halo2_aes_b200.backend import Circuit
    theta, beta, gamma = challenges
    k, N, G = 22, 1, 4
    nb = ffi.layout_query(ffi.KIND_AES, k, N)["capacity_blocks"]
    assert nb == 3082
    pts = splitmix_bytes(SEED + 60, 16 * nb).reshape(-1, 16)
    key = splitmix_bytes(SEED + 61, 16).tobytes()
    d_pts = torch.from_numpy(pts).cuda()
    shards = [Circuit(gpu, ffi.KIND_AES, k, N) for _ in range(G)]
    for g, c in enumerate(shards):
        c.set_shard(g, G, tiled=tiled)
    gidx = [torch.from_numpy(c.global_rows()).cuda() for c in shards]

    def glue(parts):   # [cols, local rows, 4] per shard -> [cols, n, 4] on the device
        out = torch.empty((parts[0].shape[0], 1 << k, 4), dtype=torch.int64, device="cuda")
        for ix, p in zip(gidx, parts):
            out[:, ix] = p
        return out

    adv = glue([c.witness(key, d_pts)[0] for c in shards])
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    for col in range(4):
        assert _same_on_device(adv[col], ref.advice(col)), col
    del adv
    hists = [c.shard_hist(theta) for c in shards]
    total = torch.stack(hists).sum(dim=0, dtype=torch.int32)
    for h in hists:
        h.copy_(total)
    perm = [c.shard_permuted(want_inputs=False) for c in shards]
    zp = [c.shard_partial(beta, gamma) for c in shards]
    gathered = torch.stack([p.clone() for _, p in zp]).contiguous()
    Zs = [c.shard_product(gathered, z) for c, (z, _) in zip(shards, zp)]
    for c in shards:
        c.lookup_check()
    got = _oracle_lookups(ref, [1, 2], theta, beta, gamma, workers=2)
    for li in (1, 2):
        for nm, parts in (("Ap", [p["Ap"] for p in perm]), ("Sp", [p["Sp"] for p in perm]), ("Z", Zs)):
            col = glue([t[li:li + 1] for t in parts])[0]
            assert _same_on_device(col, got[li][nm]), (nm, li)
    for c in shards:
        c.close()


def test_encrypt_append_equals_one_shot(gpu, orc, ffi):
    """h2aes_encrypt_append_dev (the reference's per-call `encrypt`, aes128.rs:154): schedule_key
    alone, then blocks appended in ragged batches across the column-set switch, gives the same
    advice columns and ciphertexts as one h2aes_witness_dev call; overflow is the capacity error
    and leaves the circuit untouched; append before schedule_key is KeysNotScheduled."""
    import torch
    from halo2_aes_b200.backend import Circuit
    k, N = 17, 2
    cap = ffi.layout_query(ffi.KIND_AES, k, N)["capacity_blocks"]
    pts = splitmix_bytes(SEED + 97, 16 * cap).reshape(-1, 16)
    key = splitmix_bytes(SEED + 98, 16).tobytes()
    d_pts = torch.from_numpy(pts).cuda()
    one = Circuit(gpu, ffi.KIND_AES, k, N)
    adv_ref, cts_ref = one.witness(key, d_pts)
    inc = Circuit(gpu, ffi.KIND_AES, k, N)
    with pytest.raises(ffi.KeysNotScheduled):
        inc.encrypt_append(d_pts[:1], adv_ref)
    adv, _ = inc.witness(key, None)                      # schedule_key alone
    got, at = [], 0
    for nb in (1, 7, 90, 40, cap):                       # 95 blocks fill set 0, the rest set 1
        nb = min(nb, cap - at)
        first, cts = inc.encrypt_append(d_pts[at:at + nb].contiguous(), adv)
        assert first == at
        got.append(cts)
        at += nb
    assert at == cap
    with pytest.raises(ffi.CapacityError):
        inc.encrypt_append(d_pts[:1].contiguous(), adv)
    assert torch.equal(adv, adv_ref) and torch.equal(torch.cat(got), cts_ref)
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
    assert (_host(adv) == ref.advice()).all()
    one.close()
    inc.close()


@pytest.mark.parametrize("case", ["small", "wide", "beta_zero", "gamma_cancels_pad"])
def test_challenge_sweep(gpu, orc, ffi, case):
    """theta / beta / gamma away from the seeded defaults: tiny challenges, full-width ones, and
    the two ways a denominator of the grand product can vanish — (A'+beta) with A' = 0 (beta = 0)
    and (S'+gamma) with S' = the padding value — which halo2's batch_invert turns into a zero and
    this library reports as H2AES_ERR_ZERO_DENOM instead of producing a column."""
    import torch
    from halo2_aes_b200.backend import Circuit
    rng = np.random.default_rng(0xC0FFEE)
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    big = lambda: orc.fr_from_int(int.from_bytes(rng.bytes(40), "little") % R)
    pts = splitmix_bytes(SEED + 95, 16 * 20).reshape(-1, 16)
    key = splitmix_bytes(SEED + 96, 16).tobytes()
    if case == "small":
        ch = (orc.fr_from_int(3), orc.fr_from_int(1), orc.fr_from_int(2))
    elif case == "wide":
        ch = (big(), big(), big())
    elif case == "beta_zero":
        ch = (big(), orc.fr_from_int(0), big())
    else:  # the 2-column range lookup pads its table with (1, 0): value theta -> gamma = -theta
        th = big()
        ch = (th, big(), orc.fr_from_int(R - orc.fr_to_int(th)))
    if case in ("small", "wide"):
        _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 1, key, pts, ch)
        return
    cir = Circuit(gpu, ffi.KIND_AES, 17, 1)
    cir.witness(key, torch.from_numpy(pts).cuda())
    cir.lookup_permuted(ch[0], want_inputs=False)
    cir.lookup_product(ch[1], ch[2], first=0, count=1)
    with pytest.raises(ffi.H2AesError) as ei:
        cir.lookup_check()
    assert ei.value.code == ffi.ERR_ZERO_DENOM
    cir.close()


@pytest.mark.parametrize("theta_int", [1, 2, 255])
def test_degenerate_theta_merges_equal_values(gpu, orc, ffi, theta_int):
    """With a tiny theta many different table tuples compress to the SAME field element
    (theta = 1: tag+x+y+z).  halo2 works on values, not on table rows, so bins must be merged by
    value and an input may match several table rows — the collision-heavy edge case of the
    permuted columns."""
    theta = orc.fr_from_int(theta_int)
    beta, gamma = orc.fr_from_int(5), orc.fr_from_int(7)
    pts = splitmix_bytes(SEED + 90, 16 * 30).reshape(-1, 16)
    _run_case(gpu, orc, ffi, ffi.KIND_AES, 17, 1, splitmix_bytes(SEED + 91, 16).tobytes(), pts,
              (theta, beta, gamma))


@pytest.mark.parametrize("kind,k,N,nb", [(1, 17, 1, 0), (0, 17, 1, 1), (0, 17, 2, 120), (0, 18, 3, 400)])
def test_permutation_argument(gpu, orc, ffi, challenges, kind, k, N, nb):
    """SURVEY 8(f).1 — the copy-constraint argument: sigma columns (keygen) and the chained grand
    products Z_perm against the oracle's restatement of halo2's Assembly / Argument::commit, plus
    the closing property (last set's Z[n-6] == 1 because every copy constraint holds)."""
    import torch
    from halo2_aes_b200.backend import Circuit
    _, beta, gamma = challenges
    key = splitmix_bytes(SEED + 95, 16).tobytes()
    pts = splitmix_bytes(SEED + 96, 16 * max(nb, 1)).reshape(-1, 16)[:nb]
    cir = Circuit(gpu, kind, k, N)
    assert cir.perm_shape() == orc.perm_shape(kind, k, N)
    advice, _ = cir.witness(key, torch.from_numpy(pts.copy()).cuda() if nb else None)
    sigma = cir.perm_sigma()
    Zp = cir.perm_product(beta, gamma, advice)
    gpu.synchronize()
    ref = orc.Circuit(kind, k, N).synthesize(key, pts if nb else None)
    r = ref.permutation(beta, gamma, want=("sigma", "Z"))
    assert (_host(sigma) == r["sigma"]).all(), "sigma"
    gz = _host(Zp)
    n = 1 << k
    one = orc.fr_from_int(1)
    assert (gz[0, 0] == one).all() and (gz[-1, n - 6] == one).all()
    for s in range(gz.shape[0]):
        bad = np.flatnonzero((gz[s] != r["Z"][s]).any(axis=1))
        assert bad.size == 0, f"Z_perm set {s}: {bad.size} rows differ, first {bad[:5]}"
    # the lookup product still works after the permutation product reused its challenge slots
    theta = challenges[0]
    cir.lookup_permuted(theta, first=1, count=1, want_inputs=False)
    Z = cir.lookup_product(beta, gamma, first=1, count=1)
    cir.lookup_check()
    assert (_host(Z)[0] == ref.lookup(1, theta, beta, gamma, want=("Z",))["Z"]).all()
    cir.close()
