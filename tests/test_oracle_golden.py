"""Pins the CPU oracle (oracle/) against every golden vector the reference holds for this path
 before anything is compared with it.  CPU only."""
import numpy as np
import pytest

from conftest import SEED, splitmix_bytes

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_tables_match_reference_constants(orc, golden):
    s, m2, m3 = orc.tables(fips=False)
    assert s.tolist() == golden["S_BOX"]          # src/constant.rs:1-14 (entry 255 == 23)
    assert m2.tolist() == golden["MUL_BY_2"]      # src/constant.rs:16-30
    assert m3.tolist() == golden["MUL_BY_3"]      # src/constant.rs:32-47
    sf, _, _ = orc.tables(fips=True)
    assert sf[255] == 0x16 and sf[:255].tolist() == golden["S_BOX"][:255]


def test_fr_encoding_constants(orc):
    # bn256 Fr(1) in Montgomery form: limbs of R mod r
    one = np.zeros(4, np.uint64)
    orc.lib().orc_fr_from_u64(1, orc._u64(one))
    assert [hex(int(x)) for x in one] == ["0xac96341c4ffffffb", "0x36fc76959f60cd29",
                                          "0x666ea36f7879462e", "0xe0a77c19a07df2f"]
    for v in (0, 1, 2, 255, 256, 0xFFFFFFFFFFFFFFFF):
        out = np.zeros(4, np.uint64)
        orc.lib().orc_fr_from_u64(v, orc._u64(out))
        assert orc.fr_to_int(out) == v
        assert (out == orc.fr_from_int(v)).all()


def test_fr_arithmetic_against_bigints(orc):
    rng = np.random.default_rng(SEED)
    for _ in range(200):
        a = int.from_bytes(rng.bytes(32), "little") % R_MOD
        b = int.from_bytes(rng.bytes(32), "little") % R_MOD
        A, B = orc.fr_from_int(a), orc.fr_from_int(b)
        out = np.zeros(4, np.uint64)
        orc.lib().orc_fr_mul(orc._u64(A), orc._u64(B), orc._u64(out))
        assert orc.fr_to_int(out) == a * b % R_MOD
        orc.lib().orc_fr_add(orc._u64(A), orc._u64(B), orc._u64(out))
        assert orc.fr_to_int(out) == (a + b) % R_MOD
        if a:
            orc.lib().orc_fr_inv(orc._u64(A), orc._u64(out))
            assert orc.fr_to_int(out) == pow(a, -1, R_MOD)


def test_xor_bytes_vector(orc, golden):
    v = golden["xor_bytes"]                        # src/utils.rs:40-47
    out = np.zeros(4, np.uint64)
    rc = orc.lib().orc_xor_bytes(orc._u64(orc.fr_from_int(v["x"])), orc._u64(orc.fr_from_int(v["y"])),
                                 orc._u64(out))
    assert rc == 0 and orc.fr_to_int(out) == v["z"]


def test_zero_key_expansion_matches_EXPANDED(orc, golden):
    c = orc.Circuit(orc.KIND_KEYSCHED, 17, 1).synthesize(bytes(16))
    rk = c.round_keys()
    words = ["".join(f"{b:02x}" for b in rk[4 * i:4 * i + 4]) for i in range(44)]
    assert words == golden["EXPANDED"]             # src/key_schedule.rs:337-345
    assert c.num_regions() == 421
    assert len(c.copies()) == 640


def test_key_schedule_circuit_satisfied_zero_key(orc):
    # src/key_schedule.rs:385-392 test_constraints: MockProver k=17, zero key
    c = orc.Circuit(orc.KIND_KEYSCHED, 17, 1).synthesize(bytes(16))
    rc, where = c.mock_verify()
    assert rc == 0, where


@pytest.mark.parametrize("vec", [0, 1, 2])
def test_fips197_blocks(orc, golden, vec):
    v = golden["fips197"][vec]
    c = orc.Circuit(orc.KIND_AES, 17, 1).synthesize(bytes.fromhex(v["key"]),
                                                    np.frombuffer(bytes.fromhex(v["pt"]), np.uint8))
    assert c.ciphertexts()[0].tobytes().hex() == v["ct"]
    assert c.num_regions() == 421 + 1345
    assert len(c.copies()) == 640 + 1952
    rc, where = c.mock_verify()
    assert rc == 0, where


def test_reference_unit_test_shape_satisfied(orc):
    # src/aes128.rs:409-418 test_correct_encryption: K=20, N=3, 1000 zero blocks under the zero
    # key must satisfy every constraint.  (Run at the reference's own shape.)
    c = orc.Circuit(orc.KIND_AES, 20, 3).synthesize(bytes(16), np.zeros((1000, 16), np.uint8))
    assert (c.ciphertexts() == np.frombuffer(
        bytes.fromhex("66e94bd4ef8a2c3b884cfa59ca342b2e"), np.uint8)).all()
    rc, where = c.mock_verify()
    assert rc == 0, where


def test_block_rows_and_capacity(orc, golden):
    # AES_ROWS = 1360 is exact; set 0 starts at row 400; capacity rule aes128.rs:303-325
    k, N = 17, 2
    cap0 = ((1 << k) - golden["KEY_SCHEDULE_ROWS"]) // golden["AES_ROWS"]
    cap = (1 << k) // golden["AES_ROWS"]
    pts = splitmix_bytes(SEED, 16 * (cap0 + cap)).reshape(-1, 16)
    c = orc.Circuit(orc.KIND_AES, k, N).synthesize(splitmix_bytes(SEED + 9, 16), pts)
    a0 = c.advice(0)
    for b in (0, 1, cap0 - 1):
        row = 400 + 1360 * b
        assert [orc.fr_to_int(a0[row + j]) for j in range(16)] == pts[b].tolist()
    a3 = c.advice(3)
    assert [orc.fr_to_int(a3[j]) for j in range(16)] == pts[cap0].tolist()
    rc, where = c.mock_verify()
    assert rc == 0, where
    with pytest.raises(orc.OracleError) as e:
        orc.Circuit(orc.KIND_AES, k, N).synthesize(bytes(16), np.zeros((cap0 + cap + 1, 16), np.uint8))
    assert e.value.code == 2
    with pytest.raises(orc.OracleError) as e:
        orc.Circuit(orc.KIND_AES, 16, 1).synthesize(bytes(16))
    assert e.value.code == 3


def test_random_blocks_hit_the_sbox_deviation(orc):
    # with the reference's table ~55% of random blocks differ from true AES
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    key = splitmix_bytes(SEED + 5, 16).tobytes()
    pts = splitmix_bytes(SEED + 6, 16 * 64).reshape(-1, 16)
    c = orc.Circuit(orc.KIND_AES, 17, 1).synthesize(key, pts)
    enc = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
    true = np.frombuffer(enc.update(pts.tobytes()), np.uint8).reshape(-1, 16)
    cf = orc.Circuit(orc.KIND_AES, 17, 1, fips_sbox=True).synthesize(key, pts)
    assert (cf.ciphertexts() == true).all()
    differ = int((c.ciphertexts() != true).any(axis=1).sum())
    assert 0 < differ < 64
    rc, where = c.mock_verify()
    assert rc == 0, where


def test_lookup_helper_invariants(orc, challenges):
    """No reference test pins A', S', Z (parity unpinned at that boundary): check the algebraic
    invariants halo2's verifier relies on."""
    theta, beta, gamma = challenges
    c = orc.Circuit(orc.KIND_KEYSCHED, 17, 1).synthesize(splitmix_bytes(SEED + 7, 16))
    n, u = c.n, c.n - 6
    one = orc.fr_from_int(1)
    for idx in range(3):
        r = c.lookup(idx, theta, beta, gamma)
        A, S, Ap, Sp, Z = (r[k] for k in ("A", "S", "Ap", "Sp", "Z"))
        assert (Z[0] == one).all() and (Z[u] == one).all()
        assert (Ap[0] == Sp[0]).all()
        same_as_S = (Ap[:u] == Sp[:u]).all(axis=1)
        same_as_prev = np.concatenate([[False], (Ap[1:u] == Ap[:u - 1]).all(axis=1)])
        assert (same_as_S | same_as_prev).all()
        def key(x):
            return np.sort(x[:u].view([("a", "u8"), ("b", "u8"), ("c", "u8"), ("d", "u8")]).ravel())
        assert (key(A) == key(Ap)).all() and (key(S) == key(Sp)).all()
        assert not Ap[u:].any() and not Sp[u:].any() and not Z[u + 1:].any()


def test_domain_constants_rederived_from_the_generator(orc):
    """halo2curves bn256::Fr [UPSTREAM]: MULTIPLICATIVE_GENERATOR = 7, S = 28; the oracle's
    ROOT_OF_UNITY / DELTA constants equal 7^((r-1)/2^28) and 7^(2^28), and the domain generator of
    a 2^k-row circuit has exact order 2^k."""
    assert (R_MOD - 1) % (1 << 28) == 0 and ((R_MOD - 1) >> 28) % 2 == 1
    for k in (17, 20, 22):
        om, de = orc.domain_constants(k)
        w, d = orc.fr_to_int(om), orc.fr_to_int(de)
        assert d == pow(7, 1 << 28, R_MOD)
        assert w == pow(pow(7, (R_MOD - 1) >> 28, R_MOD), 1 << (28 - k), R_MOD)
        assert pow(w, 1 << k, R_MOD) == 1 and pow(w, 1 << (k - 1), R_MOD) == R_MOD - 1


def test_permutation_argument_invariants(orc, challenges):
    """No reference test pins sigma / Z_perm (parity unpinned): check what halo2's verifier relies
    on — sigma is a permutation of the identity labels delta^p omega^i that only moves cells inside
    copy cycles of equal values, the sets chain (Z_s[0] == Z_{s-1}[n-6]) and the last one closes
    at 1 because every copy constraint holds."""
    _, beta, gamma = challenges
    k, N, nb = 17, 2, 110
    c = orc.Circuit(orc.KIND_AES, k, N).synthesize(
        splitmix_bytes(SEED + 30, 16), splitmix_bytes(SEED + 31, 16 * nb).reshape(-1, 16))
    ps = orc.perm_shape(orc.KIND_AES, k, N)
    assert ps["columns"][:5] == [("advice", 0), ("advice", 1), ("advice", 2), ("advice", 6), ("fixed", 4)]
    assert ps["n_columns"] == 3 * N + 2 and ps["chunk_len"] == 3 and ps["n_sets"] == 3
    r = c.permutation(beta, gamma)
    n = 1 << k
    m = r["mapping"].astype(np.int64)
    flat = m[:, :, 0] * n + m[:, :, 1]
    assert np.array_equal(np.sort(flat.ravel()), np.arange(ps["n_columns"] * n))   # a permutation
    cols = [c.advice(i) if t == "advice" else c.fixed(i) for t, i in ps["columns"]]
    vals = np.stack(cols)                                                     # [ncols, n, 4]
    moved = flat != np.arange(ps["n_columns"] * n).reshape(ps["n_columns"], n)
    src = vals.reshape(-1, 4)[flat[moved]]
    assert (vals[moved] == src).all()                      # cycles only join equal values
    one = orc.fr_from_int(1)
    Z = r["Z"]
    assert (Z[0, 0] == one).all() and (Z[-1, n - 6] == one).all()
    for s in range(1, Z.shape[0]):
        assert (Z[s, 0] == Z[s - 1, n - 6]).all()
    assert not Z[:, n - 5:].any()


def test_prover_side_restatements_against_first_principles_vectors(orc):
    """tests/golden/prover_side_vectors.json (pure python integers, make_prover_side_golden.py):
    the oracle's evaluation domain, G1 group law and bucket-method MSM reproduce them exactly."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "prover_side_vectors.json")) as f:
        g = json.load(f)
    R, Q = int(g["r"], 16), int(g["q"], 16)
    assert R == orc.R_MOD and Q == orc.Q_MOD
    dom = g["domain"]
    k = dom["k"]
    dq = orc.domain_query(k, dom["j"])
    assert dq["extended_k"] == dom["extended_k"]
    assert orc.fr_to_int(dq["omega"]) == int(dom["omega"], 16)
    assert orc.fr_to_int(dq["extended_omega"]) == int(dom["extended_omega"], 16)
    assert orc.fr_to_int(dq["g_coset"]) == int(g["zeta"], 16)
    _, delta = orc.domain_constants(k)
    assert orc.fr_to_int(delta) == int(g["delta"], 16)

    def col(name):
        return np.stack([orc.fr_from_int(int(v, 16)) for v in dom[name]])

    a = col("values")
    assert (orc.domain_transform(orc.FFT, k, a) == col("best_fft")).all()
    coeff = orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, a)
    assert (coeff == col("lagrange_to_coeff")).all()
    assert (orc.domain_transform(orc.COEFF_TO_EXTENDED, k, coeff, dom["j"])
            == col("coeff_to_extended_of_those_coefficients")).all()

    rq = (1 << 256) % Q

    def point(xy):
        return np.stack([orc.int_to_limbs(int(c, 16) * rq % Q) for c in xy])

    g1 = g["g1"]
    bases = orc.g1_test_bases(len(g1["bases"]), orc.fr_from_int(int(g1["d"], 16)))
    assert (bases == np.stack([point(p) for p in g1["bases"]])).all()
    scalars = np.stack([orc.fr_from_int(int(s_, 16)) for s_ in g1["scalars"]])
    assert (orc.g1_msm(scalars, bases) == point(g1["msm"])).all()
    assert (orc.g1_mul_generator(orc.fr_from_int(2)) == point(g1["two_g"])).all()
    assert (orc.g1_mul_generator(orc.fr_from_int(R - 1)) == point(g1["minus_g"])).all()
