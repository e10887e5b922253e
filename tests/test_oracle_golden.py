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
