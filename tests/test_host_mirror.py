"""The host-side mirror of the reference's chip interface (halo2_aes_b200/aes128.py): circuits
written like the reference's own test circuits (src/aes128.rs:375-418,
src/key_schedule.rs:242-320, 385-392)."""
import numpy as np
import pytest

from conftest import SEED, splitmix_bytes


class TestAesCircuit:
    """src/aes128.rs:375-407 with K, N, encrypt_num as run-time values."""

    def __init__(self, key, plaintext, K=20, N=3, encrypt_num=1000):
        self.key, self.plaintext, self.K, self.N, self.encrypt_num = key, plaintext, K, N, encrypt_num

    def configure(self, meta):
        from halo2_aes_b200.aes128 import FixedAes128Config
        return FixedAes128Config.configure(meta, self.K, self.N)

    def synthesize(self, config, layouter):
        from halo2_aes_b200.aes128 import load_enc_full_table
        load_enc_full_table(layouter, config.tables)
        config.schedule_key(layouter, self.key)
        self.ct_cells = [config.encrypt(layouter, self.plaintext) for _ in range(self.encrypt_num)]


class TestKeyScheduleCircuit:
    """src/key_schedule.rs:247-320."""

    def __init__(self, key):
        self.key = key

    def configure(self, meta):
        from halo2_aes_b200 import aes128 as m
        advices = [meta.advice_column() for _ in range(3)]
        tables = [meta.lookup_table_column() for _ in range(4)]
        q_range, q_xor, q_sbox = (meta.complex_selector() for _ in range(3))
        rc = m.U8RangeCheckChip.configure(meta, advices[0], q_range, tables[0], tables[1])
        xc = m.U8XorChip.configure(meta, advices[0], advices[1], advices[2], q_xor, *tables)
        sc = m.SboxChip.configure(meta, advices[0], advices[1], q_sbox, *tables[:3])
        return m.Aes128KeyScheduleConfig.configure(meta, advices, xc, sc, rc), tables

    def synthesize(self, config, layouter):
        from halo2_aes_b200.aes128 import load_enc_full_table
        load_enc_full_table(layouter, config[1])
        self.words = config[0].schedule_keys(layouter.namespace("AES128 schedule key"), self.key)


TestAesCircuit.__test__ = False
TestKeyScheduleCircuit.__test__ = False


def test_configure_allocation_order(ffi):
    """Column / selector / lookup indices equal the reference's configure order."""
    from halo2_aes_b200 import aes128 as m
    meta = m.ConstraintSystem()
    cfg = m.FixedAes128Config.configure(meta, 20, 5)
    assert [t.index for t in cfg.tables] == [0, 1, 2, 3]
    assert [[c.index for c in s] for s in cfg.advices] == [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(5)]
    assert cfg.key_schedule_config.words_column.index == 15
    assert cfg.key_schedule_config.round_constants.index == 4
    assert cfg.key_schedule_config.q_eq_rcon.index == 25
    assert meta.selector_kinds == ["complex"] * 25 + ["simple"]
    assert [lk.tag for lk in meta.lookups] == [1, 2, 3, 4, 5] * 5
    assert [lk.selector.index for lk in meta.lookups] == list(range(25))
    assert [len(lk.inputs) + 1 for lk in meta.lookups[:5]] == [2, 4, 3, 3, 3]
    assert meta.gates == ["Equality RC"]
    assert meta.shape() == (ffi.KIND_AES, 5)
    lay = ffi.layout_query(ffi.KIND_AES, 20, 5)
    assert (meta.num_advice, meta.num_fixed, meta.num_selectors, len(meta.lookups)) == \
        (lay["n_advice"], lay["n_fixed"], lay["n_selectors"], lay["n_lookups"])
    meta2 = m.ConstraintSystem()
    TestKeyScheduleCircuit(bytes(16)).configure(meta2)
    assert meta2.shape() == (ffi.KIND_KEYSCHED, 1)
    assert [lk.tag for lk in meta2.lookups] == [1, 2, 3]


def test_capacity_rule_and_panics(ffi):
    """aes_callable (aes128.rs:303-325), the capacity panic (:160-162) and the
    keys-not-scheduled expect (:170) — no GPU needed: recording is lazy."""
    from halo2_aes_b200 import aes128 as m
    meta = m.ConstraintSystem()
    cfg = m.FixedAes128Config.configure(meta, 17, 2)
    lay = m.B200Layouter(17)
    with pytest.raises(ffi.KeysNotScheduled):
        cfg.encrypt(lay, bytes(16))
    cfg = m.FixedAes128Config.configure(m.ConstraintSystem(), 17, 2)
    lay = m.B200Layouter(17)
    cfg.schedule_key(lay, bytes(16))
    cells = None
    for i in range(95 + 96):
        cells = cfg.encrypt(lay, bytes(16))
    assert cfg.current == 1 and cfg.count == 96
    assert (cells[0].column, cells[0].row) == (5, 95 * 1360 + 1344)
    with pytest.raises(ffi.CapacityError):
        cfg.encrypt(lay, bytes(16))
    # benches/aes128.rs asks for 6000 blocks at K=20,N=5 -> panics at block 3854
    cfg = m.FixedAes128Config.configure(m.ConstraintSystem(), 20, 5)
    lay = m.B200Layouter(20)
    cfg.schedule_key(lay, bytes(16))
    for i in range(3853):
        cfg.encrypt(lay, bytes(16))
    with pytest.raises(ffi.CapacityError):
        cfg.encrypt(lay, bytes(16))
    with pytest.raises(ffi.NotEnoughRows):
        m.FixedAes128Config.configure(m.ConstraintSystem(), 16, 1).schedule_key(
            m.B200Layouter(16), bytes(16))
    with pytest.raises(ValueError):   # chip inputs are u8 cells (checked before any device call)
        m.U8XorChip.construct(cfg.configs[1][0]).xor(lay, 256, 0)


@pytest.mark.gpu
def test_correct_encryption_reference_shape(ffi, orc, challenges):
    """src/aes128.rs:409-418 test_correct_encryption at the reference's own shape (K=20, N=3,
    1000 zero blocks under the zero key) on the GPU, then the restated MockProver on the
    GPU-produced columns; plus the ciphertext check the reference left commented out
    (:420-434)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from halo2_aes_b200.aes128 import B200Prover
    from halo2_aes_b200.backend import to_host_u64
    circuit = TestAesCircuit(bytes(16), bytes(16))
    prover = B200Prover.run(20, circuit)
    ct = prover.ciphertexts().cpu().numpy()
    assert (ct == np.frombuffer(bytes.fromhex("66e94bd4ef8a2c3b884cfa59ca342b2e"), np.uint8)).all()
    assert [c.value() for c in circuit.ct_cells[999]] == list(ct[999])
    fixed, sel = prover.fixed_and_selectors()
    one = orc.fr_from_int(1)
    sel_bits = (to_host_u64(sel) == one).all(axis=-1).astype(np.uint8)
    rc, where = orc.mock_verify_ext(orc.KIND_AES, 20, 3, to_host_u64(prover.advice()),
                                    to_host_u64(fixed), sel_bits, to_host_u64(prover.copies()))
    assert rc == 0, (rc, where)
    theta, beta, gamma = challenges
    cols = prover.lookup_columns(theta, beta, gamma, want_inputs=False)
    Z = to_host_u64(cols["Z"])
    n = 1 << 20
    assert (Z[:, 0] == one).all() and (Z[:, n - 6] == one).all()
    del cols
    Zp = to_host_u64(prover.permutation_product(beta, gamma))   # 11 columns -> 4 chained sets
    assert Zp.shape[0] == 4 and (Zp[0, 0] == one).all() and (Zp[-1, n - 6] == one).all()
    assert all((Zp[s, 0] == Zp[s - 1, n - 6]).all() for s in range(1, 4))


@pytest.mark.gpu
def test_key_schedule_constraints_and_EXPANDED(ffi, orc, golden):
    """src/key_schedule.rs:385-392 test_constraints, and the #[ignore]d
    test_correct_key_scheduling (:348-383) with today's cell positions."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from halo2_aes_b200.aes128 import B200Prover
    from halo2_aes_b200.backend import to_host_u64
    circuit = TestKeyScheduleCircuit(bytes(16))
    prover = B200Prover.run(17, circuit)
    words = ["".join(f"{circuit.words[i // 4][4 * (i % 4) + j].value():02x}" for j in range(4))
             for i in range(44)]
    assert words == golden["EXPANDED"]
    fixed, sel = prover.fixed_and_selectors()
    one = orc.fr_from_int(1)
    sel_bits = (to_host_u64(sel) == one).all(axis=-1).astype(np.uint8)
    rc, where = orc.mock_verify_ext(orc.KIND_KEYSCHED, 17, 1, to_host_u64(prover.advice()),
                                    to_host_u64(fixed), sel_bits, to_host_u64(prover.copies()))
    assert rc == 0, (rc, where)
