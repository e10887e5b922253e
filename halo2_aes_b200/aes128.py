"""Host-side mirror of the reference's chip interface for the accelerated path.

Same names, argument order and error behaviour as the reference (paths in /root/reference):
    FixedAes128Config::{configure, schedule_key, encrypt}      src/aes128.rs:46,143,154
    Aes128KeyScheduleConfig::{configure, schedule_keys}        src/key_schedule.rs:40,80
    U8XorChip / SboxChip / MulBy2Chip / MulBy3Chip / U8RangeCheckChip ::configure   src/chips/*.rs
    load_enc_full_table                                         src/table.rs:18
so a `Circuit.synthesize` body reads like the reference's (aes128.rs:389-402,
key_schedule.rs:296-315).  What changes is what happens underneath: instead of one
`assign_region` per byte operation on one CPU thread, the calls are *recorded* by a
`B200Layouter` and the whole circuit is produced by the CUDA kernels behind the C ABI
(include/h2aes.h) when the layouter is flushed.  AssignedCell values are read back lazily.

The chips' per-cell assign methods (xor / substitute / mul / range_check) run fused inside
schedule_keys / encrypt / lcon in the witness kernel.  Called on their own they assign a one-row
region whose cells come from h2aes_chip_rows_dev (a row kernel); such a region is not placed in
one of the two circuit shapes (h2aes.h says why), it is returned as a `ChipRegion`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from ._ffi import CapacityError, KeysNotScheduled

KEY_SCHEDULE_ROWS = 1760  # src/constant.rs:113 (capacity accounting only; 400 rows are used)
AES_ROWS = 1360           # src/constant.rs:114
TAG_U8, TAG_XOR, TAG_SBOX, TAG_GFMUL2, TAG_GFMUL3 = 1, 2, 3, 4, 5  # src/table.rs:10-16


# ---------------------------------------------------------------------------------------------
# the slice of halo2's ConstraintSystem the reference's configure() calls
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Column:
    kind: str   # "advice" | "fixed" | "table" | "selector" | "complex_selector"
    index: int


@dataclass
class Lookup:
    name: str
    selector: Column
    tag: int
    inputs: tuple      # advice columns, in expression order after the tag
    tables: tuple      # table columns


class ConstraintSystem:
    """Records allocation order exactly like halo2's ConstraintSystem does for these calls, so
    column / selector / lookup indices equal the reference's."""

    def __init__(self):
        self.num_advice = self.num_fixed = self.num_selectors = 0
        self.selector_kinds: list[str] = []
        self.lookups: list[Lookup] = []
        self.gates: list[str] = []
        self.equality: list[Column] = []
        self.constants: list[Column] = []

    def advice_column(self) -> Column:
        self.num_advice += 1
        return Column("advice", self.num_advice - 1)

    def fixed_column(self) -> Column:
        self.num_fixed += 1
        return Column("fixed", self.num_fixed - 1)

    def lookup_table_column(self) -> Column:  # a TableColumn wraps a fixed column
        self.num_fixed += 1
        return Column("table", self.num_fixed - 1)

    def selector(self) -> Column:
        self.selector_kinds.append("simple")
        self.num_selectors += 1
        return Column("selector", self.num_selectors - 1)

    def complex_selector(self) -> Column:
        self.selector_kinds.append("complex")
        self.num_selectors += 1
        return Column("complex_selector", self.num_selectors - 1)

    def enable_equality(self, col: Column):
        if col not in self.equality:
            self.equality.append(col)

    def enable_constant(self, col: Column):
        self.constants.append(col)
        self.enable_equality(col)

    def lookup(self, name, selector, tag, inputs, tables) -> int:
        self.lookups.append(Lookup(name, selector, tag, tuple(inputs), tuple(tables)))
        return len(self.lookups) - 1

    def create_gate(self, name: str):
        self.gates.append(name)

    def shape(self):
        """(kind, N) of the two circuit shapes the kernels implement, else NotImplementedError."""
        n_complex = self.selector_kinds.count("complex")
        tags = [lk.tag for lk in self.lookups]
        if (self.num_advice == 4 and n_complex == 3 and tags == [TAG_U8, TAG_XOR, TAG_SBOX]
                and self.num_fixed == 5):
            return _ffi.KIND_KEYSCHED, 1
        if n_complex % 5 == 0 and n_complex and self.num_advice == 3 * (n_complex // 5) + 1 \
                and self.num_fixed == 5 \
                and tags == [TAG_U8, TAG_XOR, TAG_SBOX, TAG_GFMUL2, TAG_GFMUL3] * (n_complex // 5):
            return _ffi.KIND_AES, n_complex // 5
        raise NotImplementedError("constraint system is not one of the halo2-aes circuit shapes")


# ---------------------------------------------------------------------------------------------
# chips (configure only; src/chips/*.rs)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class U8RangeCheckConfig:
    x: Column
    q: Column
    lookup: int


@dataclass(frozen=True)
class U8XorConfig:
    x: Column
    y: Column
    z: Column
    q: Column
    lookup: int


@dataclass(frozen=True)
class SboxConfig:
    x: Column
    y: Column
    q: Column
    lookup: int


@dataclass(frozen=True)
class MulBy2Config(SboxConfig):
    pass


@dataclass(frozen=True)
class MulBy3Config(SboxConfig):
    pass


def _byte_of(cell) -> int:
    v = cell.value() if hasattr(cell, "value") else int(cell)
    if not 0 <= v < 256:
        raise ValueError("chip inputs are u8 cells")
    return v


@dataclass
class ChipRegion:
    """The one-row region a stand-alone chip call assigns: `cells` = the Montgomery limbs of the
    chip's columns in that row ([3, 4] int64 on the GPU), `out` = the byte of the returned cell."""
    tag: int
    cells: object
    out: int

    def value(self) -> int:
        return self.out


def _chip_row(layouter, tag, x, y=None) -> ChipRegion:
    import torch

    from .backend import Context
    xv, yv = _byte_of(x), (_byte_of(y) if y is not None else None)
    if layouter.ctx is None:
        layouter.ctx = Context(torch.cuda.current_device())
    dev = layouter.ctx.device
    xb = torch.tensor([xv], dtype=torch.uint8, device=dev)
    yb = torch.tensor([yv], dtype=torch.uint8, device=dev) if y is not None else None
    out, cells = layouter.ctx.chip_rows(tag, xb, yb)
    region = ChipRegion(tag, cells[:, 0], int(out.item()))
    layouter.chip_regions.append(region)
    return region


class U8RangeCheckChip:
    def __init__(self, config):
        self.config = config

    construct = classmethod(lambda cls, config: cls(config))

    @staticmethod
    def configure(meta, x, selector, tag_tab, x_tab):  # u8_range_check_chip.rs:28-49
        lk = meta.lookup("Check u8 range", selector, TAG_U8, (x,), (tag_tab, x_tab))
        return U8RangeCheckConfig(x, selector, lk)

    def range_check(self, layouter, x):  # u8_range_check_chip.rs:51-70
        return _chip_row(layouter, TAG_U8, x)


class U8XorChip:
    def __init__(self, config):
        self.config = config

    construct = classmethod(lambda cls, config: cls(config))

    @staticmethod
    def configure(meta, x, y, z, selector, tag_tab, x_tab, y_tab, z_tab):  # u8_xor_chip.rs:30-61
        lk = meta.lookup("Check correct XOR of u8", selector, TAG_XOR, (x, y, z),
                         (tag_tab, x_tab, y_tab, z_tab))
        return U8XorConfig(x, y, z, selector, lk)

    def xor(self, layouter, x, y):  # u8_xor_chip.rs:63-100
        return _chip_row(layouter, TAG_XOR, x, y)


class SboxChip:
    def __init__(self, config):
        self.config = config

    construct = classmethod(lambda cls, config: cls(config))

    @staticmethod
    def configure(meta, x, y, selector, tag_tab, x_tab, y_tab):  # sbox_chip.rs:29-55
        lk = meta.lookup("Check correct Sbox substitution", selector, TAG_SBOX, (x, y),
                         (tag_tab, x_tab, y_tab))
        return SboxConfig(x, y, selector, lk)

    def substitute(self, layouter, x):  # sbox_chip.rs:57-83
        return _chip_row(layouter, TAG_SBOX, x)


class MulBy2Chip:
    def __init__(self, config):
        self.config = config

    construct = classmethod(lambda cls, config: cls(config))

    @staticmethod
    def configure(meta, x, y, selector, tag_tab, x_tab, y_tab):  # gf_mul_chip.rs:31-57
        lk = meta.lookup("Check correct GF mul by 2", selector, TAG_GFMUL2, (x, y),
                         (tag_tab, x_tab, y_tab))
        return MulBy2Config(x, y, selector, lk)

    def mul(self, layouter, x):  # gf_mul_chip.rs:59-89
        return _chip_row(layouter, TAG_GFMUL2, x)


class MulBy3Chip:
    def __init__(self, config):
        self.config = config

    construct = classmethod(lambda cls, config: cls(config))

    @staticmethod
    def configure(meta, x, y, selector, tag_tab, x_tab, y_tab):
        lk = meta.lookup("Check correct GF mul by 3", selector, TAG_GFMUL3, (x, y),
                         (tag_tab, x_tab, y_tab))
        return MulBy3Config(x, y, selector, lk)

    def mul(self, layouter, x):
        return _chip_row(layouter, TAG_GFMUL3, x)


# ---------------------------------------------------------------------------------------------
# assigned cells and the recording layouter
# ---------------------------------------------------------------------------------------------
@dataclass
class AssignedCell:
    """(advice column, absolute row); the value is read from the GPU columns after the flush."""
    column: int
    row: int
    _owner: "B200Layouter" = field(repr=False, default=None)

    def value(self) -> int:
        """canonical integer value of the cell (flushes the layouter if needed)"""
        return self._owner.cell_value(self.column, self.row)


class B200Layouter:
    """What `synthesize` receives instead of halo2's SimpleFloorPlanner layouter.  Records
    load_enc_full_table / schedule_key / encrypt; `flush()` runs the CUDA kernels once for the
    whole circuit."""

    def __init__(self, k: int, ctx=None):
        self.k = k
        self.ctx = ctx
        self.table_loaded = False
        self.key = None
        self.plaintexts: list[bytes] = []
        self.circuit = None        # backend.Circuit after the first flush
        self.advice = None
        self.ciphertexts = None
        self._flushed = 0          # plaintexts already on the GPU
        self._shape = None
        self.chip_regions = []     # one-row regions of stand-alone chip calls (ChipRegion)

    def namespace(self, _name=None):  # Layouter::namespace: same layouter
        return self

    def _bind(self, kind, N):
        if self._shape is None:
            self._shape = (kind, N)
            self.layout = _ffi.layout_query(kind, self.k, N)  # raises NotEnoughRows for k < 17
        elif self._shape != (kind, N):
            raise ValueError("one layouter serves one circuit shape")

    def flush(self):
        """Produce every advice column on the GPU.  The first flush is schedule_key + the blocks
        recorded so far (h2aes_witness_dev); `encrypt` calls made after a flush are appended by the
        next one (h2aes_encrypt_append_dev), like the reference's per-call `encrypt`."""
        import torch

        from .backend import Circuit, Context
        if self.key is None:
            raise KeysNotScheduled(_ffi.ERR_NOT_SCHEDULED)
        pending = self.plaintexts[self._flushed:]
        pts = None
        if pending:
            pts = torch.from_numpy(
                np.frombuffer(b"".join(pending), np.uint8).reshape(-1, 16).copy()
            ).to(self.ctx.device if self.ctx is not None else "cuda")
        if self.circuit is None:
            kind, N = self._shape
            if self.ctx is None:
                self.ctx = Context(torch.cuda.current_device())
            self.circuit = Circuit(self.ctx, kind, self.k, N)
            self.advice, self.ciphertexts = self.circuit.witness(self.key, pts)
        elif pending:
            first, cts = self.circuit.encrypt_append(pts, self.advice)
            assert first == self._flushed
            self.ciphertexts = cts if self.ciphertexts is None else torch.cat([self.ciphertexts, cts])
        self._flushed = len(self.plaintexts)
        return self

    def cell_value(self, column: int, row: int) -> int:
        from .backend import to_host_u64
        self.flush()
        limbs = to_host_u64(self.advice[column, row])
        v = sum(int(limbs[i]) << (64 * i) for i in range(4))
        r = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
        return v * pow(1 << 256, -1, r) % r


def load_enc_full_table(layouter: B200Layouter, tables):
    """src/table.rs:18-192.  The table columns are input independent; they are produced by
    h2aes_fixed_columns_* when asked for (B200Prover.fixed)."""
    assert len(tables) == 4
    layouter.table_loaded = True


# ---------------------------------------------------------------------------------------------
# key schedule gadget
# ---------------------------------------------------------------------------------------------
class Aes128KeyScheduleConfig:
    def __init__(self, words_column, round_constants, q_eq_rcon, advices, u8_range_check_config,
                 u8_xor_config, sbox_config):
        self.words_column = words_column
        self.round_constants = round_constants
        self.q_eq_rcon = q_eq_rcon
        self._advices = advices
        self.u8_range_check_config = u8_range_check_config
        self.u8_xor_config = u8_xor_config
        self.sbox_config = sbox_config
        self._kind = _ffi.KIND_KEYSCHED
        self._N = 1

    @staticmethod
    def configure(meta, advices, u8_xor_config, sbox_config, u8_range_check_config):
        # key_schedule.rs:40-77 — same allocation order
        words_column = meta.advice_column()
        round_constants = meta.fixed_column()
        q_eq_rcon = meta.selector()
        for a in advices:
            meta.enable_equality(a)
        meta.enable_equality(words_column)
        meta.enable_constant(round_constants)
        meta.create_gate("Equality RC")
        return Aes128KeyScheduleConfig(words_column, round_constants, q_eq_rcon, tuple(advices),
                                       u8_range_check_config, u8_xor_config, sbox_config)

    def schedule_keys(self, layouter: B200Layouter, key) -> list[list[AssignedCell]]:
        """key_schedule.rs:80-96: returns the 11x16 cells holding the round-key bytes."""
        key = bytes(key)
        assert len(key) == 16
        layouter._bind(self._kind, self._N)
        layouter.key = key
        w = self.words_column.index
        a2 = self._advices[2].index
        words = [[AssignedCell(w, i, layouter) for i in range(16)]]
        for r in range(1, 11):
            words.append([AssignedCell(a2, 40 * (r - 1) + 8 + i, layouter) for i in range(16)])
        return words


# ---------------------------------------------------------------------------------------------
# AES-128 gadget
# ---------------------------------------------------------------------------------------------
class FixedAes128Config:
    """FixedAes128Config<K, N> (src/aes128.rs:27-44); K and N are run-time values here."""

    def __init__(self, K, N, key_schedule_config, configs, advices, tables):
        self.K, self.N = K, N
        self.keys = None
        self.key_schedule_config = key_schedule_config
        self.configs = configs
        self.advices = advices
        self.tables = tables
        self.current = 0
        self.count = 0

    @staticmethod
    def configure(meta: ConstraintSystem, K: int, N: int) -> "FixedAes128Config":
        # aes128.rs:46-141 — same allocation order
        tables = tuple(meta.lookup_table_column() for _ in range(4))
        advices = tuple(tuple(meta.advice_column() for _ in range(3)) for _ in range(N))
        configs = ([], [], [], [], [])
        for i in range(N):
            q_range, q_xor, q_sbox, q_m2, q_m3 = (meta.complex_selector() for _ in range(5))
            a = advices[i]
            configs[0].append(U8RangeCheckChip.configure(meta, a[0], q_range, tables[0], tables[1]))
            configs[1].append(U8XorChip.configure(meta, a[0], a[1], a[2], q_xor, *tables))
            configs[2].append(SboxChip.configure(meta, a[0], a[1], q_sbox, *tables[:3]))
            configs[3].append(MulBy2Chip.configure(meta, a[0], a[1], q_m2, *tables[:3]))
            configs[4].append(MulBy3Chip.configure(meta, a[0], a[1], q_m3, *tables[:3]))
        ks = Aes128KeyScheduleConfig.configure(meta, advices[0], configs[1][0], configs[2][0],
                                               configs[0][0])
        ks._kind, ks._N = _ffi.KIND_AES, N
        for set_cols in advices:
            for c in set_cols:
                meta.enable_equality(c)
        return FixedAes128Config(K, N, ks, configs, advices, tables)

    def schedule_key(self, layouter: B200Layouter, key):
        self.keys = self.key_schedule_config.schedule_keys(layouter, key)

    def aes_callable(self) -> bool:  # aes128.rs:303-325
        max_row = 1 << self.K
        if self.current == 0:
            max_row -= KEY_SCHEDULE_ROWS
        if max_row >= self.count * AES_ROWS + AES_ROWS:
            return True
        if self.current < self.N - 1:
            self.current += 1
            self.count = 0
            return True
        return False

    def encrypt(self, layouter: B200Layouter, plaintext) -> list[AssignedCell]:
        """aes128.rs:154-265: returns the 16 ciphertext cells."""
        if not self.aes_callable():
            raise CapacityError(_ffi.ERR_CAPACITY)       # panic!("AES calls too many...")
        self.count += 1
        if self.keys is None:
            raise KeysNotScheduled(_ffi.ERR_NOT_SCHEDULED)  # expect("Keys should be scheduled")
        plaintext = bytes(plaintext)
        assert len(plaintext) == 16
        block = len(layouter.plaintexts)
        s, row = _ffi.block_position(self.K, self.N, block)
        assert s == self.current and row == (400 if s == 0 else 0) + AES_ROWS * (self.count - 1)
        layouter.plaintexts.append(plaintext)
        a2 = self.advices[s][2].index
        return [AssignedCell(a2, row + 1344 + i, layouter) for i in range(16)]


# ---------------------------------------------------------------------------------------------
# the caller side: run a circuit the way MockProver::run / create_proof would drive it
# ---------------------------------------------------------------------------------------------
class B200Prover:
    """configure + synthesize a circuit object exposing `configure(meta)` and
    `synthesize(config, layouter)`, then serve its columns from the GPU.  Mirrors how the
    reference's tests drive `MockProver::run(k, &circuit, vec![])`."""

    def __init__(self, k: int, circuit, ctx=None):
        self.k = k
        self.meta = ConstraintSystem()
        self.config = circuit.configure(self.meta)
        self.kind, self.N = self.meta.shape()
        self.layouter = B200Layouter(k, ctx)
        self.layouter._bind(self.kind, self.N)
        circuit.synthesize(self.config, self.layouter)
        if not self.layouter.table_loaded:
            raise ValueError("synthesize did not call load_enc_full_table")
        self.layouter.flush()
        self.circuit = self.layouter.circuit

    @classmethod
    def run(cls, k, circuit, ctx=None):
        return cls(k, circuit, ctx)

    def advice(self):
        return self.layouter.advice

    def ciphertexts(self):
        return self.layouter.ciphertexts

    def fixed_and_selectors(self):
        return self.circuit.fixed_columns()

    def copies(self):
        return self.circuit.copy_constraints()

    def permutation_sigma(self):
        """the sigma columns keygen derives from the copy constraints (pk.permutation)"""
        return self.circuit.perm_sigma()

    def permutation_product(self, beta, gamma):
        """the chained grand products Z_perm of permutation::Argument::commit"""
        return self.circuit.perm_product(beta, gamma, self.layouter.advice)

    def lookup_columns(self, theta, beta, gamma, want_inputs=True):
        out = self.circuit.lookup_permuted(theta, want_inputs=want_inputs)
        out["Z"] = self.circuit.lookup_product(beta, gamma)
        self.circuit.lookup_check()
        return out
