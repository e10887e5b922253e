"""ctypes binding of the C ABI in include/h2aes.h (halo2_aes_b200/lib/libh2aes.so).

This is the same surface a Rust `extern "C"` block would bind.  The library is
built in-tree by `__graft_entry__.build()` / `make -C halo2_aes_b200/csrc`; there is no
fallback of any kind: a missing library raises at import of the first symbol.
"""
from __future__ import annotations

import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
# H2AES_LIB selects another build of the same library (tuning experiments); never a fallback
LIB_PATH = os.environ.get("H2AES_LIB") or os.path.join(_DIR, "lib", "libh2aes.so")

OK = 0
ERR_ARG, ERR_CAPACITY, ERR_ROWS, ERR_NOT_SCHEDULED, ERR_LOOKUP = 1, 2, 3, 4, 5
ERR_CUDA, ERR_NCCL, ERR_NOMEM, ERR_STATE, ERR_ZERO_DENOM = 6, 7, 8, 9, 10

KIND_AES, KIND_KEYSCHED = 0, 1
SBOX_REFERENCE, SBOX_FIPS197 = 0, 1
OP_NONE, OP_RANGE, OP_XOR, OP_SBOX, OP_MUL2, OP_MUL3 = 0, 1, 2, 3, 4, 5
DOMAIN_FFT, DOMAIN_LAGRANGE_TO_COEFF, DOMAIN_COEFF_TO_EXTENDED, DOMAIN_EXTENDED_TO_COEFF = 0, 1, 2, 3
COMM_ID_BYTES = 128  # an ncclUniqueId
SHARD_TILE_ROWS = 43520  # rows of an ownership tile of h2aes_circuit_set_shard_tiled
CS_DEGREE = 5  # cs.degree() of both circuits (lookups: 2 + 2 + 1)


class Layout(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("k", C.c_int32), ("n_sets", C.c_int32), ("n_advice", C.c_int32),
        ("n_fixed", C.c_int32), ("n_selectors", C.c_int32), ("n_lookups", C.c_int32),
        ("reserved", C.c_int32),
        ("n_rows", C.c_uint64), ("usable_rows", C.c_uint64), ("table_rows", C.c_uint64),
        ("key_schedule_rows", C.c_uint64), ("rows_per_block", C.c_uint64),
        ("capacity_set0", C.c_uint64), ("capacity_set", C.c_uint64),
        ("capacity_blocks", C.c_uint64), ("copies_key_schedule", C.c_uint64),
        ("copies_per_block", C.c_uint64),
    ]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != "reserved"}


class H2AesError(RuntimeError):
    """Non-zero status from the C ABI.  `.code` is the H2AES_ERR_* value."""

    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = lib().h2aes_status_string(code).decode()
        super().__init__(f"h2aes status {code}: {msg}" + (f" ({detail})" if detail else ""))


class CapacityError(H2AesError):
    """The reference's panic!("AES calls too many. doesn't fit in the rows") (aes128.rs:160-162)."""


class KeysNotScheduled(H2AesError):
    """The reference's expect("Keys should be scheduled") (aes128.rs:170)."""


class NotEnoughRows(H2AesError):
    """halo2 Error::NotEnoughRowsAvailable."""


_ERR_CLASSES = {ERR_CAPACITY: CapacityError, ERR_NOT_SCHEDULED: KeysNotScheduled,
                ERR_ROWS: NotEnoughRows}

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); exactly the declarations of include/h2aes.h
SIGNATURES = {
    "h2aes_abi_version": (C.c_int32, []),
    "h2aes_status_string": (C.c_char_p, [C.c_int32]),
    "h2aes_layout_query": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(Layout)]),
    "h2aes_block_position": (C.c_int32, [C.c_int32, C.c_int32, C.c_uint64,
                                          C.POINTER(C.c_int32), _u64p]),
    "h2aes_row_op": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32,
                                  C.c_uint64]),
    "h2aes_active_rows": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, _u64p]),
    "h2aes_copy_count": (C.c_int32, [C.c_int32, C.c_uint64, _u64p]),
    "h2aes_copy_constraints_host": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                                 C.c_uint64, C.c_uint64, _u64p]),
    "h2aes_perm_query": (C.c_int32, [C.c_int32] * 3 + [C.POINTER(C.c_int32)] * 5),
    "h2aes_perm_mapping_host": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _u32p]),
    "h2aes_ctx_create": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "h2aes_ctx_destroy": (C.c_int32, [_vp]),
    "h2aes_ctx_set_stream": (C.c_int32, [_vp, _vp]),
    "h2aes_ctx_synchronize": (C.c_int32, [_vp]),
    "h2aes_ctx_last_error": (C.c_char_p, [_vp]),
    "h2aes_ctx_kernel_launches": (C.c_uint64, [_vp]),
    "h2aes_ctx_profile": (C.c_int32, [_vp, C.c_int32]),
    "h2aes_ctx_profile_read": (C.c_int32, [_vp, C.c_char_p, C.c_uint64]),
    "h2aes_host_alloc": (C.c_int32, [C.c_uint64, C.POINTER(_vp)]),
    "h2aes_host_free": (C.c_int32, [_vp]),
    "h2aes_circuit_create": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "h2aes_circuit_destroy": (C.c_int32, [_vp]),
    "h2aes_witness_dev": (C.c_int32, [_vp, _u8p, _vp, C.c_uint64, _vp, _vp]),
    "h2aes_witness_host": (C.c_int32, [_vp, _u8p, _vp, C.c_uint64, _vp, _vp]),
    "h2aes_encrypt_append_dev": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, _vp, _u64p]),
    "h2aes_chip_rows_dev": (C.c_int32, [_vp, C.c_int32, _vp, _vp, C.c_uint64, _vp, _vp]),
    "h2aes_round_keys": (C.c_int32, [_vp, _u8p]),
    "h2aes_fixed_columns_dev": (C.c_int32, [_vp, C.c_uint64, _vp, _vp]),
    "h2aes_fixed_columns_host": (C.c_int32, [_vp, C.c_uint64, _vp, _vp]),
    "h2aes_copy_constraints_dev": (C.c_int32, [_vp, C.c_uint64, _vp]),
    "h2aes_lookup_permuted_dev": (C.c_int32, [_vp, _u64p, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "h2aes_lookup_permuted_host": (C.c_int32, [_vp, _u64p, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "h2aes_lookup_product_dev": (C.c_int32, [_vp, _u64p, _u64p, C.c_int32, C.c_int32, _vp]),
    "h2aes_lookup_product_host": (C.c_int32, [_vp, _u64p, _u64p, C.c_int32, C.c_int32, _vp]),
    "h2aes_perm_sigma_dev": (C.c_int32, [_vp, C.c_uint64, _vp]),
    "h2aes_perm_product_dev": (C.c_int32, [_vp, _u64p, _u64p, _vp, _vp]),
    "h2aes_domain_query": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), _u64p]),
    "h2aes_domain_create": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "h2aes_domain_destroy": (C.c_int32, [_vp]),
    "h2aes_fft_dev": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32]),
    "h2aes_lagrange_to_coeff_dev": (C.c_int32, [_vp, _vp, C.c_int32]),
    "h2aes_coeff_to_extended_dev": (C.c_int32, [_vp, _vp, C.c_int32, _vp]),
    "h2aes_extended_to_coeff_dev": (C.c_int32, [_vp, _vp, C.c_int32]),
    "h2aes_domain_transform_host": (C.c_int32, [_vp, C.c_int32, _vp, C.c_int32, _vp]),
    "h2aes_lookup_quotient_dev": (C.c_int32, [_vp, _vp, _u64p, _u64p, _u64p, _u64p, C.c_int32, C.c_int32,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "h2aes_gate_perm_quotient_dev": (C.c_int32, [_vp, _vp, _u64p, _u64p, _u64p, _vp, _vp, _vp, _vp, _vp, _vp]),
    "h2aes_divide_by_vanishing_dev": (C.c_int32, [_vp, _vp, C.c_int32]),
    "h2aes_g1_synthetic_bases_dev": (C.c_int32, [_vp, _u64p, C.c_uint64, _vp]),
    "h2aes_msm_create": (C.c_int32, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "h2aes_msm_destroy": (C.c_int32, [_vp]),
    "h2aes_msm_enable_runs": (C.c_int32, [_vp]),
    "h2aes_msm_dev": (C.c_int32, [_vp, _vp, C.c_int32, _vp]),
    "h2aes_msm_host": (C.c_int32, [_vp, _vp, C.c_int32, _vp]),
    "h2aes_lookup_stats": (C.c_int32, [_vp, _u64p]),
    "h2aes_lookup_check": (C.c_int32, [_vp]),
    "h2aes_circuit_set_shard": (C.c_int32, [_vp, C.c_int32, C.c_int32]),
    "h2aes_circuit_set_shard_tiled": (C.c_int32, [_vp, C.c_int32, C.c_int32]),
    "h2aes_circuit_shard_query": (C.c_int32, [_vp, _u64p, _u64p]),
    "h2aes_lookup_hist_dev": (C.c_int32, [_vp, _u64p, C.POINTER(_vp), _u64p]),
    "h2aes_lookup_permuted_shard_dev": (C.c_int32, [_vp, _vp, _vp, _vp, _vp]),
    "h2aes_lookup_partial_dev": (C.c_int32, [_vp, _u64p, _u64p, _vp, C.POINTER(_vp)]),
    "h2aes_lookup_product_shard_dev": (C.c_int32, [_vp, _vp, _vp]),
    "h2aes_comm_unique_id": (C.c_int32, [_u8p]),
    "h2aes_comm_init": (C.c_int32, [_vp, _u8p, C.c_int32, C.c_int32]),
    "h2aes_comm_destroy": (C.c_int32, [_vp]),
    "h2aes_lookup_columns_sharded_dev": (C.c_int32, [_vp, _u64p, _u64p, _u64p, _vp, _vp, _vp]),
    "h2aes_dbg_fr_mul": (C.c_int32, [_u64p, _u64p, _u64p]),
    "h2aes_dbg_fr_add": (C.c_int32, [_u64p, _u64p, _u64p]),
    "h2aes_dbg_fr_sub": (C.c_int32, [_u64p, _u64p, _u64p]),
    "h2aes_dbg_fr_inv": (C.c_int32, [_u64p, _u64p]),
    "h2aes_dbg_fr_inv_fermat": (C.c_int32, [_u64p, _u64p]),
    "h2aes_dbg_fr_from_u64": (C.c_int32, [C.c_uint64, _u64p]),
    "h2aes_dbg_fr_to_canonical": (C.c_int32, [_u64p, _u64p]),
    "h2aes_dbg_fr_mul_dev": (C.c_int32, [_vp, _u64p, _u64p, C.c_uint64, _u64p]),
    "h2aes_dbg_tables": (C.c_int32, [C.c_int32, _u8p, _u8p, _u8p]),
    "h2aes_dbg_block_trace": (C.c_int32, [C.c_int32, _u8p, _u8p, _u32p]),
    "h2aes_dbg_key_schedule_trace": (C.c_int32, [C.c_int32, _u8p, _u8p, _u32p, _u8p]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C halo2_aes_b200/csrc`. There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.h2aes_abi_version() != 1:
            raise ImportError("libh2aes.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, ctx=None):
    if rc == OK:
        return
    detail = ""
    if ctx:
        detail = (lib().h2aes_ctx_last_error(ctx) or b"").decode()
    raise _ERR_CLASSES.get(rc, H2AesError)(rc, detail)


def layout_query(kind: int, k: int, N: int) -> dict:
    out = Layout()
    check(lib().h2aes_layout_query(kind, k, N, C.byref(out)))
    return out.asdict()


def domain_query(k: int, j: int = CS_DEGREE):
    """(extended_k, quotient_len) of EvaluationDomain::new(j, k)"""
    ek = C.c_int32()
    q = C.c_uint64()
    check(lib().h2aes_domain_query(j, k, C.byref(ek), C.byref(q)))
    return ek.value, q.value


def block_position(k: int, N: int, block: int):
    s = C.c_int32()
    r = C.c_uint64()
    check(lib().h2aes_block_position(k, N, block, C.byref(s), C.byref(r)))
    return s.value, r.value
