"""Device-side handles: thin Python objects over the C ABI (include/h2aes.h).

torch is used only for device memory, streams and (in sharded.py) torch.distributed; every
column is produced by the library's own sm_100a kernels.  Field elements live in int64 tensors
of shape [..., 4] holding the 4 little-endian u64 Montgomery limbs of bn256 Fr, bit-compatible
with halo2curves' `Fr` (view them as numpy uint64 on the host).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi
from ._ffi import KIND_AES, KIND_KEYSCHED, SBOX_REFERENCE, check, lib


def _fr_arg(x):
    """4 u64 Montgomery limbs -> ctypes pointer (keeps the array alive through the return)."""
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64).reshape(4))
    return a, a.ctypes.data_as(C.POINTER(C.c_uint64))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Context:
    """One per (process, GPU): uploads the S-box and the byte -> Fr(Montgomery) table."""

    def __init__(self, device: int = 0, sbox_mode: int = SBOX_REFERENCE):
        if not torch.cuda.is_available():
            raise RuntimeError("halo2_aes_b200 needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        check(lib().h2aes_ctx_create(device, sbox_mode, C.byref(h)))
        self._h = h
        self.use_stream(torch.cuda.current_stream(self.device))

    def use_stream(self, stream: "torch.cuda.Stream"):
        self.stream = stream
        check(lib().h2aes_ctx_set_stream(self._h, C.c_void_p(stream.cuda_stream)), self._h)

    def synchronize(self):
        check(lib().h2aes_ctx_synchronize(self._h), self._h)

    @property
    def kernel_launches(self) -> int:
        return int(lib().h2aes_ctx_kernel_launches(self._h))

    def profile(self, enable: bool = True):
        """bracket every kernel of this context with CUDA events (h2aes_ctx_profile)"""
        check(lib().h2aes_ctx_profile(self._h, int(enable)), self._h)

    def profile_read(self) -> dict:
        """{kernel: (launches, total_ms)} since profiling was enabled / last read"""
        buf = C.create_string_buffer(1 << 16)
        check(lib().h2aes_ctx_profile_read(self._h, buf, len(buf)), self._h)
        out = {}
        for ln in buf.value.decode().splitlines():
            name, n, ms = ln.split()
            out[name] = (int(n), float(ms))
        return out

    def chip_rows(self, tag: int, x, y=None, want_cells=True):
        """h2aes_chip_rows_dev: n one-row regions of the chip with table tag `tag` (1 range, 2 xor,
        3 sbox, 4 mul2, 5 mul3) on uint8 CUDA tensors x (and y for xor).  Returns (out bytes [n],
        cells [3, n, 4] int64 or None)."""
        n = x.numel()
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
        if y is not None:
            assert y.is_cuda and y.dtype == torch.uint8 and y.is_contiguous() and y.numel() == n
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        cells = torch.empty((3, n, 4), dtype=torch.int64, device=self.device) if want_cells else None
        check(lib().h2aes_chip_rows_dev(self._h, tag, _ptr(x) if n else None,
                                        _ptr(y) if (y is not None and n) else None, n,
                                        _ptr(cells) if n else None, _ptr(out) if n else None), self._h)
        return out, cells

    # ---- NCCL inside the library (h2aes_comm_*) ------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """the 128-byte id rank 0 ships to every rank before comm_init"""
        buf = (C.c_uint8 * _ffi.COMM_ID_BYTES)()
        check(lib().h2aes_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """collective: every rank of the row-sharded circuit calls it with rank 0's id"""
        assert len(unique_id) == _ffi.COMM_ID_BYTES
        buf = (C.c_uint8 * _ffi.COMM_ID_BYTES).from_buffer_copy(unique_id)
        check(lib().h2aes_comm_init(self._h, buf, rank, world), self._h)

    def comm_destroy(self):
        check(lib().h2aes_comm_destroy(self._h), self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().h2aes_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Circuit:
    """One halo2 circuit instance of shape (kind, k, N) resident on the GPU."""

    def __init__(self, ctx: Context, kind: int, k: int, N: int = 1):
        self.ctx = ctx
        self.kind, self.k, self.N = kind, k, N
        self.layout = _ffi.layout_query(kind, k, N)
        self.n = self.layout["n_rows"]
        self.n_blocks = 0
        self.shard, self.n_shards, self.tiled = 0, 1, False
        self._rows, self._partial_words = self.n, self.layout["n_lookups"] * 4
        h = C.c_void_p()
        check(lib().h2aes_circuit_create(ctx._h, kind, k, N, C.byref(h)), ctx._h)
        self._h = h

    # ---- helpers ----------------------------------------------------------------------------
    @property
    def rows(self) -> int:
        """rows of every column held by this (shard of the) circuit"""
        return self._rows

    def global_rows(self) -> np.ndarray:
        """global row index of every local row of this shard (host; tests and checks)"""
        if self.n_shards == 1:
            return np.arange(self.n, dtype=np.int64)
        ts = _ffi.SHARD_TILE_ROWS if self.tiled else self.n // self.n_shards
        lr = np.arange(self._rows, dtype=np.int64)
        t = lr // ts
        return (t * self.n_shards + self.shard) * ts + (lr - t * ts)

    def _cols(self, ncols: int, out=None):
        if out is None:
            out = torch.empty((ncols, self.rows, 4), dtype=torch.int64, device=self.ctx.device)
        assert out.is_cuda and out.is_contiguous() and out.numel() == ncols * self.rows * 4
        return out

    def set_shard(self, shard: int, n_shards: int, tiled: bool = False):
        """shard `shard` of `n_shards`: contiguous rows, or (tiled) ownership tiles of 43520 rows
        dealt round-robin — the balanced split for the grand product"""
        fn = lib().h2aes_circuit_set_shard_tiled if tiled else lib().h2aes_circuit_set_shard
        check(fn(self._h, shard, n_shards), self.ctx._h)
        self.shard, self.n_shards, self.tiled = shard, n_shards, bool(tiled) and n_shards > 1
        rows, pw = C.c_uint64(), C.c_uint64()
        check(lib().h2aes_circuit_shard_query(self._h, C.byref(rows), C.byref(pw)), self.ctx._h)
        self._rows, self._partial_words = rows.value, pw.value

    # ---- witness ------------------------------------------------------------------------------
    def witness(self, key, plaintexts=None, out=None, want_ciphertexts=True):
        """schedule_key(key) then encrypt(pt) for every 16-byte row of `plaintexts`.

        plaintexts: uint8 CUDA tensor [B,16] (or None).  Returns (advice [n_advice, rows, 4],
        ciphertexts [B,16] uint8 or None).  Asynchronous on the context's stream.
        """
        key_a = np.ascontiguousarray(np.frombuffer(bytes(key), dtype=np.uint8))
        assert key_a.size == 16
        B = 0 if plaintexts is None else plaintexts.shape[0]
        if B:
            assert plaintexts.is_cuda and plaintexts.dtype == torch.uint8 and plaintexts.is_contiguous()
        advice = self._cols(self.layout["n_advice"], out)
        cts = torch.empty((B, 16), dtype=torch.uint8, device=self.ctx.device) \
            if (want_ciphertexts and B) else None
        check(lib().h2aes_witness_dev(self._h, key_a.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      _ptr(plaintexts) if B else None, B, _ptr(advice),
                                      _ptr(cts)), self.ctx._h)
        self.n_blocks = B
        return advice, cts

    def encrypt_append(self, plaintexts, advice, want_ciphertexts=True):
        """h2aes_encrypt_append_dev: `encrypt` for every row of `plaintexts` [B,16] after the blocks
        this circuit already holds; `advice` is the buffer `witness` returned.  Returns
        (index of the first appended block, ciphertexts [B,16] or None)."""
        B = plaintexts.shape[0]
        assert plaintexts.is_cuda and plaintexts.dtype == torch.uint8 and plaintexts.is_contiguous()
        cts = torch.empty((B, 16), dtype=torch.uint8, device=self.ctx.device) \
            if (want_ciphertexts and B) else None
        first = C.c_uint64()
        check(lib().h2aes_encrypt_append_dev(self._h, _ptr(plaintexts) if B else None, B, _ptr(advice),
                                             _ptr(cts), C.byref(first)), self.ctx._h)
        self.n_blocks += B
        return first.value, cts

    def witness_host(self, key, plaintexts: np.ndarray, advice_out: np.ndarray,
                     ciphertexts_out: np.ndarray | None = None):
        """Same through the _host entry point: numpy in, numpy out, copies done by the library."""
        key_a = np.ascontiguousarray(np.frombuffer(bytes(key), dtype=np.uint8))
        pts = np.ascontiguousarray(plaintexts, dtype=np.uint8).reshape(-1, 16)
        B = pts.shape[0]
        assert advice_out.dtype == np.uint64 and advice_out.flags.c_contiguous
        assert advice_out.size == self.layout["n_advice"] * self.rows * 4
        check(lib().h2aes_witness_host(
            self._h, key_a.ctypes.data_as(C.POINTER(C.c_uint8)),
            C.c_void_p(pts.ctypes.data) if B else None, B, C.c_void_p(advice_out.ctypes.data),
            C.c_void_p(ciphertexts_out.ctypes.data) if ciphertexts_out is not None else None),
            self.ctx._h)
        self.n_blocks = B
        return advice_out

    def round_keys(self) -> np.ndarray:
        out = np.zeros(176, np.uint8)
        check(lib().h2aes_round_keys(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8))), self.ctx._h)
        return out

    # ---- input-independent columns ------------------------------------------------------------
    def fixed_columns(self, n_blocks=None, want_fixed=True, want_selectors=True):
        nb = self.n_blocks if n_blocks is None else n_blocks
        fixed = self._cols(self.layout["n_fixed"]) if want_fixed else None
        sel = self._cols(self.layout["n_selectors"]) if want_selectors else None
        check(lib().h2aes_fixed_columns_dev(self._h, nb, _ptr(fixed), _ptr(sel)), self.ctx._h)
        return fixed, sel

    def copy_constraints(self, n_blocks=None):
        nb = self.n_blocks if n_blocks is None else n_blocks
        total = C.c_uint64()
        check(lib().h2aes_copy_count(self.kind, nb, C.byref(total)))
        out = torch.empty((total.value, 4), dtype=torch.int64, device=self.ctx.device)
        check(lib().h2aes_copy_constraints_dev(self._h, nb, _ptr(out)), self.ctx._h)
        return out

    # ---- lookup helper columns ------------------------------------------------------------------
    def lookup_permuted(self, theta, first=0, count=None, want_inputs=True, out=None):
        """A, S, A', S' of lookups [first, first+count): dict of [count, n, 4] tensors."""
        L = self.layout["n_lookups"]
        count = L - first if count is None else count
        names = (["A", "S"] if want_inputs else []) + ["Ap", "Sp"]
        bufs = out or {}
        for nm in names:
            if nm not in bufs:
                bufs[nm] = torch.empty((count, self.n, 4), dtype=torch.int64, device=self.ctx.device)
        th, thp = _fr_arg(theta)
        check(lib().h2aes_lookup_permuted_dev(self._h, thp, first, count, _ptr(bufs.get("A")),
                                              _ptr(bufs.get("S")), _ptr(bufs["Ap"]),
                                              _ptr(bufs["Sp"])), self.ctx._h)
        return bufs

    def lookup_product(self, beta, gamma, first=0, count=None, out=None):
        L = self.layout["n_lookups"]
        count = L - first if count is None else count
        Z = out if out is not None else torch.empty((count, self.n, 4), dtype=torch.int64,
                                                     device=self.ctx.device)
        be, bep = _fr_arg(beta)
        ga, gap = _fr_arg(gamma)
        check(lib().h2aes_lookup_product_dev(self._h, bep, gap, first, count, _ptr(Z)), self.ctx._h)
        return Z

    # ---- permutation argument -----------------------------------------------------------------
    def perm_shape(self) -> dict:
        nc, cl, ns = C.c_int32(), C.c_int32(), C.c_int32()
        fx, ix = (C.c_int32 * 256)(), (C.c_int32 * 256)()
        check(lib().h2aes_perm_query(self.kind, self.k, self.N, C.byref(nc), C.byref(cl), C.byref(ns),
                                     fx, ix))
        return dict(n_columns=nc.value, chunk_len=cl.value, n_sets=ns.value,
                    columns=[("fixed" if fx[i] else "advice", ix[i]) for i in range(nc.value)])

    def perm_sigma(self, n_blocks=None) -> torch.Tensor:
        nb = self.n_blocks if n_blocks is None else n_blocks
        out = torch.empty((self.perm_shape()["n_columns"], self.n, 4), dtype=torch.int64,
                          device=self.ctx.device)
        check(lib().h2aes_perm_sigma_dev(self._h, nb, _ptr(out)), self.ctx._h)
        return out

    def perm_product(self, beta, gamma, advice: torch.Tensor) -> torch.Tensor:
        """Z_perm [n_sets, n, 4] from the advice columns of the last witness call"""
        assert advice.is_cuda and advice.is_contiguous()
        assert advice.numel() == self.layout["n_advice"] * self.n * 4
        be, bep = _fr_arg(beta)
        ga, gap = _fr_arg(gamma)
        out = torch.empty((self.perm_shape()["n_sets"], self.n, 4), dtype=torch.int64,
                          device=self.ctx.device)
        check(lib().h2aes_perm_product_dev(self._h, bep, gap, _ptr(advice), _ptr(out)), self.ctx._h)
        return out

    def lookup_quotient(self, dom: "Domain", theta, beta, gamma, y, advice_ext, table_ext, sel_ext,
                        Ap_ext, Sp_ext, Z_ext, values_ext, first=0):
        """evaluate_h, lookup part (h2aes_lookup_quotient_dev): folds the five expressions of
        lookups [first, first + len(Z_ext)) into `values_ext` ([2^extended_k, 4], in place)."""
        count = Z_ext.shape[0]
        en = dom.extended_n
        for t, nc in ((advice_ext, self.layout["n_advice"]), (table_ext, 4), (sel_ext, count),
                      (Ap_ext, count), (Sp_ext, count), (Z_ext, count)):
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.int64
            assert tuple(t.shape) == (nc, en, 4), (tuple(t.shape), nc)
        assert values_ext.is_cuda and values_ext.is_contiguous() and tuple(values_ext.shape) == (en, 4)
        keep = [_fr_arg(v) for v in (theta, beta, gamma, y)]
        check(lib().h2aes_lookup_quotient_dev(self._h, dom._h, *[k[1] for k in keep], first, count,
                                              _ptr(advice_ext), _ptr(table_ext), _ptr(sel_ext),
                                              _ptr(Ap_ext), _ptr(Sp_ext), _ptr(Z_ext),
                                              _ptr(values_ext)), self.ctx._h)
        return values_ext

    def gate_perm_quotient(self, dom: "Domain", beta, gamma, y, advice_ext, rc_ext, qrc_ext, sigma_ext,
                           Zp_ext, values_ext):
        """evaluate_h before the lookups (h2aes_gate_perm_quotient_dev): the Equality-RC gate and
        the permutation argument's expressions folded into `values_ext` (in place)."""
        en = dom.extended_n
        ps = self.perm_shape()
        for t, shape in ((advice_ext, (self.layout["n_advice"], en, 4)), (rc_ext, (en, 4)), (qrc_ext, (en, 4)),
                         (sigma_ext, (ps["n_columns"], en, 4)), (Zp_ext, (ps["n_sets"], en, 4)),
                         (values_ext, (en, 4))):
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.int64 and tuple(t.shape) == shape, \
                (tuple(t.shape), shape)
        keep = [_fr_arg(v) for v in (beta, gamma, y)]
        check(lib().h2aes_gate_perm_quotient_dev(self._h, dom._h, *[k[1] for k in keep], _ptr(advice_ext),
                                                 _ptr(rc_ext), _ptr(qrc_ext), _ptr(sigma_ext),
                                                 _ptr(Zp_ext), _ptr(values_ext)), self.ctx._h)
        return values_ext

    def lookup_stats(self) -> np.ndarray:
        """rows per lookup whose grand-product factor is not one (last product call)"""
        out = np.zeros(self.layout["n_lookups"], np.uint64)
        check(lib().h2aes_lookup_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))),
              self.ctx._h)
        return out

    def lookup_check(self):
        """Synchronize and surface device-detected conditions (ERR_LOOKUP / ERR_ZERO_DENOM)."""
        check(lib().h2aes_lookup_check(self._h), self.ctx._h)

    # ---- row-sharded phases (see sharded.py for the collective plumbing) -----------------------
    def shard_hist(self, theta) -> torch.Tensor:
        th, thp = _fr_arg(theta)
        p = C.c_void_p()
        words = C.c_uint64()
        check(lib().h2aes_lookup_hist_dev(self._h, thp, C.byref(p), C.byref(words)), self.ctx._h)
        return _wrap_device_u32(p.value, words.value, self.ctx.device)

    def shard_permuted(self, want_inputs=True, want_outputs=True):
        L = self.layout["n_lookups"]
        bufs = {}
        if want_outputs:
            for nm in (["A", "S"] if want_inputs else []) + ["Ap", "Sp"]:
                bufs[nm] = torch.empty((L, self.rows, 4), dtype=torch.int64, device=self.ctx.device)
        check(lib().h2aes_lookup_permuted_shard_dev(self._h, _ptr(bufs.get("A")), _ptr(bufs.get("S")),
                                                    _ptr(bufs.get("Ap")), _ptr(bufs.get("Sp"))),
              self.ctx._h)
        return bufs

    def shard_partial(self, beta, gamma):
        """phase 2a: returns (Z buffer holding the factors, partial products [L * segments, 4])"""
        be, bep = _fr_arg(beta)
        ga, gap = _fr_arg(gamma)
        L = self.layout["n_lookups"]
        Z = torch.empty((L, self.rows, 4), dtype=torch.int64, device=self.ctx.device)
        p = C.c_void_p()
        check(lib().h2aes_lookup_partial_dev(self._h, bep, gap, _ptr(Z), C.byref(p)), self.ctx._h)
        return Z, _wrap_device_i64(p.value, self._partial_words, self.ctx.device).view(-1, 4)

    def shard_product(self, gathered: torch.Tensor, Z: torch.Tensor):
        """phase 2b: finishes Z in place from every shard's partial products"""
        L = self.layout["n_lookups"]
        assert gathered.is_cuda and gathered.is_contiguous()
        assert gathered.numel() == self.n_shards * self._partial_words
        check(lib().h2aes_lookup_product_shard_dev(self._h, _ptr(gathered), _ptr(Z)), self.ctx._h)
        return Z

    def lookup_columns_sharded(self, theta, beta=None, gamma=None, want_permuted=True, out=None):
        """h2aes_lookup_columns_sharded_dev: this shard's rows of A', S' and (with beta, gamma) Z of
        ALL lookups, both exchanges done by the library over its NCCL communicator."""
        L = self.layout["n_lookups"]
        bufs = dict(out or {})
        names = (["Ap", "Sp"] if want_permuted else []) + (["Z"] if beta is not None else [])
        for nm in names:
            if nm not in bufs:
                bufs[nm] = torch.empty((L, self.rows, 4), dtype=torch.int64, device=self.ctx.device)
        th, thp = _fr_arg(theta)
        be, bep = _fr_arg(beta) if beta is not None else (None, None)
        ga, gap = _fr_arg(gamma) if gamma is not None else (None, None)
        check(lib().h2aes_lookup_columns_sharded_dev(
            self._h, thp, bep, gap, _ptr(bufs.get("Ap")) if want_permuted else None,
            _ptr(bufs.get("Sp")) if want_permuted else None,
            _ptr(bufs.get("Z")) if beta is not None else None), self.ctx._h)
        return bufs

    def close(self):
        if getattr(self, "_h", None):
            lib().h2aes_circuit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Domain:
    """halo2's EvaluationDomain::new(j, k) on the GPU (poly/domain.rs [UPSTREAM]): the transforms
    keygen_pk / create_proof apply to the columns the circuit kernels produce.  Columns are int64
    CUDA tensors [ncols, len, 4]; the in-place calls return their argument."""

    def __init__(self, ctx: Context, k: int, j: int = _ffi.CS_DEGREE):
        self.ctx, self.k, self.j = ctx, k, j
        self.extended_k, self.quotient_len = _ffi.domain_query(k, j)
        self.n, self.extended_n = 1 << k, 1 << self.extended_k
        h = C.c_void_p()
        check(lib().h2aes_domain_create(ctx._h, j, k, C.byref(h)), ctx._h)
        self._h = h

    def _chk(self, t: torch.Tensor, length: int) -> int:
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.int64
        assert t.dim() == 3 and t.shape[1] == length and t.shape[2] == 4, tuple(t.shape)
        return t.shape[0]

    def fft(self, cols: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """best_fft(a, omega | omega_inv, k) in place"""
        check(lib().h2aes_fft_dev(self._h, _ptr(cols), self._chk(cols, self.n), int(inverse)), self.ctx._h)
        return cols

    def lagrange_to_coeff(self, cols: torch.Tensor) -> torch.Tensor:
        check(lib().h2aes_lagrange_to_coeff_dev(self._h, _ptr(cols), self._chk(cols, self.n)), self.ctx._h)
        return cols

    def coeff_to_extended(self, coeff: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        nc = self._chk(coeff, self.n)
        if out is None:
            out = torch.empty((nc, self.extended_n, 4), dtype=torch.int64, device=self.ctx.device)
        assert self._chk(out, self.extended_n) == nc
        check(lib().h2aes_coeff_to_extended_dev(self._h, _ptr(coeff), nc, _ptr(out)), self.ctx._h)
        return out

    def extended_to_coeff(self, ext: torch.Tensor) -> torch.Tensor:
        """in place; coefficients [quotient_len, extended_n) are zero afterwards"""
        check(lib().h2aes_extended_to_coeff_dev(self._h, _ptr(ext), self._chk(ext, self.extended_n)),
              self.ctx._h)
        return ext

    def divide_by_vanishing(self, ext: torch.Tensor) -> torch.Tensor:
        """EvaluationDomain::divide_by_vanishing_poly, in place on [ncols, 2^extended_k, 4]"""
        check(lib().h2aes_divide_by_vanishing_dev(self._h, _ptr(ext), self._chk(ext, self.extended_n)),
              self.ctx._h)
        return ext

    def to_extended(self, lagrange: torch.Tensor) -> torch.Tensor:
        """Lagrange columns [ncols, 2^k, 4] -> their extended cosets (a copy goes to coefficient form)"""
        return self.coeff_to_extended(self.lagrange_to_coeff(lagrange.clone()))

    def transform_host(self, op: int, cols: np.ndarray) -> np.ndarray:
        """h2aes_domain_transform_host: numpy uint64 [ncols, len_in, 4] -> [ncols, len_out, 4]"""
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        n_in = self.extended_n if op == _ffi.DOMAIN_EXTENDED_TO_COEFF else self.n
        n_out = {_ffi.DOMAIN_FFT: self.n, _ffi.DOMAIN_LAGRANGE_TO_COEFF: self.n,
                 _ffi.DOMAIN_COEFF_TO_EXTENDED: self.extended_n,
                 _ffi.DOMAIN_EXTENDED_TO_COEFF: self.quotient_len}[op]
        assert cols.ndim == 3 and cols.shape[1:] == (n_in, 4), cols.shape
        out = np.empty((cols.shape[0], n_out, 4), dtype=np.uint64)
        check(lib().h2aes_domain_transform_host(self._h, op, C.c_void_p(cols.ctypes.data), cols.shape[0],
                                                C.c_void_p(out.ctypes.data)), self.ctx._h)
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().h2aes_domain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synthetic_bases(ctx: Context, n: int, d) -> torch.Tensor:
    """[n, 2, 4] G1Affine points (1 + i*d) * G on the device (h2aes_g1_synthetic_bases_dev)"""
    out = torch.empty((n, 2, 4), dtype=torch.int64, device=ctx.device)
    da, dp = _fr_arg(d)
    check(lib().h2aes_g1_synthetic_bases_dev(ctx._h, dp, n, _ptr(out)), ctx._h)
    return out


class Msm:
    """best_multiexp over bn256 G1 on the GPU (h2aes_msm_*): commitments of Fr columns against a
    fixed set of bases (halo2curves G1Affine: int64 CUDA tensor [n, 2, 4], x then y)."""

    def __init__(self, ctx: Context, bases: torch.Tensor):
        assert bases.is_cuda and bases.is_contiguous() and bases.dtype == torch.int64
        assert bases.dim() == 3 and bases.shape[1:] == (2, 4), tuple(bases.shape)
        self.ctx, self.bases, self.n = ctx, bases, bases.shape[0]  # keeps the bases alive
        h = C.c_void_p()
        check(lib().h2aes_msm_create(ctx._h, _ptr(bases), self.n, C.byref(h)), ctx._h)
        self._h = h

    def enable_runs(self):
        """h2aes_msm_enable_runs: commit run-structured columns through their differences"""
        check(lib().h2aes_msm_enable_runs(self._h), self.ctx._h)
        return self

    def commit(self, cols: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """cols [ncols, n, 4] (device) -> [ncols, 2, 4] affine points (device); asynchronous"""
        assert cols.is_cuda and cols.is_contiguous() and cols.dtype == torch.int64
        assert cols.dim() == 3 and cols.shape[1:] == (self.n, 4), tuple(cols.shape)
        if out is None:
            out = torch.empty((cols.shape[0], 2, 4), dtype=torch.int64, device=self.ctx.device)
        check(lib().h2aes_msm_dev(self._h, _ptr(cols), cols.shape[0], _ptr(out)), self.ctx._h)
        return out

    def commit_host(self, cols: np.ndarray) -> np.ndarray:
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        assert cols.ndim == 3 and cols.shape[1:] == (self.n, 4), cols.shape
        out = np.empty((cols.shape[0], 2, 4), dtype=np.uint64)
        check(lib().h2aes_msm_host(self._h, C.c_void_p(cols.ctypes.data), cols.shape[0],
                                   C.c_void_p(out.ctypes.data)), self.ctx._h)
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().h2aes_msm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """Expose library-owned device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
            "strides": None,
        }


def _wrap_device_u32(ptr: int, count: int, device) -> torch.Tensor:
    # int32 view: sums of row counts stay far below 2^31 (n <= 2^26)
    return torch.as_tensor(_CudaArray(ptr, count, "<i4"), device=device)


def _wrap_device_i64(ptr: int, count: int, device) -> torch.Tensor:
    return torch.as_tensor(_CudaArray(ptr, count, "<i8"), device=device)


def to_host_u64(t: torch.Tensor) -> np.ndarray:
    """int64 device tensor of limbs -> numpy uint64 array (same bits)."""
    return t.cpu().numpy().view(np.uint64)
