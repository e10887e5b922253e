"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/h2aes_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package (halo2_aes_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "libh2aes_oracle.so")

KIND_AES = 0
KIND_KEYSCHED = 1

ERRORS = {
    1: "bad argument",
    2: "AES calls too many. doesn't fit in the rows",
    3: "not enough rows available",
    4: "Keys should be scheduled",
    5: "lookup input not in table",
    6: "from_bytes returned None",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}: {ERRORS.get(code, '?')}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile beside it (gcc only; no CUDA)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_DIR, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("h2aes_oracle.cpp", "fr.hpp", "Makefile")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u64p = C.POINTER(C.c_uint64)
        u8p = C.POINTER(C.c_uint8)
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_int] * 4
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_record_copies.argtypes = [C.c_void_p, C.c_int]
        L.orc_synthesize.argtypes = [C.c_void_p, u8p, u8p, C.c_uint64]
        L.orc_num_copies.restype = C.c_uint64
        L.orc_num_copies.argtypes = [C.c_void_p]
        L.orc_num_regions.restype = C.c_uint64
        L.orc_num_regions.argtypes = [C.c_void_p]
        L.orc_get_copies.argtypes = [C.c_void_p, u64p]
        L.orc_get_advice.argtypes = [C.c_void_p, C.c_int, u64p]
        L.orc_get_fixed.argtypes = [C.c_void_p, C.c_int, u64p]
        L.orc_get_selector.argtypes = [C.c_void_p, C.c_int, u8p]
        L.orc_get_ciphertexts.argtypes = [C.c_void_p, u8p]
        L.orc_get_round_keys.argtypes = [C.c_void_p, u8p]
        L.orc_lookup.argtypes = [C.c_void_p, C.c_int] + [u64p] * 8
        L.orc_mock_verify.argtypes = [C.c_void_p, u64p]
        L.orc_mock_verify_ext.argtypes = [C.c_int] * 3 + [u64p, u64p, u8p, u64p, C.c_uint64, u64p]
        L.orc_shape.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int32)]
        L.orc_fr_from_u64.argtypes = [C.c_uint64, u64p]
        for f in (L.orc_fr_mul, L.orc_fr_add, L.orc_xor_bytes):
            f.argtypes = [u64p, u64p, u64p]
        for f in (L.orc_fr_inv, L.orc_fr_to_canonical, L.orc_fr_from_canonical):
            f.argtypes = [u64p, u64p]
        L.orc_tables.argtypes = [C.c_int, u8p, u8p, u8p]
        L.orc_bench_circuit.argtypes = (
            [C.c_int] * 3 + [u8p, u8p, C.c_uint64, u64p, u64p, u64p, C.c_int, C.c_int,
                             C.POINTER(C.c_double), u64p]
        )
        L.orc_perm_shape.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int32)] * 3
        L.orc_permutation.argtypes = [C.c_void_p, u64p, u64p, u64p, C.POINTER(C.c_uint32), u64p]
        L.orc_domain_constants.argtypes = [C.c_int, u64p, u64p]
        L.orc_domain_transform.restype = C.c_int64
        L.orc_domain_transform.argtypes = [C.c_int, C.c_int, C.c_int, u64p, u64p]
        L.orc_domain_query.argtypes = [C.c_int, C.c_int, u64p]
        L.orc_lookup_quotient.argtypes = [C.c_void_p, C.c_int, C.c_int] + [u64p] * 5
        L.orc_g1_msm.argtypes = [u64p, u64p, C.c_uint64, u64p]
        L.orc_g1_test_bases.argtypes = [u64p, C.c_uint64, u64p]
        L.orc_g1_mul_generator.argtypes = [u64p, u64p]
        L.orc_g1_on_curve.argtypes = [u64p]
        L.orc_g1_add_affine.argtypes = [u64p, u64p, u64p]
        L.orc_gate_perm_quotient.argtypes = [C.c_void_p] + [u64p] * 4
        L.orc_divide_by_vanishing.argtypes = [C.c_int, C.c_int, u64p]
        L.orc_poly_eval.argtypes = [u64p, C.c_uint64, u64p, u64p]
        L.orc_bench_batch.argtypes = (
            [C.c_int] * 4 + [u8p, u8p, C.c_uint64, u64p, u64p, u64p, C.c_int,
                             C.POINTER(C.c_double), u64p]
        )
        _lib = L
    return _lib


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R_MONT = (1 << 256) % R_MOD


def limbs_to_int(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(4))


def int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fr_from_int(v: int) -> np.ndarray:
    """canonical integer -> Montgomery limbs (pure Python big-int; independent of the C code)."""
    return int_to_limbs((v % R_MOD) * R_MONT % R_MOD)


def fr_to_int(limbs) -> int:
    """Montgomery limbs -> canonical integer."""
    return limbs_to_int(limbs) * pow(R_MONT, -1, R_MOD) % R_MOD


def shape(kind: int, k: int, N: int) -> dict:
    out = (C.c_int32 * 5)()
    rc = lib().orc_shape(kind, k, N, out)
    if rc:
        raise OracleError(rc)
    return dict(n_advice=out[0], n_fixed=out[1], n_selectors=out[2], n_lookups=out[3],
                n_sets=out[4], n=1 << k, usable=(1 << k) - 6)


def perm_shape(kind: int, k: int, N: int) -> dict:
    """columns of the permutation argument in enable_equality order, chunking"""
    out = (C.c_int32 * 3)()
    types = (C.c_int32 * 256)()
    idx = (C.c_int32 * 256)()
    lib().orc_perm_shape(kind, k, N, out, types, idx)
    return dict(n_columns=out[0], chunk_len=out[1], n_sets=out[2],
                columns=[("advice" if types[i] == 0 else "fixed", idx[i]) for i in range(out[0])])


def domain_constants(k: int):
    om, de = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    lib().orc_domain_constants(k, _u64(om), _u64(de))
    return om, de


# [UPSTREAM] poly/domain.rs EvaluationDomain restated (oracle/h2aes_oracle.cpp: domain_new ...)
FFT, LAGRANGE_TO_COEFF, COEFF_TO_EXTENDED, EXTENDED_TO_COEFF = 0, 1, 2, 3
CS_DEGREE = 5  # cs.degree() of both circuits: the lookups have degree 2 + 2 + 1


def domain_query(k: int, j: int = CS_DEGREE) -> dict:
    out = np.zeros((8, 4), np.uint64)
    ek = lib().orc_domain_query(j, k, _u64(out))
    names = ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset",
             "g_coset_inv", "ifft_divisor", "extended_ifft_divisor")
    d = {nm: out[i].copy() for i, nm in enumerate(names)}
    d["extended_k"] = ek
    d["quotient_poly_degree"] = j - 1
    return d


def domain_transform(op: int, k: int, a: np.ndarray, j: int = CS_DEGREE) -> np.ndarray:
    """One column through best_fft / lagrange_to_coeff / coeff_to_extended / extended_to_coeff."""
    a = np.ascontiguousarray(a, np.uint64)
    ek = domain_query(k, j)["extended_k"]
    n_in = (1 << ek) if op == EXTENDED_TO_COEFF else (1 << k)
    assert a.shape == (n_in, 4), a.shape
    n_out = {FFT: 1 << k, LAGRANGE_TO_COEFF: 1 << k, COEFF_TO_EXTENDED: 1 << ek,
             EXTENDED_TO_COEFF: (1 << k) * (j - 1)}[op]
    out = np.zeros((n_out, 4), np.uint64)
    got = lib().orc_domain_transform(j, k, op, _u64(a), _u64(out))
    if got != n_out:
        raise OracleError(int(-got))
    return out


def divide_by_vanishing(k: int, a: np.ndarray, j: int = CS_DEGREE) -> np.ndarray:
    """EvaluationDomain::divide_by_vanishing_poly on one extended column (returns a copy)."""
    a = np.array(a, dtype=np.uint64, order="C", copy=True)
    assert a.shape == (1 << domain_query(k, j)["extended_k"], 4)
    lib().orc_divide_by_vanishing(j, k, _u64(a))
    return a


def poly_eval(coeffs: np.ndarray, x) -> np.ndarray:
    """Horner evaluation of a coefficient column at the field element x (Montgomery limbs)."""
    coeffs = np.ascontiguousarray(coeffs, np.uint64)
    xx = np.ascontiguousarray(x, np.uint64)
    out = np.zeros(4, np.uint64)
    lib().orc_poly_eval(_u64(coeffs), coeffs.shape[0], _u64(xx), _u64(out))
    return out


# [UPSTREAM] bn256 G1 + best_multiexp restated (oracle/g1.hpp); points are [2, 4] uint64 (x, y)
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def g1_msm(scalars: np.ndarray, bases: np.ndarray) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars, np.uint64)
    bases = np.ascontiguousarray(bases, np.uint64)
    n = scalars.shape[0]
    assert scalars.shape == (n, 4) and bases.shape == (n, 2, 4)
    out = np.zeros((2, 4), np.uint64)
    lib().orc_g1_msm(_u64(scalars), _u64(bases), n, _u64(out))
    return out


def g1_test_bases(n: int, d) -> np.ndarray:
    """bases[i] = (1 + i*d) * G: an SRS stand-in whose discrete logarithms are known"""
    dd = np.ascontiguousarray(d, np.uint64)
    out = np.zeros((n, 2, 4), np.uint64)
    lib().orc_g1_test_bases(_u64(dd), n, _u64(out))
    return out


def g1_mul_generator(k) -> np.ndarray:
    kk = np.ascontiguousarray(k, np.uint64)
    out = np.zeros((2, 4), np.uint64)
    lib().orc_g1_mul_generator(_u64(kk), _u64(out))
    return out


def g1_on_curve(p) -> bool:
    return bool(lib().orc_g1_on_curve(_u64(np.ascontiguousarray(p, np.uint64))))


def g1_add(a, b) -> np.ndarray:
    out = np.zeros((2, 4), np.uint64)
    lib().orc_g1_add_affine(_u64(np.ascontiguousarray(a, np.uint64)), _u64(np.ascontiguousarray(b, np.uint64)),
                            _u64(out))
    return out


def tables(fips: bool = False):
    s = np.zeros(256, np.uint8)
    m2 = np.zeros(256, np.uint8)
    m3 = np.zeros(256, np.uint8)
    lib().orc_tables(int(fips), _u8(s), _u8(m2), _u8(m3))
    return s, m2, m3


class Circuit:
    """One synthesized circuit: the reference's `synthesize` replayed on the CPU."""

    def __init__(self, kind: int, k: int, N: int, fips_sbox: bool = False):
        self.kind, self.k, self.N = kind, k, N
        self.n = 1 << k
        self.info = shape(kind, k, N)
        self._h = lib().orc_new(kind, k, N, int(fips_sbox))
        if not self._h:
            raise OracleError(1)
        self.n_blocks = 0

    def close(self):
        if self._h:
            lib().orc_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def synthesize(self, key, plaintexts=None):
        key = np.ascontiguousarray(np.frombuffer(bytes(key), dtype=np.uint8))
        assert key.size == 16
        if plaintexts is None:
            pts = np.zeros((0, 16), np.uint8)
        else:
            pts = np.ascontiguousarray(np.asarray(plaintexts, dtype=np.uint8).reshape(-1, 16))
        self.n_blocks = pts.shape[0]
        dummy = np.zeros(16, np.uint8)
        rc = lib().orc_synthesize(self._h, _u8(key), _u8(pts if pts.size else dummy), self.n_blocks)
        if rc:
            raise OracleError(rc)
        return self

    def advice(self, col=None) -> np.ndarray:
        if col is None:
            return np.stack([self.advice(c) for c in range(self.info["n_advice"])])
        out = np.zeros((self.n, 4), np.uint64)
        rc = lib().orc_get_advice(self._h, col, _u64(out))
        if rc:
            raise OracleError(rc)
        return out

    def fixed(self, col=None) -> np.ndarray:
        if col is None:
            return np.stack([self.fixed(c) for c in range(self.info["n_fixed"])])
        out = np.zeros((self.n, 4), np.uint64)
        rc = lib().orc_get_fixed(self._h, col, _u64(out))
        if rc:
            raise OracleError(rc)
        return out

    def selectors(self, sel=None) -> np.ndarray:
        if sel is None:
            return np.stack([self.selectors(c) for c in range(self.info["n_selectors"])])
        out = np.zeros(self.n, np.uint8)
        rc = lib().orc_get_selector(self._h, sel, _u8(out))
        if rc:
            raise OracleError(rc)
        return out

    def copies(self) -> np.ndarray:
        m = lib().orc_num_copies(self._h)
        out = np.zeros((max(m, 1), 4), np.uint64)
        lib().orc_get_copies(self._h, _u64(out))
        return out[:m]

    def num_regions(self) -> int:
        return lib().orc_num_regions(self._h)

    def ciphertexts(self) -> np.ndarray:
        out = np.zeros((max(self.n_blocks, 1), 16), np.uint8)
        lib().orc_get_ciphertexts(self._h, _u8(out))
        return out[: self.n_blocks]

    def round_keys(self) -> np.ndarray:
        out = np.zeros(176, np.uint8)
        lib().orc_get_round_keys(self._h, _u8(out))
        return out

    def lookup(self, idx: int, theta, beta, gamma, want=("A", "S", "Ap", "Sp", "Z")) -> dict:
        """halo2 v0.3.0 commit_permuted + commit_product for lookup `idx` (Montgomery limbs)."""
        bufs = {}
        ptrs = []
        null = C.POINTER(C.c_uint64)()
        for name in ("A", "S", "Ap", "Sp", "Z"):
            if name in want:
                bufs[name] = np.zeros((self.n, 4), np.uint64)
                ptrs.append(_u64(bufs[name]))
            else:
                ptrs.append(null)
        th = np.ascontiguousarray(theta, dtype=np.uint64)
        be = np.ascontiguousarray(beta, dtype=np.uint64)
        ga = np.ascontiguousarray(gamma, dtype=np.uint64)
        rc = lib().orc_lookup(self._h, idx, _u64(th), _u64(be), _u64(ga), *ptrs)
        if rc:
            raise OracleError(rc)
        return bufs

    def lookup_quotient(self, theta, beta, gamma, y, first=0, count=None, values=None) -> np.ndarray:
        """halo2 v0.3.0 evaluate_h, lookup terms of lookups [first, first+count): the running
        extended-domain accumulator `values` (zero when None) after those lookups."""
        sh = shape(self.kind, self.k, self.N)
        count = sh["n_lookups"] - first if count is None else count
        en = 1 << domain_query(self.k)["extended_k"]
        acc = np.zeros((en, 4), np.uint64) if values is None else np.array(values, np.uint64, order="C", copy=True)
        assert acc.shape == (en, 4)
        args = [np.ascontiguousarray(v, dtype=np.uint64) for v in (theta, beta, gamma, y)]
        rc = lib().orc_lookup_quotient(self._h, first, count, *[_u64(v) for v in args], _u64(acc))
        if rc:
            raise OracleError(rc)
        return acc

    def gate_perm_quotient(self, beta, gamma, y, values=None) -> np.ndarray:
        """halo2 v0.3.0 evaluate_h up to the lookups: the custom gate, then the permutation
        argument's expressions, folded into `values` (zero when None)."""
        en = 1 << domain_query(self.k)["extended_k"]
        acc = np.zeros((en, 4), np.uint64) if values is None else np.array(values, np.uint64, order="C", copy=True)
        args = [np.ascontiguousarray(v, dtype=np.uint64) for v in (beta, gamma, y)]
        rc = lib().orc_gate_perm_quotient(self._h, *[_u64(v) for v in args], _u64(acc))
        if rc:
            raise OracleError(rc)
        return acc

    def permutation(self, beta, gamma, want=("sigma", "mapping", "Z")) -> dict:
        """halo2 v0.3.0 permutation argument: sigma columns (keygen), the (column, row) mapping
        they encode, and the grand products Z_perm (one per chunk of 3 columns)."""
        ps = perm_shape(self.kind, self.k, self.N)
        out = {}
        null64 = C.POINTER(C.c_uint64)()
        null32 = C.POINTER(C.c_uint32)()
        if "sigma" in want:
            out["sigma"] = np.zeros((ps["n_columns"], self.n, 4), np.uint64)
        if "mapping" in want:
            out["mapping"] = np.zeros((ps["n_columns"], self.n, 2), np.uint32)
        if "Z" in want:
            out["Z"] = np.zeros((ps["n_sets"], self.n, 4), np.uint64)
        be = np.ascontiguousarray(beta, dtype=np.uint64)
        ga = np.ascontiguousarray(gamma, dtype=np.uint64)
        rc = lib().orc_permutation(
            self._h, _u64(be), _u64(ga), _u64(out["sigma"]) if "sigma" in out else null64,
            out["mapping"].ctypes.data_as(C.POINTER(C.c_uint32)) if "mapping" in out else null32,
            _u64(out["Z"]) if "Z" in out else null64)
        if rc:
            raise OracleError(rc)
        return out

    def mock_verify(self):
        where = np.zeros(2, np.uint64)
        rc = lib().orc_mock_verify(self._h, _u64(where))
        return rc, (int(where[0]), int(where[1]))


def mock_verify_ext(kind, k, N, advice, fixed, selectors, copies):
    """Restated MockProver over caller-supplied columns (e.g. read back from the GPU)."""
    advice = np.ascontiguousarray(advice, dtype=np.uint64)
    fixed = np.ascontiguousarray(fixed, dtype=np.uint64)
    selectors = np.ascontiguousarray(selectors, dtype=np.uint8)
    copies = np.ascontiguousarray(copies, dtype=np.uint64).reshape(-1, 4)
    info = shape(kind, k, N)
    n = 1 << k
    assert advice.shape == (info["n_advice"], n, 4), advice.shape
    assert fixed.shape == (info["n_fixed"], n, 4), fixed.shape
    assert selectors.shape == (info["n_selectors"], n), selectors.shape
    where = np.zeros(2, np.uint64)
    cp = copies if copies.size else np.zeros((1, 4), np.uint64)
    rc = lib().orc_mock_verify_ext(kind, k, N, _u64(advice), _u64(fixed), _u8(selectors),
                                   _u64(cp), copies.shape[0], _u64(where))
    return rc, (int(where[0]), int(where[1]))


def bench_circuit(kind, k, N, key, plaintexts, theta, beta, gamma, threads=1, n_lookups=-1):
    """Time one whole circuit on the CPU: (synthesize seconds, lookup seconds, Z[n-6] digest)."""
    key = np.ascontiguousarray(np.frombuffer(bytes(key), dtype=np.uint8))
    pts = np.ascontiguousarray(np.asarray(plaintexts, dtype=np.uint8).reshape(-1, 16))
    times = (C.c_double * 2)()
    digest = np.zeros(4, np.uint64)
    th = np.ascontiguousarray(theta, dtype=np.uint64)
    be = np.ascontiguousarray(beta, dtype=np.uint64)
    ga = np.ascontiguousarray(gamma, dtype=np.uint64)
    dummy = np.zeros(16, np.uint8)
    rc = lib().orc_bench_circuit(kind, k, N, _u8(key), _u8(pts if pts.size else dummy),
                                 pts.shape[0], _u64(th), _u64(be), _u64(ga), threads, n_lookups,
                                 times, _u64(digest))
    if rc:
        raise OracleError(rc)
    return times[0], times[1], digest


def bench_batch(kind, k, N, keys, plaintexts, theta, beta, gamma, threads=1):
    """Time `len(keys)` independent circuits, one per OpenMP thread: (seconds, digest)."""
    keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint8).reshape(-1, 16))
    n_circ = keys.shape[0]
    pts = np.ascontiguousarray(np.asarray(plaintexts, dtype=np.uint8).reshape(n_circ, -1, 16))
    nb = pts.shape[1]
    sec = C.c_double()
    digest = np.zeros(4, np.uint64)
    th = np.ascontiguousarray(theta, dtype=np.uint64)
    be = np.ascontiguousarray(beta, dtype=np.uint64)
    ga = np.ascontiguousarray(gamma, dtype=np.uint64)
    rc = lib().orc_bench_batch(kind, k, N, n_circ, _u8(keys), _u8(pts), nb, _u64(th), _u64(be),
                               _u64(ga), threads, C.byref(sec), _u64(digest))
    if rc:
        raise OracleError(rc)
    return sec.value, digest
