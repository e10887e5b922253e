"""Evaluation domain: halo2's EvaluationDomain transforms.

CPU part: the oracle restatement of poly/domain.rs + arithmetic.rs best_fft [UPSTREAM] against
first-principles big-integer DFTs, the re-derived constants, and the host-side shape query of the
C ABI.  GPU part (`-m gpu`): the CUDA four-step transforms, called through the C ABI, bit-exact
against the oracle, plus size-independent properties at the bench size (round trips, linearity,
point evaluations by Horner's rule).  Parity unpinned in the sense of DESIGN.md section 2: no
reference test holds transform outputs; the DFT definition itself is the anchor."""
import ctypes as C

import numpy as np
import pytest

from conftest import R_MOD, SEED, splitmix_bytes

R_MONT = (1 << 256) % R_MOD


def rand_fr_cols(seed, ncols, n):
    """uniform-ish Montgomery limbs < r: top limb below r's top limb"""
    a = splitmix_bytes(seed, ncols * n * 32).view(np.uint64).reshape(ncols, n, 4).copy()
    a[..., 3] %= np.uint64(0x30644E72E131A029)
    return a


def to_ints(orc, col):
    return [orc.fr_to_int(x) for x in col]


# ---- CPU: oracle ------------------------------------------------------------------------------
def test_domain_constants_rederived(orc):
    d = orc.domain_query(17)
    assert d["extended_k"] == 19 and d["quotient_poly_degree"] == 4
    zeta = orc.fr_to_int(d["g_coset"])
    assert zeta == pow(7, (R_MOD - 1) // 3, R_MOD) and pow(zeta, 3, R_MOD) == 1 and zeta != 1
    assert orc.fr_to_int(d["g_coset_inv"]) == zeta * zeta % R_MOD
    root = pow(7, (R_MOD - 1) >> 28, R_MOD)
    assert orc.fr_to_int(d["extended_omega"]) == pow(root, 1 << (28 - 19), R_MOD)
    assert orc.fr_to_int(d["omega"]) == pow(root, 1 << (28 - 17), R_MOD)
    assert orc.fr_to_int(d["omega"]) * orc.fr_to_int(d["omega_inv"]) % R_MOD == 1
    assert orc.fr_to_int(d["ifft_divisor"]) == pow(1 << 17, -1, R_MOD)
    assert orc.fr_to_int(d["extended_ifft_divisor"]) == pow(1 << 19, -1, R_MOD)
    # extended_k rule of EvaluationDomain::new: smallest 2^ek >= n * (j - 1)
    for j, want in ((2, 10), (3, 11), (4, 12), (5, 12), (6, 13), (9, 13), (10, 14)):
        assert orc.domain_query(10, j)["extended_k"] == want


@pytest.mark.parametrize("k", [1, 2, 5, 9])
def test_oracle_transforms_against_bigint_dft(orc, k):
    n = 1 << k
    d = orc.domain_query(k)
    w, we = orc.fr_to_int(d["omega"]), orc.fr_to_int(d["extended_omega"])
    zeta = orc.fr_to_int(d["g_coset"])
    a = rand_fr_cols(SEED + k, 1, n)[0]
    vals = to_ints(orc, a)
    f = orc.domain_transform(orc.FFT, k, a)
    for j in range(0, n, max(1, n // 7)):
        assert orc.fr_to_int(f[j]) == sum(v * pow(w, i * j, R_MOD) for i, v in enumerate(vals)) % R_MOD
    assert (orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, f) == a).all()
    e = orc.domain_transform(orc.COEFF_TO_EXTENDED, k, a)
    en = 1 << d["extended_k"]
    assert e.shape == (en, 4)
    for j in range(0, en, max(1, en // 5)):
        x = zeta * pow(we, j, R_MOD) % R_MOD
        assert orc.fr_to_int(e[j]) == sum(v * pow(x, i, R_MOD) for i, v in enumerate(vals)) % R_MOD
    back = orc.domain_transform(orc.EXTENDED_TO_COEFF, k, e)
    assert back.shape == (n * 4, 4) and (back[:n] == a).all() and not back[n:].any()


def test_domain_query_abi_matches_oracle(ffi, orc):
    for k in (2, 10, 17, 20):
        for j in (2, 3, 5, 9):
            ek, q = ffi.domain_query(k, j)
            assert ek == orc.domain_query(k, j)["extended_k"] and q == (1 << k) * (j - 1)
    for bad in ((0, 5), (27, 5)):
        with pytest.raises(ffi.H2AesError):
            ffi.domain_query(*bad)
    assert ffi.lib().h2aes_domain_create(None, 5, 17, None) == ffi.ERR_ARG
    assert ffi.lib().h2aes_fft_dev(None, None, 1, 0) == ffi.ERR_ARG


# ---- GPU --------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu(ffi):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from halo2_aes_b200.backend import Context
    ctx = Context(0)
    yield ctx
    ctx.close()


def _dev(a):
    import torch
    return torch.from_numpy(a.view(np.int64)).cuda()


def _host(t):
    from halo2_aes_b200.backend import to_host_u64
    return to_host_u64(t)


@pytest.mark.gpu
@pytest.mark.parametrize("k,ncols", [(2, 2), (5, 3), (12, 9), (13, 2), (17, 2)])
def test_transforms_equal_the_oracle(gpu, orc, ffi, k, ncols):
    from halo2_aes_b200.backend import Domain
    dom = Domain(gpu, k)
    n, en = dom.n, dom.extended_n
    assert dom.extended_k == orc.domain_query(k)["extended_k"]
    a = rand_fr_cols(SEED + 100 + k, ncols, n)
    a[0, 1::2] = 0           # sparse column
    if ncols > 2:
        a[2] = 0             # all-zero column
    check_cols = range(ncols) if k < 17 else (0, ncols - 1)
    f = _host(dom.fft(_dev(a)))
    fi = _host(dom.fft(_dev(a), inverse=True))
    c = _host(dom.lagrange_to_coeff(_dev(a)))
    e_t = dom.coeff_to_extended(_dev(a))
    e = _host(e_t)
    q = _host(dom.extended_to_coeff(e_t))
    for ci in check_cols:
        assert (f[ci] == orc.domain_transform(orc.FFT, k, a[ci])).all(), "fft"
        ref_c = orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, a[ci])
        assert (c[ci] == ref_c).all(), "lagrange_to_coeff"
        assert (e[ci] == orc.domain_transform(orc.COEFF_TO_EXTENDED, k, a[ci])).all(), "coeff_to_extended"
        ref_q = orc.domain_transform(orc.EXTENDED_TO_COEFF, k, e[ci])
        assert (q[ci, :dom.quotient_len] == ref_q).all() and not q[ci, dom.quotient_len:].any()
    # fft with omega^-1 is lagrange_to_coeff without the 1/n divisor: n * coeff
    nfr = orc.fr_from_int(n)
    for i in (0, 1, n - 1):
        got = np.zeros(4, np.uint64)
        orc.lib().orc_fr_mul(c[0, i].ctypes.data_as(C.POINTER(C.c_uint64)),
                             nfr.ctypes.data_as(C.POINTER(C.c_uint64)),
                             got.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert (got == fi[0, i]).all()
    # the host entry point gives the same columns
    assert (dom.transform_host(ffi.DOMAIN_LAGRANGE_TO_COEFF, a) == c).all()
    assert (dom.transform_host(ffi.DOMAIN_COEFF_TO_EXTENDED, a[:1]) == e[:1]).all()
    assert (dom.transform_host(ffi.DOMAIN_EXTENDED_TO_COEFF, e[:1])[0] == q[0, :dom.quotient_len]).all()
    dom.close()


@pytest.mark.gpu
def test_bench_size_properties(gpu, orc, ffi):
    """K=20 (extended 2^22): round trips and Horner evaluations."""
    import torch
    from halo2_aes_b200.backend import Domain
    k = 20
    dom = Domain(gpu, k)
    n = dom.n
    a = rand_fr_cols(SEED + 7, 3, n)
    d_a = _dev(a)
    coeff = dom.lagrange_to_coeff(d_a.clone())
    # round trip: best_fft(lagrange_to_coeff(a)) == a
    back = dom.fft(coeff.clone())
    assert torch.equal(back, d_a)
    ext = dom.coeff_to_extended(coeff)
    h_ext, h_coeff = _host(ext[:1]), _host(coeff[:1])
    # E[j] == p(zeta * extended_omega^j) by Horner's rule on the coefficients
    d = orc.domain_query(k)
    zeta, we = orc.fr_to_int(d["g_coset"]), orc.fr_to_int(d["extended_omega"])
    rinv = pow(R_MONT, -1, R_MOD)
    cvals = [orc.limbs_to_int(x) * rinv % R_MOD for x in h_coeff[0]]
    for j in (0, 3, (1 << 22) - 1):
        x = zeta * pow(we, j, R_MOD) % R_MOD
        acc = 0
        for v in reversed(cvals):
            acc = (acc * x + v) % R_MOD
        assert orc.fr_to_int(h_ext[0, j]) == acc, j
    # extended_to_coeff(coeff_to_extended(c)) == c padded with zeros
    q = dom.extended_to_coeff(ext)
    assert torch.equal(q[:, :n], coeff) and not bool(q[:, n:].any())
    dom.close()


@pytest.mark.gpu
def test_linearity(gpu, orc, ffi):
    """T(x + y) == T(x) + T(y) for every transform (k = 13, sums formed with python integers)."""
    from halo2_aes_b200.backend import Domain
    k = 13
    dom = Domain(gpu, k)
    a = rand_fr_cols(SEED + 9, 2, dom.n)

    def add(u, v):  # Montgomery limbs add like the values they stand for
        return np.stack([orc.int_to_limbs((orc.limbs_to_int(p) + orc.limbs_to_int(q)) % R_MOD)
                         for p, q in zip(u, v)])

    cols = np.stack([a[0], a[1], add(a[0], a[1])])
    for op in (ffi.DOMAIN_FFT, ffi.DOMAIN_LAGRANGE_TO_COEFF, ffi.DOMAIN_COEFF_TO_EXTENDED):
        t = dom.transform_host(op, cols)
        assert (add(t[0], t[1]) == t[2]).all(), op
    dom.close()


@pytest.mark.gpu
def test_witness_column_through_the_domain(gpu, orc, ffi):
    """The advice columns the circuit kernels produce, taken to coefficient form on the device
    and compared with the oracle's lagrange_to_coeff of the oracle's advice columns."""
    from halo2_aes_b200.backend import Circuit, Domain
    k = 17
    key = splitmix_bytes(SEED, 16).tobytes()
    cir = Circuit(gpu, ffi.KIND_KEYSCHED, k, 1)
    advice, _ = cir.witness(key, None)
    dom = Domain(gpu, k)
    coeff = _host(dom.lagrange_to_coeff(advice.clone()))
    ref = orc.Circuit(orc.KIND_KEYSCHED, k, 1).synthesize(key, None)
    for col in (0, 3):
        assert (coeff[col] == orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, ref.advice()[col])).all()
    dom.close()
    cir.close()


@pytest.mark.gpu
def test_domain_error_behaviour(gpu, ffi):
    from halo2_aes_b200.backend import Domain
    h = C.c_void_p()
    lib = ffi.lib()
    assert lib.h2aes_domain_create(gpu._h, 5, 23, C.byref(h)) == ffi.ERR_ARG  # extended_k = 25
    assert lib.h2aes_domain_create(gpu._h, 1, 17, C.byref(h)) == ffi.ERR_ARG
    dom = Domain(gpu, 10)
    assert lib.h2aes_fft_dev(dom._h, None, 1, 0) == ffi.ERR_ARG
    assert lib.h2aes_coeff_to_extended_dev(dom._h, None, 1, None) == ffi.ERR_ARG
    assert lib.h2aes_domain_transform_host(dom._h, 7, None, 1, None) == ffi.ERR_ARG
    dom.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [21, 22])
def test_three_pass_transforms_equal_the_oracle(gpu, orc, ffi, k):
    """extended_k = 23, 24 (the K=21 / K=22 circuits of BASELINE config 5): a 2^extended_k-point
    transform no longer fits two shared-memory passes and runs as three (N = 2^7..8 x 2^8 x 2^8);
    one column through coeff_to_extended and extended_to_coeff against the oracle, the 2^k-point
    transforms (two passes) beside it."""
    from halo2_aes_b200.backend import Domain
    dom = Domain(gpu, k)
    assert dom.extended_k == k + 2
    a = rand_fr_cols(SEED + 300 + k, 1, dom.n)
    c = _host(dom.lagrange_to_coeff(_dev(a)))
    assert (c[0] == orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, a[0])).all(), "lagrange_to_coeff"
    e_t = dom.coeff_to_extended(_dev(a))
    e = _host(e_t)
    assert (e[0] == orc.domain_transform(orc.COEFF_TO_EXTENDED, k, a[0])).all(), "coeff_to_extended"
    q = _host(dom.extended_to_coeff(e_t))
    ref_q = orc.domain_transform(orc.EXTENDED_TO_COEFF, k, e[0])
    assert (q[0, :dom.quotient_len] == ref_q).all() and not q[0, dom.quotient_len:].any()
    dom.close()


@pytest.mark.gpu
def test_domain_limit_is_reported(gpu, ffi):
    """The documented limit (include/h2aes.h): extended_k <= 24, i.e. k <= 22 for these circuits."""
    import ctypes as C
    L = ffi.lib()
    h = C.c_void_p()
    assert L.h2aes_domain_create(gpu._h, ffi.CS_DEGREE, 23, C.byref(h)) == ffi.ERR_ARG
    assert not h.value
    assert b"extended_k <= 24" in L.h2aes_ctx_last_error(gpu._h)
    assert ffi.domain_query(22)[0] == 24
