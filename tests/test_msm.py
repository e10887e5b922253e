"""Commitments: multi-scalar multiplication over bn256 G1, halo2's
best_multiexp [UPSTREAM halo2_proofs v0.3.0 arithmetic.rs, halo2curves 0.6.1 bn256].

CPU part: the oracle's Fq / Jacobian group law / bucket-method restatement against a
first-principles affine group law on python integers and against the discrete-log anchor
(bases[i] = (1 + i d) G, so sum a_i bases[i] = (sum a_i (1 + i d)) G).  GPU part (`-m gpu`): the CUDA
MSM through the C ABI, bit-exact (affine coordinates as Montgomery limbs) against the oracle and
against the anchor.  Parity unpinned: no reference test holds a commitment."""
import numpy as np
import pytest

from conftest import R_MOD, SEED, splitmix_bytes

Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
RQ = (1 << 256) % Q
D_STEP = 0x1D2C3B4A5968778695A4B3C2D1E0F  # d of the test bases


def fq_int(limbs):
    v = sum(int(limbs[i]) << (64 * i) for i in range(4))
    return v * pow(RQ, -1, Q) % Q


def pt(p):
    x, y = fq_int(p[0]), fq_int(p[1])
    return None if (x, y) == (0, 0) else (x, y)


def ec_add(P, Qp):
    if P is None:
        return Qp
    if Qp is None:
        return P
    (x1, y1), (x2, y2) = P, Qp
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return x3, (lam * (x1 - x3) - y1) % Q


def ec_mul(k, P):
    R = None
    for b in bin(k)[2:] if k else "":
        R = ec_add(R, R)
        if b == "1":
            R = ec_add(R, P)
    return R


G = (1, 2)


def rand_scalars(orc, seed, n):
    a = splitmix_bytes(seed, n * 32).view(np.uint64).reshape(n, 4).copy()
    a[:, 3] %= np.uint64(0x30644E72E131A029)  # Montgomery limbs of some element < r
    return a


def scalar_values(orc, s):
    rinv = pow((1 << 256) % R_MOD, -1, R_MOD)
    return [orc.limbs_to_int(x) * rinv % R_MOD for x in s]


def anchor(orc, s):
    """(sum a_i (1 + i d)) * G by one scalar multiplication"""
    tot = sum(a * (1 + i * D_STEP) for i, a in enumerate(scalar_values(orc, s))) % R_MOD
    return ec_mul(tot, G)


# ---- CPU --------------------------------------------------------------------------------------
def test_oracle_group_law_against_python_integers(orc):
    k = int.from_bytes(splitmix_bytes(SEED + 40, 32).tobytes(), "little") % R_MOD
    g = orc.g1_mul_generator(orc.fr_from_int(k))
    assert pt(g) == ec_mul(k, G) and orc.g1_on_curve(g)
    assert pt(orc.g1_mul_generator(orc.fr_from_int(0))) is None
    assert pt(orc.g1_mul_generator(orc.fr_from_int(R_MOD - 1))) == (1, Q - 2)  # -G
    B = orc.g1_test_bases(64, orc.fr_from_int(D_STEP))
    for i in (0, 1, 5, 63):
        assert pt(B[i]) == ec_mul((1 + i * D_STEP) % R_MOD, G)
    assert pt(orc.g1_add(B[3], B[3])) == ec_mul(2 * (1 + 3 * D_STEP) % R_MOD, G)  # doubling
    neg = B[3].copy()
    neg[1] = orc.int_to_limbs((Q - fq_int(B[3][1])) * RQ % Q)
    assert pt(orc.g1_add(B[3], neg)) is None  # P + (-P)


def test_oracle_msm_against_the_anchor(orc):
    n = 3000
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    s = rand_scalars(orc, SEED + 41, n)
    s[7] = 0
    s[8] = orc.fr_from_int(R_MOD - 1)
    assert pt(orc.g1_msm(s, B)) == anchor(orc, s)
    assert pt(orc.g1_msm(np.zeros((n, 4), np.uint64), B)) is None


def test_signed_digit_recoding_model_for_every_window_size():
    """python model of k_msm_digits' integer logic (csrc/msm.cu): for every window size c the
    library can pick (4..16) the W = 254 // c + 1 signed digits stay within +-2^(c-1), rebuild the
    scalar, and leave no carry after the top window — on the extremes and on random scalars."""
    import random
    rng = random.Random(SEED)
    edge = [0, 1, 2, R_MOD - 1, R_MOD - 2, (1 << 253), (1 << 253) - 1, (1 << 128) - 1, (1 << 128),
            int("f" * 60, 16) % R_MOD, sum(1 << (16 * i + 15) for i in range(15))]
    cases = edge + [rng.randrange(R_MOD) for _ in range(300)]
    for c in range(4, 17):
        W, NB, full = 254 // c + 1, 1 << (c - 1), 1 << c
        for k in cases:
            limbs = [(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
            carry, total = 0, 0
            for w in range(W):
                bit = w * c
                lo, sh = bit >> 5, bit & 31
                assert lo < 8
                d = limbs[lo] >> sh
                if sh + c > 32 and lo + 1 < 8:
                    d |= (limbs[lo + 1] << (32 - sh)) & 0xFFFFFFFF
                d = (d & (full - 1)) + carry
                if d > NB:
                    digit, carry = d - full, 1
                else:
                    digit, carry = d, 0
                assert -NB <= digit <= NB
                total += digit << bit
            assert carry == 0 and total == k, (c, hex(k))


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
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def _host(t):
    from halo2_aes_b200.backend import to_host_u64
    return to_host_u64(t)


@pytest.mark.gpu
@pytest.mark.parametrize("c", [6, 7, 9, 11, 13, 16])
def test_msm_every_window_size_on_the_device(gpu, orc, ffi, c, monkeypatch):
    """The window size is chosen per column (k_msm_stats + msm_choose in csrc/msm.cu); H2AES_MSM_C
    pins it, so every size the choice ranges over is checked against the oracle on columns of all
    shapes (full width, short scalars that need fewer windows, bytes, 0/1, extremes), directly and
    through the differences."""
    from halo2_aes_b200.backend import Msm
    monkeypatch.setenv("H2AES_MSM_C", str(c))
    n = 3000
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    runs = rand_scalars(orc, SEED + 71, 40)
    cols = np.stack([
        rand_scalars(orc, SEED + 70 + c, n),
        np.repeat(runs, n // 40, axis=0),                           # long runs: the difference path
        np.stack([orc.fr_from_int(int(b)) for b in splitmix_bytes(SEED + 72, n)]),
        np.stack([orc.fr_from_int(int(b) & 1) for b in splitmix_bytes(SEED + 73, n)]),
        np.stack([orc.fr_from_int((int(b) << 9) | 0x1ff) for b in splitmix_bytes(SEED + 74, n)]),  # 17 bits
        np.tile(orc.fr_from_int(R_MOD - 1), (n, 1)),
        np.tile(orc.fr_from_int((1 << c) - 1), (n, 1)),             # a full digit in the lowest window
        np.tile(orc.fr_from_int(1 << (c - 1)), (n, 1)),             # the largest positive digit
    ])
    cols[0, 0] = orc.fr_from_int((1 << 253) + 12345)
    want = [orc.g1_msm(col, B) for col in cols]
    for with_runs in (False, True):
        msm = Msm(gpu, _dev(B))
        if with_runs:
            msm.enable_runs()
        got = _host(msm.commit(_dev(cols)))
        for ci in range(cols.shape[0]):
            assert (got[ci] == want[ci]).all(), (c, with_runs, ci)
        msm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 33, 1000, 1 << 12, 1 << 14])
def test_msm_equals_oracle_and_anchor(gpu, orc, ffi, n):
    from halo2_aes_b200.backend import Msm
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    msm = Msm(gpu, _dev(B))
    cols = np.stack([
        rand_scalars(orc, SEED + 50 + n, n),                       # uniform field elements
        np.zeros((n, 4), np.uint64),                                # all zero -> identity
        np.tile(orc.fr_from_int(1), (n, 1)),                        # sum of the bases
        np.tile(orc.fr_from_int(R_MOD - 1), (n, 1)),                # -(sum of the bases)
        np.stack([orc.fr_from_int(int(b)) for b in splitmix_bytes(SEED + 51, n)]),  # bytes (advice-like)
        np.stack([orc.fr_from_int(int(b) & 1) for b in splitmix_bytes(SEED + 52, n)]),  # 0/1 (selector-like)
    ])
    cols[0, 0] = orc.fr_from_int((1 << 253) + 12345)  # top window with a carry chain below it
    if n > 2:
        cols[0, 1] = orc.fr_from_int((1 << 128) - 1)
    got = _host(msm.commit(_dev(cols)))
    for ci in range(cols.shape[0]):
        want = orc.g1_msm(cols[ci], B)
        assert (got[ci] == want).all(), f"column {ci}: {pt(got[ci])} != {pt(want)}"
        assert orc.g1_on_curve(got[ci])
    assert pt(got[0]) == anchor(orc, cols[0])
    assert pt(got[1]) is None
    # host entry point
    assert (msm.commit_host(cols[:2]) == got[:2]).all()
    msm.close()


@pytest.mark.gpu
def test_msm_repeated_points_hit_the_doubling_case(gpu, orc, ffi):
    """equal bases with equal scalars land in one bucket: acc == point -> the doubling branch"""
    from halo2_aes_b200.backend import Msm
    n = 256
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    B[1::2] = B[0::2]                       # pairs of identical points
    s = rand_scalars(orc, SEED + 60, n)
    s[1::2] = s[0::2]                       # with identical scalars
    neg = B[10].copy()
    neg[1] = orc.int_to_limbs((Q - fq_int(B[10][1])) * RQ % Q)
    B[11] = neg                             # and one P, -P pair: cancellation
    msm = Msm(gpu, _dev(B))
    got = _host(msm.commit(_dev(s[None])))[0]
    assert (got == orc.g1_msm(s, B)).all()
    msm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 65, 1000, (1 << 14) + 3])
def test_run_structured_columns_through_differences(gpu, orc, ffi, n):
    """h2aes_msm_enable_runs: the same commitments, bit for bit, through sum (s_i - s_{i+1}) P_i"""
    from halo2_aes_b200.backend import Msm
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    vals = rand_scalars(orc, SEED + 80 + n, 8)
    rng = np.random.default_rng(n)
    cols = np.stack([
        vals[np.sort(rng.integers(0, 8, n))],            # sorted over 8 values (A'-like)
        vals[np.repeat(rng.integers(0, 8, (n + 15) // 16), 16)[:n]],  # constant stretches (Z-like)
        np.tile(vals[3], (n, 1)),                        # one run: a single term s * P_{n-1}
        rand_scalars(orc, SEED + 81 + n, n),             # uniform: keeps the direct form
        np.zeros((n, 4), np.uint64),
    ])
    plain = Msm(gpu, _dev(B))
    want = _host(plain.commit(_dev(cols)))
    runs = Msm(gpu, _dev(B)).enable_runs()
    got = _host(runs.commit(_dev(cols)))
    assert (got == want).all()
    for ci in range(cols.shape[0]):
        assert (got[ci] == orc.g1_msm(cols[ci], B)).all(), ci
    plain.close()
    runs.close()


@pytest.mark.gpu
def test_commitment_of_lookup_columns_with_runs(gpu, orc, ffi, challenges):
    """A', S', Z and Z_perm of a k=17 AES circuit: direct form == difference form == oracle"""
    import torch
    from halo2_aes_b200.backend import Circuit, Msm
    k = 17
    theta, beta, gamma = challenges
    key = splitmix_bytes(SEED + 62, 16).tobytes()
    pts = splitmix_bytes(SEED + 63, 60 * 16).reshape(60, 16)
    cir = Circuit(gpu, ffi.KIND_AES, k, 1)
    advice, _ = cir.witness(key, torch.from_numpy(pts).cuda())
    perm = cir.lookup_permuted(theta, first=1, count=2, want_inputs=False)
    Z = cir.lookup_product(beta, gamma, first=1, count=2)
    Zp = cir.perm_product(beta, gamma, advice)
    cir.lookup_check()
    cols = torch.cat([perm["Ap"], perm["Sp"], Z, Zp[:1]])
    B = orc.g1_test_bases(1 << k, orc.fr_from_int(D_STEP))
    plain, runs = Msm(gpu, _dev(B)), Msm(gpu, _dev(B)).enable_runs()
    want, got = _host(plain.commit(cols)), _host(runs.commit(cols))
    assert (got == want).all()
    h = _host(cols)
    for ci in (0, 2, 4, 6):
        assert (got[ci] == orc.g1_msm(h[ci], B)).all(), ci
    plain.close()
    runs.close()
    cir.close()


@pytest.mark.gpu
def test_commitment_of_witness_columns(gpu, orc, ffi):
    """advice columns of the key-schedule circuit (k = 17) committed against 2^17 bases"""
    from halo2_aes_b200.backend import Circuit, Msm
    k = 17
    key = splitmix_bytes(SEED + 61, 16).tobytes()
    cir = Circuit(gpu, ffi.KIND_KEYSCHED, k, 1)
    advice, _ = cir.witness(key, None)
    B = orc.g1_test_bases(1 << k, orc.fr_from_int(D_STEP))
    msm = Msm(gpu, _dev(B))
    got = _host(msm.commit(advice))
    ref = orc.Circuit(orc.KIND_KEYSCHED, k, 1).synthesize(key, None)
    for col in range(advice.shape[0]):
        assert (got[col] == orc.g1_msm(ref.advice()[col], B)).all(), col
    msm.close()
    cir.close()


@pytest.mark.gpu
def test_msm_bench_size_anchor_and_linearity(gpu, orc, ffi):
    """n = 2^20 (the bench shape's column length): anchor on uniform scalars, and
    commit(a) + commit(b) == commit(a + b) with the sum formed by the oracle's field addition"""
    import ctypes as C
    from halo2_aes_b200.backend import Msm
    n = 1 << 20
    B = orc.g1_test_bases(n, orc.fr_from_int(D_STEP))
    msm = Msm(gpu, _dev(B))
    # a, b supported on the first 4096 rows (their sum is formed row by row with the oracle's
    # field addition); the anchor below uses all 2^20 rows
    a, b = np.zeros((n, 4), np.uint64), np.zeros((n, 4), np.uint64)
    a[:4096], b[:4096] = rand_scalars(orc, SEED + 70, 4096), rand_scalars(orc, SEED + 71, 4096)
    s = np.zeros_like(a)
    u64p = C.POINTER(C.c_uint64)
    for i in range(4096):
        orc.lib().orc_fr_add(a[i].ctypes.data_as(u64p), b[i].ctypes.data_as(u64p),
                             s[i].ctypes.data_as(u64p))
    got = _host(msm.commit(_dev(np.stack([a, b, s]))))
    assert (orc.g1_add(got[0], got[1]) == got[2]).all()
    full = rand_scalars(orc, SEED + 72, n)
    got_full = _host(msm.commit(_dev(full[None])))[0]
    assert pt(got_full) == anchor(orc, full)
    msm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 5000, 1 << 17])
def test_synthetic_bases_equal_the_oracle(gpu, orc, ffi, n):
    from halo2_aes_b200.backend import synthetic_bases
    got = _host(synthetic_bases(gpu, n, orc.fr_from_int(D_STEP)))
    assert (got == orc.g1_test_bases(n, orc.fr_from_int(D_STEP))).all()


@pytest.mark.gpu
def test_msm_error_behaviour(gpu, ffi):
    import ctypes as C
    lib = ffi.lib()
    h = C.c_void_p()
    assert lib.h2aes_msm_create(gpu._h, None, 16, C.byref(h)) == ffi.ERR_ARG
    assert lib.h2aes_msm_create(None, None, 16, C.byref(h)) == ffi.ERR_ARG
    assert lib.h2aes_msm_dev(None, None, 1, None) == ffi.ERR_ARG
    assert lib.h2aes_msm_host(None, None, 1, None) == ffi.ERR_ARG
    assert lib.h2aes_msm_enable_runs(None) == ffi.ERR_ARG
    assert lib.h2aes_g1_synthetic_bases_dev(gpu._h, None, 4, None) == ffi.ERR_ARG
