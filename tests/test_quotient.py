"""Lookup terms of the quotient polynomial: halo2 v0.3.0
plonk/evaluation.rs evaluate_h (lookup part) + divide_by_vanishing_poly [UPSTREAM].

CPU part: the oracle restatement is checked against the equation halo2's VERIFIER uses — at a
random point x, h(x) * (x^n - 1) must equal the y-folded lookup expressions built from the
column polynomials evaluated at x, omega*x and x/omega — which is independent of the FFTs, the
coset and the evaluation order of the restatement.  GPU part (`-m gpu`): the CUDA kernels through
the C ABI, bit-exact against the oracle, and the same verifier equation at the bench shape.
Parity unpinned: no reference test holds these values."""
import numpy as np
import pytest

from conftest import R_MOD, SEED, seeded_fr, splitmix_bytes

TAGS = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5}  # table.rs Tag: U8, Xor, Sbox, GfMul2, GfMul3
ARITY = {0: 2, 1: 4, 2: 3, 3: 3, 4: 3}  # columns per lookup (tag + operands)


def lagrange_at(k, i, x):
    """l_i(x) for the 2^k domain by the closed form (omega^i / n) (x^n - 1) / (x - omega^i)."""
    n = 1 << k
    w = pow(pow(7, (R_MOD - 1) >> 28, R_MOD), 1 << (28 - k), R_MOD)
    wi = pow(w, i, R_MOD)
    return wi * pow(n, -1, R_MOD) * (pow(x, n, R_MOD) - 1) * pow(x - wi, -1, R_MOD) % R_MOD


def verifier_numerator(orc, ref, kind, k, lookups, theta, beta, gamma, y, x, start=0):
    """fold of the five lookup expressions per lookup at the point x, from polynomial evaluations"""
    n = 1 << k
    w = pow(pow(7, (R_MOD - 1) >> 28, R_MOD), 1 << (28 - k), R_MOD)
    th, be, ga, yy = (orc.fr_to_int(v) for v in (theta, beta, gamma, y))

    def ev(col, at):
        c = orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, col)
        return orc.fr_to_int(orc.poly_eval(c, orc.fr_from_int(at % R_MOD)))

    l0 = lagrange_at(k, 0, x)
    l_last = lagrange_at(k, n - 6, x)
    l_blind = sum(lagrange_at(k, n - 1 - i, x) for i in range(5)) % R_MOD
    l_active = (1 - l_last - l_blind) % R_MOD
    fixed = ref.fixed()
    tab = [ev(fixed[t], x) for t in range(4)]
    adv_cache = {}
    acc = start
    one = orc.fr_from_int(1)
    for li in lookups:
        set_, t = (li // 5, li % 5) if kind == orc.KIND_AES else (0, li)
        r = ref.lookup(li, theta, beta, gamma, want=("Ap", "Sp", "Z"))
        q = ev(np.where(ref.selectors()[li][:, None] != 0, one[None, :], np.uint64(0)), x)
        a = q * TAGS[t] % R_MOD
        s = tab[0]
        for j in range(ARITY[t] - 1):
            col = 3 * set_ + j
            if col not in adv_cache:
                adv_cache[col] = ev(ref.advice()[col], x)
            a = (a * th + q * adv_cache[col]) % R_MOD
            s = (s * th + tab[1 + j]) % R_MOD
        ap, sp, z = ev(r["Ap"], x), ev(r["Sp"], x), ev(r["Z"], x)
        z_next, ap_prev = ev(r["Z"], x * w), ev(r["Ap"], x * pow(w, -1, R_MOD))
        terms = [(1 - z) * l0,
                 (z * z - z) * l_last,
                 (z_next * (ap + be) * (sp + ga) - z * (a + be) * (s + ga)) * l_active,
                 (ap - sp) * l0,
                 (ap - sp) * (ap - ap_prev) * l_active]
        for term in terms:
            acc = (acc * yy + term) % R_MOD
    return acc


def verifier_gate_perm(orc, ref, kind, k, N, beta, gamma, y, x):
    """the custom gate, then the permutation argument's expressions at the point x"""
    n = 1 << k
    w = pow(pow(7, (R_MOD - 1) >> 28, R_MOD), 1 << (28 - k), R_MOD)
    delta = pow(7, 1 << 28, R_MOD)
    be, ga, yy = (orc.fr_to_int(v) for v in (beta, gamma, y))
    one = orc.fr_from_int(1)

    def ev(col, at):
        c = orc.domain_transform(orc.LAGRANGE_TO_COEFF, k, col)
        return orc.fr_to_int(orc.poly_eval(c, orc.fr_from_int(at % R_MOD)))

    sels, adv, fixed = ref.selectors(), ref.advice(), ref.fixed()
    q = ev(np.where(sels[-1][:, None] != 0, one[None, :], np.uint64(0)), x)  # q_eq_rcon is last
    acc = q * (ev(adv[-1], x) - ev(fixed[4], x)) % R_MOD  # words_column is the last advice column
    l0 = lagrange_at(k, 0, x)
    l_last = lagrange_at(k, n - 6, x)
    l_active = (1 - l_last - sum(lagrange_at(k, n - 1 - i, x) for i in range(5))) % R_MOD
    pr = ref.permutation(beta, gamma, want=("sigma", "Z"))
    cols = orc.perm_shape(kind, k, N)["columns"]
    z = [ev(zc, x) for zc in pr["Z"]]
    z_next = [ev(zc, x * w) for zc in pr["Z"]]
    z_last_rot = [ev(zc, x * pow(w, -6, R_MOD)) for zc in pr["Z"]]
    acc = (acc * yy + (1 - z[0]) * l0) % R_MOD
    acc = (acc * yy + (z[-1] * z[-1] - z[-1]) * l_last) % R_MOD
    for i in range(1, len(z)):
        acc = (acc * yy + (z[i] - z_last_rot[i - 1]) * l0) % R_MOD
    cur = be * x % R_MOD
    for i in range(len(z)):
        left, right = z_next[i], z[i]
        chunk = list(range(3 * i, min(3 * i + 3, len(cols))))
        vals = [ev(adv[cols[p][1]] if cols[p][0] == "advice" else fixed[cols[p][1]], x) for p in chunk]
        for p, v in zip(chunk, vals):
            left = left * (v + be * ev(pr["sigma"][p], x) + ga) % R_MOD
        for v in vals:
            right = right * (v + cur + ga) % R_MOD
            cur = cur * delta % R_MOD
        acc = (acc * yy + (left - right) * l_active) % R_MOD
    return acc


def check_verifier_equation(orc, ref, kind, k, lookups, chal, values, x, start=0):
    """h = values / (X^n - 1) on the extended coset, to coefficients; h(x) (x^n - 1) == N(x)"""
    theta, beta, gamma, y = chal
    h_ext = orc.divide_by_vanishing(k, values)
    h = orc.domain_transform(orc.EXTENDED_TO_COEFF, k, h_ext)
    hx = orc.fr_to_int(orc.poly_eval(h, orc.fr_from_int(x)))
    lhs = hx * (pow(x, 1 << k, R_MOD) - 1) % R_MOD
    rhs = verifier_numerator(orc, ref, kind, k, lookups, theta, beta, gamma, y, x, start)
    assert lhs == rhs


@pytest.fixture(scope="module")
def chal4(challenges):
    return (*challenges, seeded_fr(SEED + 4))


@pytest.fixture(scope="module")
def keysched(orc):
    key = splitmix_bytes(SEED + 11, 16).tobytes()
    return key, orc.Circuit(orc.KIND_KEYSCHED, 17, 1).synthesize(key, None)


def test_oracle_full_numerator_satisfies_the_verifier_equation(orc, keysched, chal4):
    """gate, permutation, lookups in halo2's order (evaluate_h of the whole constraint system)"""
    _, ref = keysched
    theta, beta, gamma, y = chal4
    # lookups accumulated in two calls on top of the gate + permutation terms: one fold
    part = ref.lookup_quotient(*chal4, first=0, count=1, values=ref.gate_perm_quotient(beta, gamma, y))
    values = ref.lookup_quotient(*chal4, first=1, count=2, values=part)
    x = int.from_bytes(splitmix_bytes(SEED + 13, 40).tobytes(), "little") % R_MOD
    start = verifier_gate_perm(orc, ref, orc.KIND_KEYSCHED, 17, 1, beta, gamma, y, x)
    check_verifier_equation(orc, ref, orc.KIND_KEYSCHED, 17, range(3), chal4, values, x, start=start)


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


def _host(t):
    from halo2_aes_b200.backend import to_host_u64
    return to_host_u64(t)


def gpu_lookup_quotient(gpu, ffi, kind, k, N, key, pts, chal, first=0, count=None, start=None):
    """witness -> lookup columns -> extended cosets -> lookup terms, all on the device"""
    import torch
    from halo2_aes_b200.backend import Circuit, Domain
    theta, beta, gamma, y = chal
    cir = Circuit(gpu, kind, k, N)
    d_pts = torch.from_numpy(np.ascontiguousarray(pts)).cuda() if pts is not None and len(pts) else None
    advice, _ = cir.witness(key, d_pts)
    fixed, sel = cir.fixed_columns()
    L = cir.layout["n_lookups"]
    count = L - first if count is None else count
    perm = cir.lookup_permuted(theta, first=first, count=count, want_inputs=False)
    Z = cir.lookup_product(beta, gamma, first=first, count=count)
    cir.lookup_check()
    dom = Domain(gpu, k)
    ext = {nm: dom.to_extended(t) for nm, t in
           (("advice", advice), ("table", fixed[:4].contiguous()),
            ("sel", sel[first:first + count].contiguous()), ("Ap", perm["Ap"]), ("Sp", perm["Sp"]), ("Z", Z))}
    values = torch.zeros((dom.extended_n, 4), dtype=torch.int64, device="cuda")
    if start is not None:
        values.copy_(torch.from_numpy(start.view(np.int64)))
    cir.lookup_quotient(dom, theta, beta, gamma, y, ext["advice"], ext["table"], ext["sel"],
                        ext["Ap"], ext["Sp"], ext["Z"], values, first=first)
    out = _host(values)
    h = dom.extended_to_coeff(dom.divide_by_vanishing(values[None].clone()))
    h_coeff = _host(h)[0]
    dom.close()
    cir.close()
    return out, h_coeff


@pytest.mark.gpu
def test_key_schedule_circuit_equals_the_oracle(gpu, orc, ffi, keysched, chal4):
    key, ref = keysched
    values, h = gpu_lookup_quotient(gpu, ffi, ffi.KIND_KEYSCHED, 17, 1, key, None, chal4)
    want = ref.lookup_quotient(*chal4)
    bad = np.flatnonzero((values != want).any(axis=1))
    assert bad.size == 0, f"{bad.size} extended rows differ, first {bad[:5]}"
    # the quotient h = values / (X^n - 1) as the device computes it
    h_ref = orc.domain_transform(orc.EXTENDED_TO_COEFF, 17, orc.divide_by_vanishing(17, want))
    assert (h[:h_ref.shape[0]] == h_ref).all() and not h[h_ref.shape[0]:].any()


@pytest.mark.gpu
def test_aes_blocks_partial_range_and_start_value(gpu, orc, ffi, chal4):
    """k=17, N=1 AES circuit with 40 random blocks: lookups 1..3 on top of a non-zero accumulator."""
    key = splitmix_bytes(SEED + 21, 16).tobytes()
    pts = splitmix_bytes(SEED + 22, 40 * 16).reshape(40, 16)
    ref = orc.Circuit(orc.KIND_AES, 17, 1).synthesize(key, pts)
    start = np.stack([seeded_fr(SEED + 30 + i) for i in range(8)] * (1 << 16))
    values, _ = gpu_lookup_quotient(gpu, ffi, ffi.KIND_AES, 17, 1, key, pts, chal4, first=1, count=3,
                                    start=start)
    want = ref.lookup_quotient(*chal4, first=1, count=3, values=start)
    assert (values == want).all()


@pytest.mark.gpu
def test_verifier_equation_two_column_sets(gpu, orc, ffi, chal4):
    """k=17, N=2 (10 lookups, 150 blocks over both column sets): the device quotient satisfies
    h(x) (x^n - 1) == folded lookup expressions at a random x (no oracle quotient involved)."""
    key = splitmix_bytes(SEED + 23, 16).tobytes()
    pts = splitmix_bytes(SEED + 24, 150 * 16).reshape(150, 16)
    ref = orc.Circuit(orc.KIND_AES, 17, 2).synthesize(key, pts)
    _, h = gpu_lookup_quotient(gpu, ffi, ffi.KIND_AES, 17, 2, key, pts, chal4)
    x = int.from_bytes(splitmix_bytes(SEED + 25, 40).tobytes(), "little") % R_MOD
    hx = orc.fr_to_int(orc.poly_eval(h, orc.fr_from_int(x)))
    lhs = hx * (pow(x, 1 << 17, R_MOD) - 1) % R_MOD
    rhs = verifier_numerator(orc, ref, orc.KIND_AES, 17, range(10), *chal4, x)
    assert lhs == rhs


def gpu_full_numerator(gpu, ffi, kind, k, N, key, pts, chal):
    """gate + permutation + every lookup, in halo2's order, all on the device; returns the
    accumulator after the gate/permutation terms, after the lookups, and h's coefficients"""
    import torch
    from halo2_aes_b200.backend import Circuit, Domain
    theta, beta, gamma, y = chal
    cir = Circuit(gpu, kind, k, N)
    d_pts = torch.from_numpy(np.ascontiguousarray(pts)).cuda() if pts is not None and len(pts) else None
    advice, _ = cir.witness(key, d_pts)
    fixed, sel = cir.fixed_columns()
    L = cir.layout["n_lookups"]
    perm = cir.lookup_permuted(theta, want_inputs=False)
    Z = cir.lookup_product(beta, gamma)
    sigma = cir.perm_sigma()
    Zp = cir.perm_product(beta, gamma, advice)
    cir.lookup_check()
    dom = Domain(gpu, k)
    ext = {nm: dom.to_extended(t) for nm, t in
           (("advice", advice), ("fixed", fixed), ("sel", sel), ("Ap", perm["Ap"]), ("Sp", perm["Sp"]),
            ("Z", Z), ("sigma", sigma), ("Zp", Zp))}
    values = torch.zeros((dom.extended_n, 4), dtype=torch.int64, device="cuda")
    cir.gate_perm_quotient(dom, beta, gamma, y, ext["advice"], ext["fixed"][4], ext["sel"][L],
                           ext["sigma"], ext["Zp"], values)
    after_perm = _host(values)
    cir.lookup_quotient(dom, theta, beta, gamma, y, ext["advice"], ext["fixed"][:4].contiguous(),
                        ext["sel"][:L].contiguous(), ext["Ap"], ext["Sp"], ext["Z"], values)
    after_lookups = _host(values)
    h = _host(dom.extended_to_coeff(dom.divide_by_vanishing(values[None])))[0]
    dom.close()
    cir.close()
    return after_perm, after_lookups, h


@pytest.mark.gpu
def test_full_numerator_key_schedule_equals_the_oracle(gpu, orc, ffi, keysched, chal4):
    key, ref = keysched
    theta, beta, gamma, y = chal4
    after_perm, after_lookups, _ = gpu_full_numerator(gpu, ffi, ffi.KIND_KEYSCHED, 17, 1, key, None, chal4)
    want = ref.gate_perm_quotient(beta, gamma, y)
    bad = np.flatnonzero((after_perm != want).any(axis=1))
    assert bad.size == 0, f"gate/permutation: {bad.size} extended rows differ, first {bad[:5]}"
    assert (after_lookups == ref.lookup_quotient(*chal4, values=want)).all()


@pytest.mark.gpu
def test_full_numerator_verifier_equation_two_column_sets(gpu, orc, ffi, chal4):
    """k=17, N=2, 150 blocks over both column sets (8 permutation columns in 3 chained sets, 10
    lookups): the device's h satisfies the verifier's equation for the WHOLE constraint system."""
    key = splitmix_bytes(SEED + 26, 16).tobytes()
    pts = splitmix_bytes(SEED + 27, 150 * 16).reshape(150, 16)
    theta, beta, gamma, y = chal4
    ref = orc.Circuit(orc.KIND_AES, 17, 2).synthesize(key, pts)
    _, _, h = gpu_full_numerator(gpu, ffi, ffi.KIND_AES, 17, 2, key, pts, chal4)
    x = int.from_bytes(splitmix_bytes(SEED + 28, 40).tobytes(), "little") % R_MOD
    hx = orc.fr_to_int(orc.poly_eval(h, orc.fr_from_int(x)))
    lhs = hx * (pow(x, 1 << 17, R_MOD) - 1) % R_MOD
    start = verifier_gate_perm(orc, ref, orc.KIND_AES, 17, 2, beta, gamma, y, x)
    rhs = verifier_numerator(orc, ref, orc.KIND_AES, 17, range(10), theta, beta, gamma, y, x, start)
    assert lhs == rhs


@pytest.mark.gpu
def test_quotient_error_behaviour(gpu, ffi):
    import ctypes as C
    from halo2_aes_b200.backend import Circuit, Domain
    lib = ffi.lib()
    cir = Circuit(gpu, ffi.KIND_KEYSCHED, 17, 1)
    dom = Domain(gpu, 12)
    assert lib.h2aes_lookup_quotient_dev(cir._h, None, *[None] * 4, 0, 1, *[None] * 7) == ffi.ERR_ARG
    assert lib.h2aes_lookup_quotient_dev(cir._h, dom._h, *[None] * 4, 0, 1, *[None] * 7) == ffi.ERR_ARG
    assert lib.h2aes_gate_perm_quotient_dev(cir._h, dom._h, *[None] * 9) == ffi.ERR_ARG
    assert lib.h2aes_divide_by_vanishing_dev(dom._h, None, 1) == ffi.ERR_ARG
    assert lib.h2aes_divide_by_vanishing_dev(None, None, 1) == ffi.ERR_ARG
    dom.close()
    cir.close()
