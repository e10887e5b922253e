"""Timing of the prover-side steps that follow the witness path at the
bench shape: every column of one K=20, N=5 circuit (advice, table, selectors, A'/S'/Z, sigma, Z_perm) to coefficient form and to the extended
coset, the gate / permutation / lookup terms of the quotient, division by the vanishing polynomial and the way back
to coefficients.  CUDA events on the context's stream, inputs resident in HBM.  Prints one JSON
object; not part of the headline metric."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo2_aes_b200 import _ffi  # noqa: E402
from halo2_aes_b200.backend import Circuit, Context, Domain, Msm, synthetic_bases  # noqa: E402

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FR_MUL_PEAK_G = 67.0  # measured ceiling of the Montgomery multiply on B200 (profiles/microbench)


def fr(seed):
    v = (seed * 0x9E3779B97F4A7C15 + 12345) ** 5 % R_MOD
    m = v * ((1 << 256) % R_MOD) % R_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def measure(ctx, K=20, N=5):
    cir = Circuit(ctx, _ffi.KIND_AES, K, N)
    nb = cir.layout["capacity_blocks"]
    rng = np.random.default_rng(0x68326165)
    key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    pts = torch.from_numpy(rng.integers(0, 256, (nb, 16), dtype=np.uint8)).cuda()
    theta, beta, gamma, y = fr(1), fr(2), fr(3), fr(4)
    advice, _ = cir.witness(key, pts)
    fixed, sel = cir.fixed_columns()
    L = cir.layout["n_lookups"]
    perm = cir.lookup_permuted(theta, want_inputs=False)
    Z = cir.lookup_product(beta, gamma)
    cir.lookup_check()
    sigma = cir.perm_sigma()
    Zp = cir.perm_product(beta, gamma, advice)
    dom = Domain(ctx, K)
    en, ek = dom.extended_n, dom.extended_k
    stream = ctx.stream

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r

    # commitments of the witness-dependent columns (best_multiexp against synthetic bases)
    ms_bases, bases = timed(lambda: synthetic_bases(ctx, 1 << K, fr(5)))
    msm = Msm(ctx, bases)
    msm.commit(advice[:1])  # warm-up
    ctx.profile(True)
    ms_c_adv, c_adv = timed(lambda: msm.commit(advice))
    helper = torch.cat([perm["Ap"], perm["Sp"], Z, Zp])
    ms_c_help, c_help = timed(lambda: msm.commit(helper))
    kern_msm = ctx.profile_read()
    # the same columns through their differences against the running sums of the bases
    ms_prefix, _ = timed(lambda: msm.enable_runs())
    msm.commit(helper[:1])
    ctx.profile_read()
    ms_c_runs, c_runs = timed(lambda: msm.commit(helper))
    kern_runs = ctx.profile_read()
    ctx.profile(False)
    assert bool(c_adv.any()) and bool(c_help.any()) and torch.equal(c_runs, c_help)
    n_help = int(helper.shape[0])
    del helper
    msm.close()
    del bases

    cols = {"advice": advice, "table": fixed[:4].contiguous(), "sel": sel[:L].contiguous(),
            "Ap": perm["Ap"], "Sp": perm["Sp"], "Z": Z,
            "rc_qrc": torch.stack([fixed[4], sel[L]]), "sigma": sigma, "Zp": Zp}
    ncols = sum(t.shape[0] for t in cols.values())
    dom.to_extended(cols["table"])  # warm-up (twiddles, kernel attributes)
    ctx.profile(True)
    ms_l2c, coeff = timed(lambda: {nm: dom.lagrange_to_coeff(t) for nm, t in cols.items()})  # in place
    ms_c2e, ext = timed(lambda: {nm: dom.coeff_to_extended(t) for nm, t in coeff.items()})
    values = torch.zeros((en, 4), dtype=torch.int64, device="cuda")
    cir.lookup_quotient(dom, theta, beta, gamma, y, ext["advice"], ext["table"], ext["sel"][:1],
                        ext["Ap"][:1], ext["Sp"][:1], ext["Z"][:1], values.clone())  # Lagrange columns
    ms_gp, _ = timed(lambda: cir.gate_perm_quotient(dom, beta, gamma, y, ext["advice"], ext["rc_qrc"][0],
                                                    ext["rc_qrc"][1], ext["sigma"], ext["Zp"], values))
    ms_q, _ = timed(lambda: cir.lookup_quotient(dom, theta, beta, gamma, y, ext["advice"], ext["table"],
                                                ext["sel"], ext["Ap"], ext["Sp"], ext["Z"], values))
    ms_div, _ = timed(lambda: dom.divide_by_vanishing(values[None]))
    ms_e2c, _ = timed(lambda: dom.extended_to_coeff(values[None]))
    kern = ctx.profile_read()
    ctx.profile(False)
    n = 1 << K
    mul_l2c = ncols * (n // 2 * (K - 2) + 2 * n)
    mul_c2e = ncols * (en // 2 * (ek - 2) + en + n)
    arity = {0: 2, 1: 4, 2: 3, 3: 3, 4: 3}
    mul_q = en * (3 + sum(17 + arity[l % 5] - 1 for l in range(L)))
    bytes_q = en * 32 * (3 + 4 + cir.layout["n_advice"] + 6 * L + 2)
    out = {
        "what": "prover-side steps after the witness path (SURVEY 8f items 3/4), one circuit, "
                "not part of the headline metric",
        "shape": {"k": K, "N": N, "blocks": int(nb), "extended_k": ek, "columns": int(ncols)},
        "commitments": {"bases": "synthetic: (1 + i d) G, built on the device", "bases_ms": ms_bases,
                        "advice": {"columns": int(advice.shape[0]), "ms": ms_c_adv,
                                   "ms_per_column": ms_c_adv / advice.shape[0]},
                        "Ap_Sp_Z_Zperm": {"columns": n_help, "ms": ms_c_help,
                                          "ms_per_column": ms_c_help / n_help},
                        "Ap_Sp_Z_Zperm_through_differences": {
                            "columns": n_help, "ms": ms_c_runs, "ms_per_column": ms_c_runs / n_help,
                            "running_sums_of_the_bases_ms_once": ms_prefix,
                            "kernels": {k_: {"launches": v[0], "ms": v[1]} for k_, v in kern_runs.items()}},
                        "kernels": {k_: {"launches": v[0], "ms": v[1]} for k_, v in kern_msm.items()}},
        "lagrange_to_coeff": {"ms": ms_l2c, "ms_per_column": ms_l2c / ncols, "fr_multiplies": mul_l2c,
                              "gmul_s": mul_l2c / ms_l2c / 1e6,
                              "frac_of_mul_peak": mul_l2c / ms_l2c / 1e6 / FR_MUL_PEAK_G},
        "coeff_to_extended": {"ms": ms_c2e, "ms_per_column": ms_c2e / ncols, "fr_multiplies": mul_c2e,
                              "gmul_s": mul_c2e / ms_c2e / 1e6,
                              "frac_of_mul_peak": mul_c2e / ms_c2e / 1e6 / FR_MUL_PEAK_G},
        "lookup_quotient": {"ms": ms_q, "lookups": L, "fr_multiplies": mul_q, "gmul_s": mul_q / ms_q / 1e6,
                            "frac_of_mul_peak": mul_q / ms_q / 1e6 / FR_MUL_PEAK_G,
                            "algorithmic_bytes": bytes_q, "hbm_gbs": bytes_q / ms_q / 1e6},
        "gate_perm_quotient": {"ms": ms_gp, "permutation_columns": int(sigma.shape[0]),
                               "permutation_sets": int(Zp.shape[0])},
        "divide_by_vanishing": {"ms": ms_div, "hbm_gbs": 2 * en * 32 / ms_div / 1e6},
        "extended_to_coeff": {"ms": ms_e2c},
        "kernels": {k_: {"launches": v[0], "ms": v[1]} for k_, v in kern.items()},
        "fr_mul_peak_gmul_s": FR_MUL_PEAK_G,
    }
    dom.close()
    cir.close()
    return out


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    print(json.dumps(measure(Context(0), K, N)))
