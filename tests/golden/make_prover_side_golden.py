"""First-principles known answers for the prover-side steps (evaluation domain, G1, MSM), written
to tests/golden/prover_side_vectors.json.  Pure python integers: no oracle, no GPU, no reference
code — the definitions themselves:

  * bn256 Fr / Fq moduli, the multiplicative generator 7, S = 28 (halo2curves 0.6.1 [UPSTREAM]):
    ROOT_OF_UNITY = 7^((r-1)/2^28), DELTA = 7^(2^28), ZETA = 7^((r-1)/3)
  * best_fft(a, omega, k)[j] = sum_i a[i] omega^(ij); lagrange_to_coeff = its inverse;
    coeff_to_extended(p)[j] = p(zeta * extended_omega^j)   (halo2_proofs poly/domain.rs [UPSTREAM])
  * G1: y^2 = x^3 + 3 over Fq, generator (1, 2), affine chord-and-tangent law
  * best_multiexp(scalars, bases) = sum_i scalars[i] * bases[i]

    python tests/golden/make_prover_side_golden.py
All values are canonical integers as hex strings (tests move them to Montgomery limbs).
"""
import json
import os

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prover_side_vectors.json")


def splitmix(seed, n):
    out, x = [], seed
    for _ in range(n):
        x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        out.append(z ^ (z >> 31))
    return out


def field_elems(seed, n, mod):
    w = splitmix(seed, 5 * n)
    return [sum(w[5 * i + j] << (64 * j) for j in range(5)) % mod for i in range(n)]


def ec_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    (x1, y1), (x2, y2) = P, S
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return x3, (lam * (x1 - x3) - y1) % Q


def ec_mul(k, P):
    acc = None
    for b in bin(k)[2:] if k else "":
        acc = ec_add(acc, acc)
        if b == "1":
            acc = ec_add(acc, P)
    return acc


def hx(v):
    return hex(v)


def main():
    root = pow(7, (R - 1) >> 28, R)
    k, j = 5, 5                      # 32-point domain, cs.degree() = 5 -> extended_k = 7
    n, ek = 1 << k, 7
    omega, ext_omega = pow(root, 1 << (28 - k), R), pow(root, 1 << (28 - ek), R)
    zeta, delta = pow(7, (R - 1) // 3, R), pow(7, 1 << 28, R)
    a = field_elems(0x68326165, n, R)
    fft = [sum(v * pow(omega, i * t, R) for i, v in enumerate(a)) % R for t in range(n)]
    ninv = pow(n, -1, R)
    coeff = [ninv * sum(v * pow(omega, -i * t, R) for i, v in enumerate(a)) % R for t in range(n)]
    ext = []
    for t in range(1 << ek):
        x = zeta * pow(ext_omega, t, R) % R
        ext.append(sum(c * pow(x, i, R) for i, c in enumerate(coeff)) % R)
    # G1 and a 16-term multi-scalar multiplication
    d = 0x1D2C3B4A5968778695A4B3C2D1E0F
    G = (1, 2)
    bases = [ec_mul((1 + i * d) % R, G) for i in range(16)]
    scalars = field_elems(0x68326166, 16, R)
    scalars[3], scalars[4], scalars[5] = 0, 1, R - 1
    acc = None
    for s, b in zip(scalars, bases):
        acc = ec_add(acc, ec_mul(s, b))
    assert acc == ec_mul(sum(s * (1 + i * d) for i, s in enumerate(scalars)) % R, G)
    out = {
        "source": "first principles, see make_prover_side_golden.py",
        "r": hx(R), "q": hx(Q), "root_of_unity_2_28": hx(root), "zeta": hx(zeta), "delta": hx(delta),
        "domain": {"k": k, "j": j, "extended_k": ek, "omega": hx(omega), "extended_omega": hx(ext_omega),
                   "values": [hx(v) for v in a], "best_fft": [hx(v) for v in fft],
                   "lagrange_to_coeff": [hx(v) for v in coeff],
                   "coeff_to_extended_of_those_coefficients": [hx(v) for v in ext]},
        "g1": {"d": hx(d), "bases": [[hx(p[0]), hx(p[1])] for p in bases],
               "scalars": [hx(s) for s in scalars], "msm": [hx(acc[0]), hx(acc[1])],
               "two_g": [hx(v) for v in ec_add(G, G)], "minus_g": [hx(1), hx(Q - 2)]},
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
