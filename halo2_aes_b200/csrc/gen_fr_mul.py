#!/usr/bin/env python
"""Generator + CPU proof for the device Montgomery multiply of bn256 Fr (fr_mul_ptx.inc).

The multiply is an operand-scanning Montgomery product on 8x32-bit limbs whose partial products
are kept in TWO accumulators, one holding the 64-bit products that start at even limb positions
and one holding those that start at odd positions.  Inside each accumulator the lo/hi halves of
consecutive products land on consecutive limbs, so a whole row `acc += a[j..]*b_i` is ONE
uninterrupted carry chain of mad.lo.cc / madc.hi.cc (ptxas turns each lo/hi pair into one
IMAD.WIDE.U32 with carry), instead of the carry save/restore a single accumulator needs.
After every row the running value is divided by 2^32, which swaps the roles of the two
accumulators.

This script is the executable specification: it builds the instruction list once, runs it on
a Python model of the PTX carry flag against big-integer arithmetic (random and extreme
operands), and only then prints the same list as one inline-asm block.

    python gen_fr_mul.py            # self-test, then (re)write fr_mul_ptx.inc
    python gen_fr_mul.py fq         # the same for the base field Fq -> fq_mul_ptx.inc
"""
import os
import random
import sys

FIELDS = {
    # bn256 scalar field (halo2curves bn256::Fr) and base field (bn256::Fq)
    "fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "fq": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
}
FIELD = sys.argv[1] if len(sys.argv) > 1 else "fr"
P = FIELDS[FIELD]
N = 8
M32 = 0xFFFFFFFF
PL = [(P >> (32 * i)) & M32 for i in range(N)]
INV = (-pow(P, -1, 1 << 32)) % (1 << 32)  # -p^-1 mod 2^32 = 0xefffffff


def build():
    """Instruction list.  Register names: a0..a7, b0..b7 inputs; e0..e7 / o0..o7 accumulators;
    m scratch; r0..r7 outputs.  Constants are ('imm', value)."""
    ops = []
    E = [f"e{i}" for i in range(N)]
    O = [f"o{i}" for i in range(N)]
    A = [f"a{i}" for i in range(N)]
    B = [f"b{i}" for i in range(N)]
    Pc = [("imm", v) for v in PL]

    def mul_row(acc, src, s0, bi):  # acc = src[s0], src[s0+2].. * bi, no carries involved
        for j in range(0, N, 2):
            ops.append(("mul.lo", acc[j], src[s0 + j], bi))
            ops.append(("mul.hi", acc[j + 1], src[s0 + j], bi))

    def cmad_row(acc, src, s0, bi, carry_in=False):  # acc += src[s0,s0+2,..]*bi, one chain
        for j in range(0, N, 2):
            first = j == 0 and not carry_in
            ops.append(("mad.lo.cc" if first else "madc.lo.cc", acc[j], src[s0 + j], bi, acc[j]))
            ops.append(("madc.hi.cc", acc[j + 1], src[s0 + j], bi, acc[j + 1]))

    def cmad_row_shift(acc, src, s0, bi):
        # acc[j] = acc[j+2] + product, consuming the pending carry: the accumulator slides down
        # by two limbs (its two lowest limbs left through the other accumulator)
        for j in range(0, N - 2, 2):
            ops.append(("madc.lo.cc", acc[j], src[s0 + j], bi, acc[j + 2]))
            ops.append(("madc.hi.cc", acc[j + 1], src[s0 + j], bi, acc[j + 3]))
        ops.append(("madc.lo.cc", acc[N - 2], src[s0 + N - 2], bi, ("imm", 0)))
        ops.append(("madc.hi", acc[N - 1], src[s0 + N - 2], bi, ("imm", 0)))

    def reduce_row(x, y):
        # x is aligned at limb 0, y at limb 1: add m*p so that limb 0 becomes zero
        ops.append(("mul.lo", "m", x[0], ("imm", INV)))
        cmad_row(y, Pc, 1, "m")           # top product p7*m + a7*b < 2^63: no carry out
        cmad_row(x, Pc, 0, "m")
        ops.append(("addc", y[N - 1], y[N - 1], ("imm", 0)))  # carry out of limb 8 -> y[7]

    x, y = E, O
    for i in range(N):
        if i == 0:
            mul_row(y, A, 1, B[0])
            mul_row(x, A, 0, B[0])
        else:
            # the previous row left limb 0 of the old x zero; dividing by 2^32 makes the old y
            # the limb-0 accumulator.  old x[1] joins it; old x[2..] becomes the limb-1 one.
            ops.append(("add.cc", x[0], x[0], y[1]))
            cmad_row_shift(y, A, 1, B[i])
            cmad_row(x, A, 0, B[i])
            ops.append(("addc", y[N - 1], y[N - 1], ("imm", 0)))
        reduce_row(x, y)
        x, y = y, x
    # after 8 rows x == E again: result = x + (y >> 32 limbs) with the stray y[1]
    ops.append(("add.cc", "r0", x[0], y[1]))
    for j in range(1, N - 1):
        ops.append(("addc.cc", f"r{j}", x[j], y[j + 1]))
    ops.append(("addc", f"r{N - 1}", x[N - 1], ("imm", 0)))
    return ops


class Overflow(Exception):
    pass


def run(ops, a, b):
    reg = {}
    for i in range(N):
        reg[f"a{i}"] = (a >> (32 * i)) & M32
        reg[f"b{i}"] = (b >> (32 * i)) & M32
    cc = 0

    def val(x):
        return x[1] if isinstance(x, tuple) else reg[x]

    for op in ops:
        name, d = op[0], op[1]
        s = [val(x) for x in op[2:]]
        if name == "mul.lo":
            reg[d] = (s[0] * s[1]) & M32
        elif name == "mul.hi":
            reg[d] = (s[0] * s[1]) >> 32
        elif name in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc", "madc.hi"):
            prod = s[0] * s[1]
            part = (prod >> 32) if ".hi" in name else (prod & M32)
            t = part + s[2] + (cc if name.startswith("madc") else 0)
            reg[d] = t & M32
            if name.endswith(".cc"):
                cc = t >> 32
            elif t >> 32:
                raise Overflow(op)
        elif name in ("add.cc", "addc.cc", "addc"):
            t = s[0] + s[1] + (cc if name.startswith("addc") else 0)
            reg[d] = t & M32
            if name.endswith(".cc"):
                cc = t >> 32
            elif t >> 32:
                raise Overflow(op)
        else:
            raise ValueError(name)
    return sum(reg[f"r{i}"] << (32 * i) for i in range(N))


def selftest(ops, iters=20000):
    rinv = pow(1 << 256, -1, P)
    rng = random.Random(0x68326165)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 254) % P, (1 << 256) % P, P >> 1, 0xFFFFFFFF,
            (P - 1) & ~M32, sum(M32 << (32 * i) for i in range(7)) % P]
    cases = [(x, y) for x in edge for y in edge]
    cases += [(rng.randrange(P), rng.randrange(P)) for _ in range(iters)]
    for a, b in cases:
        got = run(ops, a, b)
        assert got < 2 * P, "result not below 2p"
        assert got % P == a * b * rinv % P, (hex(a), hex(b))
    return len(cases)


def operand(x):
    if isinstance(x, tuple):
        return f"0x{x[1]:08x}"
    if x[0] == "a":
        return f"%{8 + int(x[1:])}"
    if x[0] == "b":
        return f"%{16 + int(x[1:])}"
    if x[0] == "r":
        return f"%{int(x[1:])}"
    return x  # e*, o*, m: asm-local registers


def emit(ops):
    lines = [f"// GENERATED by gen_fr_mul.py {FIELD} (self-tested against big-integer arithmetic) — do not edit.",
             "// Montgomery product r = a*b*2^-256 mod p on 8x32-bit limbs, r < 2p (caller reduces once).",
             "// Operands: %0-%7 = r (out), %8-%15 = a, %16-%23 = b.",
             "asm(\"{\\n\\t\"",
             "    \".reg .u32 e0,e1,e2,e3,e4,e5,e6,e7,o0,o1,o2,o3,o4,o5,o6,o7,m;\\n\\t\""]
    for op in ops:
        name, d = op[0], op[1]
        src = list(op[2:])
        if len(src) >= 2 and isinstance(src[0], tuple) and not isinstance(src[1], tuple):
            src[0], src[1] = src[1], src[0]  # keep immediates out of the first multiplicand slot
        srcs = ", ".join(operand(x) for x in src)
        lines.append(f"    \"{name}.u32 {operand(d)}, {srcs};\\n\\t\"")
    lines.append("    \"}\"")
    lines.append("    : \"=r\"(r0), \"=r\"(r1), \"=r\"(r2), \"=r\"(r3), \"=r\"(r4), \"=r\"(r5), \"=r\"(r6), \"=r\"(r7)")
    lines.append("    : \"r\"(a.v[0]), \"r\"(a.v[1]), \"r\"(a.v[2]), \"r\"(a.v[3]), \"r\"(a.v[4]), \"r\"(a.v[5]), \"r\"(a.v[6]), \"r\"(a.v[7]),")
    lines.append("      \"r\"(b.v[0]), \"r\"(b.v[1]), \"r\"(b.v[2]), \"r\"(b.v[3]), \"r\"(b.v[4]), \"r\"(b.v[5]), \"r\"(b.v[6]), \"r\"(b.v[7]));")
    return "\n".join(lines) + "\n"


def main():
    ops = build()
    n = selftest(ops)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{FIELD}_mul_ptx.inc")
    with open(out, "w") as f:
        f.write(emit(ops))
    print(f"self-test ok on {n} operand pairs; {len(ops)} PTX instructions -> {out}")


if __name__ == "__main__":
    sys.exit(main())
