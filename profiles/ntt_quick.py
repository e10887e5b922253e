"""Quick timing of the evaluation-domain transforms on one GPU (CUDA events, warm, inputs in HBM)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from halo2_aes_b200.backend import Context, Domain  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = Context(0)
dom = Domain(ctx, k)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(0, 2**62, (ncols, dom.n, 4), dtype=torch.int64, device="cuda", generator=g)
a[..., 3] &= (1 << 60) - 1  # < r
ext = torch.empty((ncols, dom.extended_n, 4), dtype=torch.int64, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ctx.profile(True)
res = {}
for name, fn, K in (("lagrange_to_coeff", lambda: dom.lagrange_to_coeff(a), k),
                    ("coeff_to_extended", lambda: dom.coeff_to_extended(a, ext), dom.extended_k),
                    ("extended_to_coeff", lambda: dom.extended_to_coeff(ext), dom.extended_k)):
    ms = timed(fn)
    N = 1 << K
    muls = ncols * (N // 2 * (K - 2) + N + (N if name != "coeff_to_extended" else dom.n))
    res[name] = {"ms": ms, "ms_per_column": ms / ncols, "fr_multiplies": muls,
                 "gmul_s": muls / ms / 1e6, "hbm_gbs": ncols * 4 * 32 * N / ms / 1e6}
res["kernels"] = ctx.profile_read()
print(json.dumps({"k": k, "ncols": ncols, **res}))
