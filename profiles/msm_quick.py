"""Quick timing of the G1 multi-scalar multiplication on one GPU (CUDA events, warm)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (only to build a valid set of bases: test infrastructure)
from halo2_aes_b200.backend import Context, Msm  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
ctx = Context(0)
B = oracle.g1_test_bases(n, oracle.fr_from_int(0x1D2C3B4A5968778695A4B3C2D1E0F))
msm = Msm(ctx, torch.from_numpy(B.view(np.int64)).cuda())
g = torch.Generator(device="cuda").manual_seed(1)
rnd = torch.randint(0, 2**62, (4, n, 4), dtype=torch.int64, device="cuda", generator=g)
rnd[..., 3] &= (1 << 60) - 1
one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f],
                   dtype=torch.int64, device="cuda")
sel = torch.zeros((4, n, 4), dtype=torch.int64, device="cuda")
sel[:, ::3] = one  # 0/1 column (selector-like)


def timed(cols, reps=3):
    msm.commit(cols)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        msm.commit(cols)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / cols.shape[0]


ctx.profile(True)
res = {"k": k, "uniform_ms_per_column": timed(rnd), "kernels_uniform": ctx.profile_read()}
res["zero_one_ms_per_column"] = timed(sel)
res["kernels_zero_one"] = ctx.profile_read()
print(json.dumps(res))
