"""Quick timing of the G1 multi-scalar multiplication on one GPU (CUDA events, warm)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo2_aes_b200.backend import Context, Msm, synthetic_bases  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
ctx = Context(0)
one_limbs = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f],
                     dtype=np.uint64)  # d = 1 in Montgomery form: bases (1 + i) * G
msm = Msm(ctx, synthetic_bases(ctx, n, one_limbs))
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
