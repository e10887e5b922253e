"""Commitment time of the columns of ONE K=20, N=5 circuit by kind (advice, A', S', Z, Z_perm),
directly and through the differences (h2aes_msm_enable_runs).  CUDA events, warm.  One JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo2_aes_b200 import _ffi  # noqa: E402
from halo2_aes_b200.backend import Circuit, Context, Msm, synthetic_bases  # noqa: E402
from profiles.prover_side import fr  # noqa: E402
import numpy as np  # noqa: E402

K, N = 20, 5
ctx = Context(0)
cir = Circuit(ctx, _ffi.KIND_AES, K, N)
nb = cir.layout["capacity_blocks"]
rng = np.random.default_rng(0x68326165)
key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
pts = torch.from_numpy(rng.integers(0, 256, (nb, 16), dtype=np.uint8)).cuda()
advice, _ = cir.witness(key, pts)
perm = cir.lookup_permuted(fr(1), want_inputs=False)
Z = cir.lookup_product(fr(2), fr(3))
Zp = cir.perm_product(fr(2), fr(3), advice)
msm = Msm(ctx, synthetic_bases(ctx, 1 << K, fr(5)))


def timed(cols, reps=2):
    msm.commit(cols[:1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for _ in range(reps):
        out = msm.commit(cols)
    e1.record(ctx.stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / cols.shape[0], out


kinds = {"advice": advice, "Ap": perm["Ap"], "Sp": perm["Sp"], "Z": Z, "Zperm": Zp}
res = {"direct_ms_per_column": {}, "runs_ms_per_column": {}}
direct = {}
for nm, t in kinds.items():
    res["direct_ms_per_column"][nm], direct[nm] = timed(t)
msm.enable_runs()
for nm, t in kinds.items():
    res["runs_ms_per_column"][nm], out = timed(t)
    assert torch.equal(out, direct[nm]), nm
allc = torch.cat([advice, perm["Ap"], perm["Sp"], Z])
ctx.profile(True)
ms, _ = timed(allc, reps=1)
res["committed_set_91_columns_ms"] = ms * allc.shape[0]
res["kernels"] = {k: round(v[1], 3) for k, v in ctx.profile_read().items()}
res["env"] = {k: v for k, v in os.environ.items() if k.startswith("H2AES_MSM")}
print(json.dumps(res))
