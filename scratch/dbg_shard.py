import sys, numpy as np, torch
sys.path.insert(0, '.')
import oracle as orc
from halo2_aes_b200 import _ffi as ffi
from halo2_aes_b200.backend import Circuit, Context, to_host_u64
from bench import splitmix_bytes, SEED
gpu = Context(0)
rng = np.random.default_rng(5)
theta, beta, gamma = (orc.fr_from_int(int.from_bytes(rng.bytes(31), 'little')) for _ in range(3))
k, N = 17, 2
pts = splitmix_bytes(SEED + 50, 16 * 120).reshape(-1, 16)
key = splitmix_bytes(SEED + 51, 16).tobytes()
d_pts = torch.from_numpy(pts).cuda()
ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(key, pts)
for G in (1, 2, 8):
    shards = [Circuit(gpu, ffi.KIND_AES, k, N) for _ in range(G)]
    for g, c in enumerate(shards):
        c.set_shard(g, G)
    adv = [c.witness(key, d_pts)[0] for c in shards]
    hists = [c.shard_hist(theta) for c in shards]
    total = torch.stack(hists).sum(dim=0, dtype=torch.int32)
    for h in hists:
        h.copy_(total)
    perm = [c.shard_permuted(want_inputs=False) for c in shards]
    zp = [c.shard_partial(beta, gamma) for c in shards]
    gathered = torch.stack([p.clone() for _, p in zp]).contiguous()
    Zs = [c.shard_product(gathered, z) for c, (z, _) in zip(shards, zp)]
    torch.cuda.synchronize()
    gZ = np.concatenate([to_host_u64(z) for z in Zs], axis=1)
    for li in range(10):
        r = ref.lookup(li, theta, beta, gamma)
        bad = np.nonzero((gZ[li] != r["Z"]).any(axis=1))[0]
        print("G", G, "lookup", li, "bad rows", len(bad), bad[:6], bad[-3:] if len(bad) else "")
