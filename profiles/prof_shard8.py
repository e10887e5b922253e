"""Per-kernel times of ONE shard of a K=22,N=5 circuit row-sharded 8 ways (tiled), with the 8 shards
emulated on one GPU and the exchanges done by hand: where the 1.1 ms of the 8-GPU run goes."""
import sys, numpy as np, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from halo2_aes_b200 import _ffi as ffi
from halo2_aes_b200.backend import Circuit, Context
from bench import splitmix_bytes, SEED, seeded_fr
gpu = Context(0)
th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
import os
k, N, G = 22, 5, int(os.environ.get("G", "8"))
TILED = os.environ.get("TILED", "1") == "1"
shards = [Circuit(gpu, ffi.KIND_AES, k, N) for _ in range(G)]
for g, c in enumerate(shards):
    c.set_shard(g, G, tiled=TILED)
nb = shards[0].layout["capacity_blocks"]
pts = torch.from_numpy(splitmix_bytes(SEED + 1, 16 * nb).reshape(-1, 16)).cuda()
key = splitmix_bytes(SEED, 16).tobytes()
allk = []
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        gpu.profile(True)
    phases = {}
    def mark(name, fn, only0=True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); phases[name] = round(1e3 * (time.perf_counter() - t0), 3)
        if it == 1:
            for kn, (cnt, ms) in gpu.profile_read().items():
                allk.append((name, kn, cnt, ms))
        return r
    adv0 = mark("witness0", lambda: shards[0].witness(key, pts, want_ciphertexts=False))
    gpu.profile(False) if it == 1 else None
    for c in shards[1:]:
        c.witness(key, pts, want_ciphertexts=False)
    gpu.profile(True) if it == 1 else None
    h0 = mark("hist0", lambda: shards[0].shard_hist(th))
    gpu.profile(False) if it == 1 else None
    hists = [h0] + [c.shard_hist(th) for c in shards[1:]]
    total = torch.stack(hists).sum(dim=0, dtype=torch.int32)
    for h in hists:
        h.copy_(total)
    gpu.profile(True) if it == 1 else None
    p0 = mark("permuted0", lambda: shards[0].shard_permuted(want_inputs=False))
    zp0 = mark("partial0", lambda: shards[0].shard_partial(be, ga))
    gpu.profile(False) if it == 1 else None
    zps = [zp0] + [(c.shard_permuted(want_inputs=False), c.shard_partial(be, ga))[1] for c in shards[1:]]
    gathered = torch.stack([p.clone() for _, p in zps]).contiguous()
    gpu.profile(True) if it == 1 else None
    Z0 = mark("product0", lambda: shards[0].shard_product(gathered, zp0[0]))
    shards[0].lookup_check()
    print("iter", it, phases, "sum", round(sum(phases.values()), 3))
tot = 0
for ph, kname, cnt, ms in allk:
    tot += ms
    print(f"  {ph:10s} {kname:20s} {cnt:3d} {ms:8.4f}")
print("sum of kernels", round(tot, 3))
