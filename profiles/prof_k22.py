import sys, numpy as np, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from halo2_aes_b200 import _ffi as ffi
from halo2_aes_b200.backend import Circuit, Context
from bench import splitmix_bytes, SEED, seeded_fr
gpu = Context(0)
th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
k, N = 22, 5
cir = Circuit(gpu, ffi.KIND_AES, k, N)
nb = cir.layout["capacity_blocks"]
pts = torch.from_numpy(splitmix_bytes(SEED + 1, 16 * nb).reshape(-1, 16)).cuda()
key = splitmix_bytes(SEED, 16).tobytes()
out = None
adv = None
for it in range(3):
    th_i = th.copy(); th_i[0] ^= np.uint64(it + 5)
    if it == 2:
        gpu.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    adv, _ = cir.witness(key, pts, out=adv, want_ciphertexts=False)
    out = cir.lookup_columns_sharded(th_i, be, ga, out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("iter", it, "cpu enqueue ms", round(1e3 * (t1 - t0), 3), "total ms", round(1e3 * (t2 - t0), 3))
prof = gpu.profile_read()
tot = 0
for kname, (cnt, ms) in prof.items():
    tot += ms
    print(f"  {kname:20s} {cnt:3d} {ms:8.4f}")
print("sum of kernels", round(tot, 3))
