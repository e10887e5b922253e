import sys, numpy as np, torch
sys.path.insert(0, '.')
import oracle as orc
from halo2_aes_b200 import _ffi as ffi
from halo2_aes_b200.backend import Circuit, Context
from bench import splitmix_bytes, SEED
gpu = Context(0)
cir = Circuit(gpu, ffi.KIND_AES, 20, 5)
nb = cir.layout["capacity_blocks"]
pts = torch.from_numpy(splitmix_bytes(SEED + 1, 16 * nb).reshape(-1, 16)).cuda()
cir.witness(splitmix_bytes(SEED, 16).tobytes(), pts)
one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=torch.int64, device="cuda")
be, ga = orc.fr_from_int(12345678901234567), orc.fr_from_int(98765432109876543)
bufs = None
for name, th in (("full", int.from_bytes(np.random.default_rng(1).bytes(31), "little")), ("2^61", (1 << 61) + 12345), ("2^20", (1 << 20) + 7), ("3", 3), ("1", 1)):
    theta = orc.fr_from_int(th)
    bufs = cir.lookup_permuted(theta, want_inputs=False, out=bufs)   # warm
    torch.cuda.synchronize()
    gpu.profile(True)
    th2 = orc.fr_from_int(th)  # same theta: force rebuild by toggling
    cir.witness(splitmix_bytes(SEED, 16).tobytes(), pts)
    bufs = cir.lookup_permuted(th2, want_inputs=False, out=bufs)
    Z = cir.lookup_product(be, ga)
    prof = gpu.profile_read()
    gpu.profile(False)
    cir.lookup_check()
    ok = bool((Z[:, (1 << 20) - 6] == one).all())
    print(name, "Z closes:", ok, {k: round(v[1], 4) for k, v in prof.items() if k in ("k_splitters", "k_bucket_hist", "k_bucket_rank", "k_permuted", "k_bin_fill")})
