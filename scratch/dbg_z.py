import sys, numpy as np, torch
sys.path.insert(0, '.')
import oracle as orc
from halo2_aes_b200 import _ffi as ffi
from halo2_aes_b200.backend import Circuit, Context, to_host_u64
from bench import splitmix_bytes, SEED
gpu = Context(0)
rng = np.random.default_rng(5)
theta, beta, gamma = (orc.fr_from_int(int.from_bytes(rng.bytes(31), 'little')) for _ in range(3))
k = 17
key = splitmix_bytes(SEED + 51, 16).tobytes()
cir = Circuit(gpu, ffi.KIND_KEYSCHED, k, 1)
cir.witness(key, None)
perm = cir.lookup_permuted(theta)
Z = cir.lookup_product(beta, gamma)
torch.cuda.synchronize()
hZ = to_host_u64(Z)
ref = orc.Circuit(orc.KIND_KEYSCHED, k, 1).synthesize(key, None)
for li in range(3):
    r = ref.lookup(li, theta, beta, gamma)
    bad = np.nonzero((hZ[li] != r["Z"]).any(axis=1))[0]
    print("lookup", li, "bad rows", len(bad), bad[:12], bad[-4:] if len(bad) else "")
    if len(bad):
        b0 = bad[0]
        print("  got", [hex(int(x)) for x in hZ[li, b0]], "\n  ref", [hex(int(x)) for x in r["Z"][b0]])
        m = np.nonzero((r["Z"] == hZ[li, b0]).all(axis=1))[0]
        print("  value found at ref rows", m[:5])
