"""Row-sharded circuit over real NCCL (BASELINE config 5 at test size): one process per GPU,
`sharded.RowShardedCircuit` drives the shard entry points and the two exchanges.  Needs >= 2
GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, SEED, seeded_fr, splitmix_bytes

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, N, nb, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from halo2_aes_b200 import _ffi, sharded
        from halo2_aes_b200.backend import Context, to_host_u64
        th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
        key = splitmix_bytes(SEED + 60, 16).tobytes()
        pts = splitmix_bytes(SEED + 61, 16 * nb).reshape(-1, 16)
        ctx = Context(rank)
        cir = sharded.RowShardedCircuit(ctx, _ffi.KIND_AES, k, N)
        advice, _ = cir.witness(key, torch.from_numpy(pts).cuda())
        cols = cir.lookup_columns(th, be, ga, want_inputs=True)
        cir.circuit.lookup_check()
        lo, hi = cir.row_range
        out = {"rows": (lo, hi), "advice": to_host_u64(advice)}
        for nm in ("A", "S", "Ap", "Sp", "Z"):
            out[nm] = to_host_u64(cols[nm])
        ret[rank] = out
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_row_sharded_over_nccl(orc):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 8)
    world = 1 << (world.bit_length() - 1)  # power of two shards
    k, N, nb = 17, 2, 120
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), k, N, nb, ret), nprocs=world, join=True)
    th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
    ref = orc.Circuit(orc.KIND_AES, k, N).synthesize(
        splitmix_bytes(SEED + 60, 16).tobytes(), splitmix_bytes(SEED + 61, 16 * nb).reshape(-1, 16))
    parts = [ret[r] for r in range(world)]
    assert [p["rows"] for p in parts] == [(r * (1 << k) // world, (r + 1) * (1 << k) // world)
                                         for r in range(world)]
    cat = lambda nm: np.concatenate([p[nm] for p in parts], axis=1)
    assert (cat("advice") == ref.advice()).all()
    got = {nm: cat(nm) for nm in ("A", "S", "Ap", "Sp", "Z")}
    for li in range(10):
        r = ref.lookup(li, th, be, ga)
        for nm in ("A", "S", "Ap", "Sp", "Z"):
            assert (got[nm][li] == r[nm]).all(), (nm, li)
