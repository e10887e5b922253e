#!/usr/bin/env python
"""BASELINE config 5: ONE K=22 circuit of batched AES blocks, row-sharded over the GPUs of one
node, the lookup grand products stitched with two small NCCL exchanges (histogram all-reduce,
partial-product all-gather).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29517 examples/config5_sharded.py [--k 22] [--sets 5] [--iters 5]

Checks size-independent properties on the device (Z[0] = 1 on shard 0, Z[n-6] = 1 on the last
shard, permuted-pair constraints inside every shard) and prints one JSON line from rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SEED, seeded_fr, splitmix_bytes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=22)
    ap.add_argument("--sets", type=int, default=5)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    if "RANK" not in os.environ:  # plain `python examples/config5_sharded.py`: one shard
        os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29519")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from halo2_aes_b200 import _ffi, sharded
    from halo2_aes_b200.backend import Context

    ctx = Context(local)
    cir = sharded.RowShardedCircuit(ctx, _ffi.KIND_AES, args.k, args.sets)
    lay = cir.circuit.layout
    n, nb, L = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"]
    th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
    key = splitmix_bytes(SEED + 80, 16).tobytes()
    pts = torch.from_numpy(splitmix_bytes(SEED + 81, 16 * nb).reshape(-1, 16)).cuda()
    one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e,
                        0x0e0a77c19a07df2f], dtype=torch.int64, device="cuda")
    times = []
    for it in range(args.iters + 1):
        torch.cuda.synchronize()
        dist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        advice, _ = cir.witness(key, pts)
        e[1].record()
        th_i = th.copy()
        th_i[0] ^= np.uint64(it + 1)
        cols = cir.lookup_columns(th_i, be, ga, want_inputs=False)
        e[2].record()
        torch.cuda.synchronize()
        cir.circuit.lookup_check()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it:
            times.append(t.tolist())
        lo, hi = cir.row_range
        ok = True
        if rank == 0:
            ok &= bool((cols["Z"][:, 0] == one).all())
        if lo <= n - 6 < hi:
            ok &= bool((cols["Z"][:, n - 6 - lo] == one).all())
            ok &= not bool(cols["Z"][:, n - 5 - lo:].any())
        # (A'_i - S'_i)(A'_i - A'_{i-1}) = 0 on the shard's usable rows
        Ap, Sp = cols["Ap"], cols["Sp"]
        u = min(hi, n - 6) - lo
        same_s = (Ap[:, 1:u] == Sp[:, 1:u]).all(dim=-1)
        same_prev = (Ap[:, 1:u] == Ap[:, :u - 1]).all(dim=-1)
        ok &= bool((same_s | same_prev).all())
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag) == 1, "sharded lookup columns violate an invariant"
        del advice, cols
    if rank == 0:
        w = float(np.mean([t[0] for t in times])), float(np.mean([t[1] for t in times]))
        print(json.dumps({"config": f"one K={args.k}, N={args.sets} circuit, {nb} blocks, "
                                    f"row-sharded over {world} GPUs",
                          "n_gpus": world, "blocks": nb, "witness_ms": w[0], "lookup_columns_ms": w[1],
                          "blocks_per_s": nb / ((w[0] + w[1]) * 1e-3),
                          "bytes_per_gpu": (lay["n_advice"] + 3 * L) * (n // world) * 32,
                          "invariants": "Z[0]=1, Z[n-6]=1, permuted-pair constraint: ok"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
