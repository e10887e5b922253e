#!/usr/bin/env python
"""BASELINE config 4: 2^20 independent AES-128 blocks as independent circuits (K=20, N=5: 3853
blocks per circuit -> 273 circuits), dealt round-robin to the GPUs of one node.  No data-path
collective: every rank runs its own circuits (witness + A', S', Z of all 25 lookups) into reused
buffers; the only communication is the final MAX over ranks of the elapsed time.

    python examples/config4_batches.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29518 examples/config4_batches.py [--blocks 1048576] [--k 20] [--sets 5]

Ran on a B200 in round 2 (one GPU: 273 circuits in 498 ms with the Z[n-6] = 1 check of every
lookup after every circuit and one circuit in flight; `bench.py`'s `config4_stream` record is the
measured form of the same workload: two circuits in flight, 2.97 M blocks/s on one GPU, 23 M on
eight).  Checks Z[n-6] = 1 for every lookup of every circuit and prints one JSON line from rank 0.
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
    ap.add_argument("--blocks", type=int, default=1 << 20)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--sets", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from halo2_aes_b200 import _ffi, sharded
    from halo2_aes_b200.backend import Circuit, Context

    ctx = Context(local)
    cir = Circuit(ctx, _ffi.KIND_AES, args.k, args.sets)
    lay = cir.layout
    n, cap, L, nadv = lay["n_rows"], lay["capacity_blocks"], lay["n_lookups"], lay["n_advice"]
    circuits = sharded.split_into_circuits(args.blocks, cap)
    mine = sharded.circuits_of_rank(len(circuits), rank, world)
    th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
    advice = torch.empty((nadv, n, 4), dtype=torch.int64, device="cuda")
    bufs = {"Ap": torch.empty((L, n, 4), dtype=torch.int64, device="cuda"),
            "Sp": torch.empty((L, n, 4), dtype=torch.int64, device="cuda")}
    Z = torch.empty((L, n, 4), dtype=torch.int64, device="cuda")
    pts_all = torch.from_numpy(splitmix_bytes(SEED + 90 + rank, 16 * cap).reshape(cap, 16)).cuda()
    one = torch.tensor([-0x5369cbe3b0000005, 0x36fc76959f60cd29, 0x666ea36f7879462e,
                        0x0e0a77c19a07df2f], dtype=torch.int64, device="cuda")
    closed = torch.ones((), dtype=torch.bool, device="cuda")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    for ci in mine:
        first, count = circuits[ci]
        key = splitmix_bytes(SEED + 1000 + ci, 16).tobytes()   # one key per circuit
        th_i = th.copy()
        th_i[0] ^= np.uint64(ci + 1)
        cir.witness(key, pts_all[:count], out=advice, want_ciphertexts=False)
        cir.lookup_permuted(th_i, first=0, count=L, want_inputs=False, out=bufs)
        cir.lookup_product(be, ga, first=0, count=L, out=Z)
        closed &= (Z[:, n - 6] == one).all()
        done += count
    e1.record()
    torch.cuda.synchronize()
    cir.lookup_check()
    assert bool(closed), "a grand product does not close at 1"
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    total = torch.tensor([done], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total)
    if rank == 0:
        assert int(total) == args.blocks
        print(json.dumps({"config": f"{args.blocks} independent AES-128 blocks as {len(circuits)} "
                                    f"K={args.k},N={args.sets} circuits over {world} GPUs",
                          "n_gpus": world, "circuits": len(circuits), "blocks": int(total),
                          "ms": float(t), "blocks_per_s": int(total) / (float(t) * 1e-3),
                          "invariants": "Z[n-6]=1 for every lookup of every circuit: ok"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
