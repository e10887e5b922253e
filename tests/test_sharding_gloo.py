"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in halo2_aes_b200/sharded.py:
the circuit partitioning of independent batches, and the two collectives of the row-sharded
circuit, checked against the oracle's whole-circuit columns."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, SEED, seeded_fr, splitmix_bytes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from halo2_aes_b200 import sharded

        # ---- independent circuits (BASELINE config 4) -------------------------------------
        circuits = sharded.split_into_circuits(10000, 3853)
        assert circuits == [(0, 3853), (3853, 3853), (7706, 2294)]
        mine = sharded.circuits_of_rank(len(circuits), rank, world)
        counts = torch.tensor([sum(circuits[i][1] for i in mine)])
        dist.all_reduce(counts)
        assert int(counts) == 10000

        # ---- row-sharded single circuit (BASELINE config 5), emulated with the oracle -------
        k, u = 17, (1 << 17) - 6
        n = 1 << k
        th, be, ga = seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
        key = splitmix_bytes(SEED + 40, 16).tobytes()
        pts = splitmix_bytes(SEED + 41, 16 * 40).reshape(-1, 16)
        full = oracle.Circuit(oracle.KIND_AES, k, 1).synthesize(key, pts)
        lo, hi = rank * n // world, (rank + 1) * n // world
        L = oracle.lib()
        for li in (1, 2):
            r = full.lookup(li, th, be, ga)
            # exchange 1: histogram of the shard's input values over the distinct table values
            tab_vals, tab_inv = np.unique(r["S"][:u], axis=0, return_inverse=True)
            index = {v.tobytes(): i for i, v in enumerate(tab_vals)}
            hist = torch.zeros(len(tab_vals), dtype=torch.int32)
            for row in range(lo, min(hi, u)):
                hist[index[r["A"][row].tobytes()]] += 1
            sharded.allreduce_histogram(hist)
            whole = np.zeros(len(tab_vals), np.int64)
            for row in range(u):
                whole[index[r["A"][row].tobytes()]] += 1
            assert (hist.numpy() == whole).all()
            assert int(hist.sum()) == u
            # exchange 2: per-shard partial product of Z's factors, gathered, gives the carry-in
            Z = r["Z"]
            def fr_mul(a, b):
                out = np.zeros(4, np.uint64)
                L.orc_fr_mul(oracle._u64(np.ascontiguousarray(a)), oracle._u64(np.ascontiguousarray(b)),
                             oracle._u64(out))
                return out
            def fr_inv(a):
                out = np.zeros(4, np.uint64)
                L.orc_fr_inv(oracle._u64(np.ascontiguousarray(a)), oracle._u64(out))
                return out
            last = min(hi, u)  # Z[last] / Z[lo] = product of the shard's factors
            partial = fr_mul(Z[last], fr_inv(Z[lo]))
            t = torch.from_numpy(partial.view(np.int64).copy()).view(1, 4)
            gathered = sharded.allgather_partials(t)
            assert gathered.shape == (world, 1, 4)
            carry = oracle.fr_from_int(1)
            for s in range(rank):
                carry = fr_mul(carry, gathered[s, 0].numpy().view(np.uint64))
            assert (carry == Z[lo]).all(), "carry-in of the shard != Z at its first row"
        # ---- prover-side columns dealt to the ranks; only the commitments travel --------------
        ncols = 5
        n_pts = 64
        bases = oracle.g1_test_bases(n_pts, oracle.fr_from_int(12345))
        cols = np.stack([np.stack([oracle.fr_from_int((7 * c + 3 * i + 1) ** 3) for i in range(n_pts)])
                         for c in range(ncols)])
        mine = sharded.columns_of_rank(ncols, rank, world)
        got = np.stack([oracle.g1_msm(cols[c], bases) for c in mine]) if len(mine) else np.zeros((0, 2, 4), np.uint64)
        allc = sharded.allgather_commitments(torch.from_numpy(got.view(np.int64).copy()), ncols)
        want = np.stack([oracle.g1_msm(cols[c], bases) for c in range(ncols)])
        assert (allc.numpy().view(np.uint64) == want).all()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_partition_helpers():
    from halo2_aes_b200 import sharded
    assert sharded.split_into_circuits(0, 10) == []
    assert sharded.split_into_circuits(10, 10) == [(0, 10)]
    assert sharded.split_into_circuits(11, 10) == [(0, 10), (10, 1)]
    c = sharded.split_into_circuits(1 << 20, 3853)
    assert len(c) == 273 and sum(x[1] for x in c) == 1 << 20   # 273 circuits
    seen = sorted(i for r in range(8) for i in sharded.circuits_of_rank(273, r, 8))
    assert seen == list(range(273))
    with pytest.raises(ValueError):
        sharded.split_into_circuits(5, 0)
    # prover-side columns: contiguous blocks, sizes within one of each other, every column once
    for ncols, world in ((97, 8), (5, 2), (3, 8), (0, 4), (145, 3)):
        blocks = [sharded.columns_of_rank(ncols, r, world) for r in range(world)]
        assert [c for b in blocks for c in b] == list(range(ncols))
        sizes = [len(b) for b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.columns_of_rank(5, 2, 2)


def test_config4_dealing_covers_every_block_once():
    """BASELINE config 4 as bench.py's config4_stream runs it: 2^20 blocks = 272 circuits at the
    K=20, N=5 capacity + one of 560 blocks; the round-robin deal gives every circuit to exactly
    one rank."""
    from halo2_aes_b200 import sharded
    total, cap = 1 << 20, 3853
    circuits = sharded.split_into_circuits(total, cap)
    assert len(circuits) == 273 and circuits[-1] == (272 * cap, 560)
    assert sum(c for _, c in circuits) == total
    assert all(s == i * cap for i, (s, _) in enumerate(circuits))
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in sharded.circuits_of_rank(len(circuits), r, world))
        assert seen == list(range(273))
        sizes = [len(sharded.circuits_of_rank(273, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_comm_entry_points_reject_bad_arguments():
    """h2aes_comm_* argument checking needs neither a GPU nor NCCL"""
    import ctypes as C
    from halo2_aes_b200 import _ffi
    L = _ffi.lib()
    assert L.h2aes_comm_unique_id(None) == _ffi.ERR_ARG
    buf = (C.c_uint8 * _ffi.COMM_ID_BYTES)()
    assert L.h2aes_comm_init(None, buf, 0, 1) == _ffi.ERR_ARG
    assert L.h2aes_comm_destroy(None) == _ffi.ERR_ARG
    assert L.h2aes_lookup_columns_sharded_dev(None, None, None, None, None, None, None) == _ffi.ERR_ARG
