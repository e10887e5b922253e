"""Multi-GPU partitioning of the path (one process per GPU, torch.distributed for plumbing).

Two regimes (BASELINE configs 4 and 5):

* independent circuits / AES batches: blocks are cut into whole circuits of the reference's
  capacity and dealt round-robin to the ranks.  No data-path collective at all.
* one large circuit, row-sharded: shard g of G owns rows [g*n/G, (g+1)*n/G) of every column.
  The witness needs no exchange (every rank expands the key and the block states itself and
  stores only its rows).  The lookup helper columns need exactly two small exchanges:
    1. all-reduce(sum) of the per-lookup table-bin histogram (66568 u32 per lookup) before
       A'/S' can be expanded, because the permuted columns are a global sort;
    2. all-gather of one 32-byte partial product per shard per lookup, from which each shard
       derives its carry-in for the grand product Z.
  Both are KiB..MiB-sized and latency-bound on NVLink5/NVSwitch, so they go through NCCL as
  they are; the bulk columns never leave the GPU that owns their rows.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------
# independent circuits
# ---------------------------------------------------------------------------------------------
def split_into_circuits(n_blocks: int, capacity: int) -> list[tuple[int, int]]:
    """[(first_block, count)] of the circuits needed for n_blocks AES blocks, each filled to the
    reference's capacity rule (aes128.rs:303-325) except the last."""
    if capacity <= 0:
        raise ValueError("capacity must be positive")
    return [(s, min(capacity, n_blocks - s)) for s in range(0, n_blocks, capacity)]


def circuits_of_rank(n_circuits: int, rank: int, world: int) -> list[int]:
    """round-robin deal of circuit indices to ranks"""
    return list(range(rank, n_circuits, world))


# ---------------------------------------------------------------------------------------------
# collectives of the row-sharded circuit (device-agnostic: NCCL on GPUs, gloo in the CPU tests)
# ---------------------------------------------------------------------------------------------
def allreduce_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """in-place SUM of the [n_lookups * BIN_STRIDE] int32 bin histogram across shards"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def allgather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """[n_lookups, 4] int64 per shard -> [n_shards, n_lookups, 4] on every shard"""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return partial.unsqueeze(0).contiguous()
    world = dist.get_world_size(group)
    flat = partial.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view((world,) + tuple(partial.shape))


# ---------------------------------------------------------------------------------------------
# prover-side steps of one circuit (commitments, domain transforms): the columns are
# independent of one another, so they are dealt to the ranks in contiguous blocks and only the
# 64-byte commitments travel
# ---------------------------------------------------------------------------------------------
def columns_of_rank(n_columns: int, rank: int, world: int) -> range:
    """contiguous block of column indices owned by `rank` (sizes differ by at most one)"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    base, extra = divmod(n_columns, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def allgather_commitments(mine: torch.Tensor, n_columns: int, group=None) -> torch.Tensor:
    """[len(columns_of_rank), 2, 4] int64 commitments of this rank's columns -> [n_columns, 2, 4]
    on every rank, in column order (ranks own contiguous blocks; padded to equal length for the
    collective)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        assert mine.shape[0] == n_columns
        return mine
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_columns + world - 1) // world
    assert mine.shape[0] == len(columns_of_rank(n_columns, rank, world))
    pad = torch.zeros((per, 2, 4), dtype=mine.dtype, device=mine.device)
    pad[: mine.shape[0]] = mine
    out = torch.empty((world, per, 2, 4), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return torch.cat([out[r, : len(columns_of_rank(n_columns, r, world))] for r in range(world)])


def library_comm_init(ctx, group=None):
    """Give `ctx` the library's own NCCL communicator (h2aes_comm_init) spanning `group`: rank 0
    draws the unique id, torch.distributed (any backend) ships its 128 bytes."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    box = [ctx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group else 0, group=group)
    ctx.comm_init(box[0], rank, world)


class RowShardedCircuit:
    """One circuit spread over the ranks of `group` by rows.

    library_comm=False: the two exchanges go through torch.distributed on the buffers the phase
    entry points expose.  library_comm=True: one call, h2aes_lookup_columns_sharded_dev, with the
    exchanges done inside the library over its own NCCL communicator (library_comm_init first)."""

    def __init__(self, ctx, kind: int, k: int, N: int, group=None, library_comm: bool = False,
                 tiled: bool = False):
        from .backend import Circuit
        self.group = group
        self.library_comm = library_comm
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.circuit = Circuit(ctx, kind, k, N)
        self.circuit.set_shard(self.rank, self.world, tiled=tiled)

    @property
    def row_range(self):
        """(first, end) row of a contiguous shard"""
        assert not self.circuit.tiled, "a tiled shard holds several row ranges: Circuit.global_rows()"
        per = self.circuit.n // self.world
        return per * self.rank, per * (self.rank + 1)

    def witness(self, key, plaintexts=None, out=None):
        """every rank passes the same key/plaintexts; returns this shard's rows of every advice
        column ([n_advice, n/G, 4]).  No communication."""
        return self.circuit.witness(key, plaintexts, out=out)

    def fixed_columns(self, **kw):
        return self.circuit.fixed_columns(**kw)

    def lookup_columns(self, theta, beta, gamma, want_inputs=False, out=None):
        c = self.circuit
        if self.library_comm:
            assert not want_inputs, "A, S are local data: use Circuit.shard_permuted for them"
            return c.lookup_columns_sharded(theta, beta, gamma, out=out)
        hist = c.shard_hist(theta)
        allreduce_histogram(hist, self.group)                 # exchange 1
        cols = c.shard_permuted(want_inputs=want_inputs)
        Z, partial = c.shard_partial(beta, gamma)
        gathered = allgather_partials(partial, self.group)    # exchange 2
        cols["Z"] = c.shard_product(gathered, Z)
        return cols
