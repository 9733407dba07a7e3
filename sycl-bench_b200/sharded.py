"""Multi-GPU data path: one process per GPU, contiguous shards, ONE tiny exchange per operation.

    reduction : "fused": sb_reduce_all — the last block of the reduce kernel exchanges the per-GPU sums through the
                         peer-memory mailboxes (NVLink) and adds them in rank order: one launch, no collective call
                "nccl" : local sb_reduce -> all-reduce(sum) of one element per rank     (traffic N*s/P per GPU)
    prefix sum: "reduce_then_scan": local sb_reduce -> all-gather of the P shard totals
                                    -> sb_exclusive_carry -> sb_scan seeded with the carry  (3*N*s/P per GPU)
                "scan_then_carry" : local sb_scan -> all-gather of the shard totals
                                    -> sb_exclusive_carry -> sb_carry_add                    (4*N*s/P per GPU,
                                    the literal BASELINE.json north_star recipe)
                "cyclic"          : BLOCK-CYCLIC shards (block = one round of the persistent kernel, 10.1 MB);
                                    sb_scan_cyclic on every rank: the round totals cross NVLink inside the kernel
                                    (peer-memory mailboxes), nothing is read twice              (2*N*s/P per GPU,
                                    no torch.distributed call in the data path — only the one-off exchange of
                                    the mailbox handles)
    segmented reduction / segmented scan / group-scope scan: shards cut on segment boundaries are
    independent — no collective at all (`foreach_*`).

The class only sequences local operations and torch.distributed collectives; the local operations
are passed in (libsb_cuda through sbcuda.Context on the GPU; the CPU tests pass the oracle, which is
how the world_size-2 gloo tests cover this logic without a GPU).
"""
from __future__ import annotations

from typing import Callable


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous shard [begin, end) of rank; every shard but the last is a multiple of `align`
    (segment length for the segmented operations, 1 otherwise)."""
    per = -(-n // world)
    per = -(-per // align) * align
    begin = min(n, rank * per)
    end = min(n, begin + per)
    return begin, end


def cyclic_global_index(i: int, n: int, block: int, world: int, rank: int) -> int:
    """Global index of local element i of `rank` in the block-cyclic layout of sb_scan_cyclic (include/sb_cuda.h):
    every rank holds n elements; block k of `block` elements lives on rank k % world; the blocks of the last
    round are all n % block long and follow each other in rank order."""
    full, last = divmod(n, block)
    b, j = divmod(i, block)
    if b < full:
        return (b * world + rank) * block + j
    return full * world * block + rank * last + j


def cyclic_global_indices(n: int, block: int, world: int, rank: int):
    """cyclic_global_index for every local element, as a numpy int64 array."""
    import numpy as np

    i = np.arange(n, dtype=np.int64)
    full, last = divmod(n, block)
    b, j = np.divmod(i, block)
    return np.where(b < full, (b * world + rank) * block + j, full * world * block + rank * last + j)


class ShardedOps:
    def __init__(self, dist, rank: int, world: int, reduce: Callable, scan: Callable, carry_add: Callable, exclusive_carry: Callable,
                 scan_cyclic: Callable | None = None, segreduce: Callable | None = None, segscan: Callable | None = None,
                 reduce_all: Callable | None = None):
        self.dist, self.rank, self.world = dist, rank, world
        self._reduce, self._scan, self._carry_add, self._exclusive_carry = reduce, scan, carry_add, exclusive_carry
        self._scan_cyclic, self._segreduce, self._segscan, self._reduce_all = scan_cyclic, segreduce, segscan, reduce_all

    # -- replicas: shards cut on segment boundaries need nothing from each other
    def foreach_segreduce(self, x, seg_len: int):
        """In-place segmented reduction of this rank's shard (a multiple of seg_len elements)."""
        self._check_segments(x, seg_len)
        self._segreduce(x, seg_len)
        return x

    def foreach_segscan(self, x, y, seg_len: int, exclusive: bool = False):
        """Segmented scan (implicit heads every seg_len) of this rank's shard."""
        self._check_segments(x, seg_len)
        self._segscan(x, y, seg_len, exclusive)
        return y

    def foreach_group_scan(self, x, y, group_len: int, exclusive: bool = True):
        """Group-scope scan (restart every group_len = --local) of this rank's shard."""
        self._check_segments(x, group_len)
        self._scan(x, y, exclusive, None, group_len)
        return y

    @staticmethod
    def _check_segments(x, seg_len: int):
        n = x.numel() if hasattr(x, "numel") else x.size
        if seg_len <= 0 or n % seg_len != 0:
            raise ValueError(f"shard of {n} elements is not cut on segment boundaries (segment length {seg_len})")

    def reduce(self, x, out, exchange: str = "fused"):
        """out[0] = sum over ALL shards (every rank gets it).  exchange: "fused" (in-kernel, needs a peer group;
        falls back to "nccl" without one) or "nccl" (torch.distributed all-reduce of one element)."""
        if exchange == "fused" and self._reduce_all is not None and self.world > 1:
            self._reduce_all(x, out)
            return out
        self._reduce(x, out)
        if self.world > 1:
            self.dist.all_reduce(out)
        return out

    def prefixsum(self, x, y, totals, carry, strategy: str = "reduce_then_scan", exclusive: bool = False):
        """y = this rank's shard of the GLOBAL scan.  totals: world elements, carry: 1 element (scratch)."""
        if strategy == "cyclic":
            if self._scan_cyclic is None:
                raise ValueError("no block-cyclic scan available in this ShardedOps")
            self._scan_cyclic(x, y, exclusive)
            return y
        if self.world == 1:
            self._scan(x, y, exclusive, None)
            return y
        if strategy == "reduce_then_scan":
            self._reduce(x, carry)
            self.dist.all_gather_into_tensor(totals, carry)
            self._exclusive_carry(totals, self.rank, carry)
            self._scan(x, y, exclusive, carry)
        elif strategy == "scan_then_carry":
            self._scan(x, y, False, None)
            # shard total = last element of the inclusive scan
            carry.copy_(y[-1:])
            self.dist.all_gather_into_tensor(totals, carry)
            self._exclusive_carry(totals, self.rank, carry)
            if exclusive:
                raise NotImplementedError("scan_then_carry is provided for the inclusive prefix sum")
            self._carry_add(y, carry)
        else:
            raise ValueError(strategy)
        return y


def make_peer_group(ctx, dist, rank: int, world: int):
    """One sb_peer per rank, connected to every other rank: the 64-byte CUDA IPC handles of the mailboxes are
    exchanged ONCE through torch.distributed (all_gather_object); the scans themselves make no collective call."""
    from .sbcuda import Peer

    peer = Peer(ctx, rank, world)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, peer.handle())
        for r, h in enumerate(handles):
            if r != rank:
                peer.connect(r, h)
        dist.barrier()  # every mailbox is mapped everywhere before the first kernel writes into one
    return peer


def for_context(ctx, dist, rank: int, world: int, peer=None) -> ShardedOps:
    """ShardedOps over libsb_cuda (sbcuda.Context).  `peer` (make_peer_group) enables strategy "cyclic" for world > 1;
    with world == 1 the block-cyclic scan needs no peer."""
    scan_cyclic = None
    if peer is not None or world == 1:
        scan_cyclic = lambda x, y, exclusive: ctx.scan_cyclic(x, y, peer=peer, exclusive=exclusive)  # noqa: E731
    return ShardedOps(
        dist, rank, world,
        reduce=lambda x, out: ctx.reduce(x, out),
        scan=lambda x, y, exclusive, seed, group_len=0: ctx.scan(x, y, exclusive=exclusive, seed=seed, group_len=group_len),
        carry_add=lambda y, c: ctx.carry_add(y, c),
        exclusive_carry=lambda t, r, c: ctx.exclusive_carry(t, r, c),
        scan_cyclic=scan_cyclic,
        reduce_all=(lambda x, out: ctx.reduce_all(x, out, peer=peer)) if peer is not None else None,
        segreduce=lambda x, L: ctx.segreduce(x, L),
        segscan=lambda x, y, L, exclusive: ctx.segscan(x, y, seg_len=L, exclusive=exclusive),
    )
