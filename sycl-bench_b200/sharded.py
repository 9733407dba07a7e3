"""Multi-GPU data path: one process per GPU, contiguous shards, ONE tiny exchange per operation.

    reduction : local sb_reduce  -> all-reduce(sum) of one element per rank            (traffic N*s/P per GPU)
    prefix sum: "reduce_then_scan": local sb_reduce -> all-gather of the P shard totals
                                    -> sb_exclusive_carry -> sb_scan seeded with the carry  (3*N*s/P per GPU)
                "scan_then_carry" : local sb_scan -> all-gather of the shard totals
                                    -> sb_exclusive_carry -> sb_carry_add                    (4*N*s/P per GPU,
                                    the literal BASELINE.json north_star recipe)
    segmented reduction / segmented scan / group-scope scan: shards cut on segment boundaries are
    independent — no collective at all.

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


class ShardedOps:
    def __init__(self, dist, rank: int, world: int, reduce: Callable, scan: Callable, carry_add: Callable, exclusive_carry: Callable):
        self.dist, self.rank, self.world = dist, rank, world
        self._reduce, self._scan, self._carry_add, self._exclusive_carry = reduce, scan, carry_add, exclusive_carry

    def reduce(self, x, out):
        """out[0] = sum over ALL shards (every rank gets it)."""
        self._reduce(x, out)
        if self.world > 1:
            self.dist.all_reduce(out)
        return out

    def prefixsum(self, x, y, totals, carry, strategy: str = "reduce_then_scan", exclusive: bool = False):
        """y = this rank's shard of the GLOBAL scan.  totals: world elements, carry: 1 element (scratch)."""
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


def for_context(ctx, dist, rank: int, world: int) -> ShardedOps:
    """ShardedOps over libsb_cuda (sbcuda.Context)."""
    return ShardedOps(
        dist, rank, world,
        reduce=lambda x, out: ctx.reduce(x, out),
        scan=lambda x, y, exclusive, seed: ctx.scan(x, y, exclusive=exclusive, seed=seed),
        carry_add=lambda y, c: ctx.carry_add(y, c),
        exclusive_carry=lambda t, r, c: ctx.exclusive_carry(t, r, c),
    )
