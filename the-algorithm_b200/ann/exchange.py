"""Shard merge fused with its exchange: the multi-GPU half of ComposedQueryable (ShardApi.scala:72-86).

The reference fans a query out to S shards, collects S lists of k and sorts the S*k entries on one thread.  Row-sharded
over the GPUs of one box the same step used to be `all_gather` (every rank receives world*b*k entries) + a merge kernel
over all b queries on every rank.  Here it is ONE kernel over NVLink peer memory (`ann_exchange_merge_device`): every
rank's local top-k lives in a result block its peers have mapped, rank r pulls only the rows of ITS slice of the batch
from the `world` blocks, merges them, and pushes the merged rows into every rank's final block.  Each rank moves and
merges 1/world of the batch; two stream-ordered cross-rank barriers bracket the kernel.

    px = PeerExchange(b, k, device)                   # collective: allocates + maps the symmetric blocks
    ix.query_batch_device(q, k, *px.local.tensors, stream)   # the shard's results land in the mapped block directly
    ids, dist, cnt = px.exchange_merge(stream)        # every rank ends up with the whole merged batch

or, sharing the shards' threshold seeds first (`ann_query_seed_device` / `ann_query_finish_device`):

    ix.query_seed_device(q, k, px.seed_keys, stream)  # this shard's k bounds per query, into its mapped key array
    px.seed_barrier()
    ix.query_finish_device(q, k, px.seed_ptrs, *px.local.tensors, stream)   # scored against the global threshold
    ids, dist, cnt = px.exchange_merge(stream)

Peer mapping uses torch's symmetric memory (cuMem allocations exchanged between the ranks of a process group), which is
plumbing in the same sense as the NCCL communicator; the kernel and the layout are this repo's.  `PeerExchange` raises
if the ranks cannot map each other (no P2P); callers then keep the all-gather + `merge_topk_device` route.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

from .. import _capi


def result_block_bytes(b: int, k: int) -> int:
    """Size of one result block: [ids b*k int64][dist b*k float32][count b int32] (include/b200ann.h)."""
    return int(_capi.lib().ann_result_block_bytes(b, k))


class ResultBlock:
    """Typed views of one result block inside a uint8 CUDA tensor."""

    def __init__(self, buf, b: int, k: int, offset: int = 0):
        import torch

        nb = b * k * 12 + b * 4
        assert buf.dtype == torch.uint8 and buf.numel() >= offset + nb
        self.b, self.k, self.offset = b, k, offset
        raw = buf[offset:offset + nb]
        self.ids = raw[: b * k * 8].view(torch.int64).view(b, k)
        self.dist = raw[b * k * 8: b * k * 12].view(torch.float32).view(b, k)
        self.count = raw[b * k * 12:].view(torch.int32)
        self.ptr = buf.data_ptr() + offset

    @property
    def tensors(self) -> Tuple:
        return self.ids, self.dist, self.count


def exchange_merge_blocks(local_ptrs: Sequence[int], final_ptrs: Sequence[int], b: int, k: int, q_begin: int, q_count: int,
                          device: int, stream: int = 0) -> None:
    """`ann_exchange_merge_device` on raw block pointers (all mapped into this process)."""
    world = len(local_ptrs)
    assert world == len(final_ptrs) and world >= 1
    arr = ctypes.c_void_p * world
    _capi.check(_capi.lib().ann_exchange_merge_device(device, arr(*local_ptrs), arr(*final_ptrs), world, b, k, q_begin, q_count,
                                                      ctypes.c_void_p(stream)))


def exchange_merge_slice(local_ptrs: Sequence[int], b: int, k: int, q_begin: int, q_count: int, out_ids_t, out_dist_t, out_count_t,
                         device: int, stream: int = 0) -> None:
    """`ann_exchange_merge_slice_device`: pull + merge this rank's slice into plain [q_count, k] CUDA tensors (no push)."""
    world = len(local_ptrs)
    arr = ctypes.c_void_p * world
    _capi.check(_capi.lib().ann_exchange_merge_slice_device(
        device, arr(*local_ptrs), world, b, k, q_begin, q_count, ctypes.c_void_p(out_ids_t.data_ptr()),
        ctypes.c_void_p(out_dist_t.data_ptr()), ctypes.c_void_p(out_count_t.data_ptr()), ctypes.c_void_p(stream)))


def slice_of(rank: int, world: int, b: int) -> Tuple[int, int]:
    """The queries rank `rank` merges: [rank*b/world, (rank+1)*b/world) -- together the slices tile the batch."""
    return rank * b // world, (rank + 1) * b // world


class PeerExchange:
    """Symmetric result blocks of one (b, k) shape across the ranks of a process group + the fused exchange/merge."""

    def __init__(self, b: int, k: int, device, group=None, dim: int = 0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.b, self.k = b, k
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        nb = result_block_bytes(b, k)
        self._stride = (nb + 255) // 256 * 256
        seed_bytes = (b * k * 4 + 255) // 256 * 256          # this rank's published seed bounds, [b][k] uint32
        # layout: [local block][final block][seed bounds][k-best bounds of the second round]
        #         [seed receive buffer: world blocks][k-best receive buffer: world blocks]   (push delivery)
        #         [query batch b x dim fp32]   (dim > 0: partitioned H2D + push all-gather, `gather_queries`)
        q_bytes = (b * dim * 4 + 255) // 256 * 256
        self._buf = symm_mem.empty(2 * self._stride + 2 * seed_bytes + 2 * self.world * seed_bytes + q_bytes, dtype=torch.uint8,
                                   device=self.device)
        self._hdl = symm_mem.rendezvous(self._buf, self.group)
        ptrs: List[int] = [int(p) for p in self._hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self._buf.data_ptr()
        self._local_ptrs = ptrs
        self._final_ptrs = [p + self._stride for p in ptrs]
        self.seed_ptrs = [p + 2 * self._stride for p in ptrs]
        self.local = ResultBlock(self._buf, b, k, 0)
        self.final = ResultBlock(self._buf, b, k, self._stride)
        self.seed_keys = self._buf[2 * self._stride: 2 * self._stride + b * k * 4].view(torch.int32).view(b, k)
        self.kth_ptrs = [p + 2 * self._stride + seed_bytes for p in ptrs]
        o2 = 2 * self._stride + seed_bytes
        self.kth_keys = self._buf[o2: o2 + b * k * 4].view(torch.int32).view(b, k)
        # push delivery: rank r writes its bounds into block r of EVERY rank's receive buffer; consumers read their own buffer
        o3 = 2 * self._stride + 2 * seed_bytes
        o4 = o3 + self.world * seed_bytes
        me = ptrs[self.rank]
        self.seed_push_dst = [p + o3 + self.rank * seed_bytes for p in ptrs]
        self.kth_push_dst = [p + o4 + self.rank * seed_bytes for p in ptrs]
        # sliced seeding: ONE bound per query, written by the query's slice owner into the head of every rank's seed receive buffer
        self.bound_dst = [p + o3 for p in ptrs]
        self.bound_src = me + o3
        self.seed_recv_src = [me + o3 + s * seed_bytes for s in range(self.world)]
        self.kth_recv_src = [me + o4 + s * seed_bytes for s in range(self.world)]
        self.dim = dim
        o5 = o4 + self.world * seed_bytes
        self._q_ptrs = [p + o5 for p in ptrs]
        self.queries = self._buf[o5: o5 + b * dim * 4].view(torch.float32).view(b, dim) if dim else None
        self.q_begin, q_end = slice_of(self.rank, self.world, b)
        self.q_count = q_end - self.q_begin
        qn = max(self.q_count, 1)
        self.slice_ids = torch.empty((qn, k), dtype=torch.int64, device=self.device)[: self.q_count]
        self.slice_dist = torch.empty((qn, k), dtype=torch.float32, device=self.device)[: self.q_count]
        self.slice_count = torch.empty((qn,), dtype=torch.int32, device=self.device)[: self.q_count]

    def gather_queries(self, host_slice, stream: int = 0):
        """Collective.  `host_slice`: this rank's rows [q_begin, q_begin + q_count) of the query batch (pinned CPU tensor).
        Copies them into the peer-mapped batch buffer, pushes them into the same rows of every peer's buffer
        (`ann_peer_push_device`) and waits at one barrier: the batch crosses PCIe once instead of `world` times.  Returns the
        complete [b, dim] device batch.  Needs dim * 4 * q_begin to be a multiple of 16 (any dim that is a multiple of 4)."""
        assert self.queries is not None, "PeerExchange was created without dim"
        if self.q_count:
            mine = self.queries[self.q_begin: self.q_begin + self.q_count]
            mine.copy_(host_slice, non_blocking=True)
            off = self.q_begin * self.dim * 4
            dst = [p + off for s, p in enumerate(self._q_ptrs) if s != self.rank]
            if dst:
                arr = (ctypes.c_void_p * len(dst))(*dst)
                _capi.check(_capi.lib().ann_peer_push_device(self.device.index or 0, ctypes.c_void_p(self._q_ptrs[self.rank] + off), arr,
                                                             len(dst), self.q_count * self.dim * 4, ctypes.c_void_p(stream)))
        self._hdl.barrier(channel=4)
        return self.queries

    def kth_barrier(self) -> None:
        """Collective, between `query_filter_device` and `query_rescore_device`: every rank's k-best bounds are complete."""
        self._hdl.barrier(channel=3)

    def merge_slice(self, stream: int = 0):
        """Collective.  One barrier (every rank's local block is complete), then this rank pulls and merges ITS slice of
        the batch into its own (slice_ids, slice_dist, slice_count) -- rows [q_begin, q_begin + q_count) of the answer.
        Nothing is pushed and no barrier follows: a peer overwrites the block this kernel reads only after the barriers of
        the next batch, which this rank enters after the kernel (stream order)."""
        self._hdl.barrier(channel=0)
        if self.q_count:
            exchange_merge_slice(self._local_ptrs, self.b, self.k, self.q_begin, self.q_count, self.slice_ids, self.slice_dist,
                                 self.slice_count, self.device.index or 0, stream)
        return self.slice_ids, self.slice_dist, self.slice_count

    def seed_barrier(self) -> None:
        """Collective, between `query_seed_device` and `query_finish_device`: every rank's published bounds are complete.
        No second barrier is needed before the next batch overwrites `seed_keys`: a rank gets there only through the
        barriers of `exchange_merge`, which every peer enters after its own finish call has read the keys."""
        self._hdl.barrier(channel=2)

    def exchange_merge(self, stream: int = 0):
        """Collective.  `local` must have been written on the CURRENT torch stream (== `stream`): the barriers are enqueued
        on torch's current stream, the kernel on `stream`.  Returns the final block's (ids, dist, count) views."""
        self._hdl.barrier(channel=0)     # every rank's local block is complete
        exchange_merge_blocks(self._local_ptrs, self._final_ptrs, self.b, self.k, self.q_begin, self.q_count,
                              self.device.index or 0, stream)
        self._hdl.barrier(channel=1)     # every rank's slice has landed in every final block; locals may be reused
        return self.final.tensors
