"""Row-sharded index across the GPUs of one box, one process per GPU: ShardedAppendable + ComposedQueryable
(ShardApi.scala:21-48, 58-87) with the shards in different processes.

The reference routes every appended row to one of S sub-indices and, per query, fans out to all S, flattens the S*k
results, sorts and takes k.  Here rank r of R owns one `BruteForceIndex` shard on its GPU; queries are replicated (every
rank sees the same batch), each rank answers over its rows, and the per-rank top-k lists are exchanged and merged:

  route "fused"     -- the default on GPUs that can map each other's memory: the shard's results land in a peer-mapped
                       result block and ONE kernel (`ann_exchange_merge_device`, ann/exchange.py) pulls this rank's slice
                       of the batch from every peer over NVLink, merges it and pushes the merged rows to every peer.
                       With `share_seeds` (default) the local query is the two-phase one: every shard first publishes k
                       bounds per query learnt from a short prefix of its rows (`ann_query_seed_device`), and after one
                       barrier scores its rows against the k-th best bound of ALL shards (`ann_query_finish_device`) -- a
                       threshold as tight as one shard would get from `world` times the prefix.  With `two_round`
                       (default) a second exchange follows the last chunk (`ann_query_filter_device` publishes every shard's
                       k best bounds, `ann_query_rescore_device` rescores only what the k-th best bound of ALL shards lets
                       through), so the exact rescoring shrinks with the shard count too; and with deliver="slice" the
                       merged answer stays partitioned: rank r receives only its 1/world slice, nothing is pushed.
  route "allgather" -- `all_gather` of the three result arrays in the [shards][b][k] layout (NCCL; gloo on CPU) followed
                       by the merge kernel (`ann_merge_topk_device`) over the whole batch on every rank.

Both give every rank the complete merged batch, ordered by (Float.compare(distance), id): with globally unique ids that
is bit for bit the single-shard answer (tests/test_distributed_gloo.py on CPU with test doubles for the two device
calls; tests/checks/dist_check.py and bench.py --gpus N on GPUs).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row partition: rank r holds rows [r*n//R, (r+1)*n//R); sizes differ by at most one."""
    return rank * n // world, (rank + 1) * n // world


def route_batch(batch_index: int, world: int) -> int:
    """Streaming appends go to the shards round-robin by batch -- the deterministic stand-in for RandomShardFunction
    (ShardApi.scala:21-25); shard sizes stay within one batch of each other."""
    return batch_index % world


class ShardedBruteForceIndex:
    """One rank's handle on the sharded index.  `local` is this rank's shard (a BruteForceIndex, or any object with
    `query_batch_device(queries, k, out_ids, out_dist, out_count, stream)`); `merge` is the [S,b,k] merge used by the
    all-gather route (default: the CUDA merge kernel)."""

    def __init__(self, local, group=None, route: str = "auto", merge: Optional[Callable] = None, device=None,
                 share_seeds: bool = True, two_round: bool = True, push: bool = True, sliced_seeds: bool = True):
        import torch
        import torch.distributed as dist

        self.local = local
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self._merge = merge
        self.share_seeds = share_seeds   # fused route: two-phase local query around a cross-shard threshold exchange
        self.two_round = two_round       # ... plus a second exchange of the k best bounds after the last chunk (three phases)
        self.push = push                 # bounds are pushed into the peers' receive buffers instead of pulled by every consumer
        self.sliced_seeds = sliced_seeds  # first round: every rank seeds its SLICE of the batch and delivers one bound per query
        self._px = {}          # (b, k) -> PeerExchange
        self._gather = {}      # (b, k) -> gathered buffers
        self._own = {}         # (b, k) -> this rank's result arrays (all-gather route)
        self._seeds = {}       # (b, k) -> (this rank's published seed bounds, everyone's) (all-gather route)
        self._appended_batches = 0
        if route not in ("auto", "fused", "allgather"):
            raise ValueError(f"route must be auto, fused or allgather, not {route!r}")
        self.route = route
        self.route_note = ""
        if route == "auto":
            self.route = "fused" if (self.device.type == "cuda" and self.world > 1) else "allgather"

    # ------------------------------------------------------------------ appends
    def append_shard(self, ids, rows) -> None:
        """Rows the caller has already assigned to this rank (e.g. its `shard_range` of a bulk load)."""
        self.local.append_batch(ids, rows)

    def append_routed(self, ids, rows) -> bool:
        """Collective-free streaming append: every rank is offered every batch and keeps the ones routed to it.
        Returns whether this rank kept the batch."""
        mine = route_batch(self._appended_batches, self.world) == self.rank
        self._appended_batches += 1
        if mine:
            self.local.append_batch(ids, rows)
        return mine

    # ------------------------------------------------------------------ queries
    def _exchange_for(self, b: int, k: int):
        from .exchange import PeerExchange

        key = (b, k)
        if key not in self._px:
            self._px[key] = PeerExchange(b, k, self.device, self.group, dim=int(getattr(self.local, "dim", 0) or 0))
        return self._px[key]

    def gather_queries(self, host_slice, b: int, k: int, stream: int = 0):
        """Collective (fused route).  Every rank passes ITS rows `slice_range(b)` of the query batch (pinned CPU tensor) and
        gets the complete [b, dim] batch on its device: 1/world of the batch crosses this rank's PCIe link, the rest arrives
        from the peers over NVLink.  Safe to call for the next batch as soon as the previous `batch_query_device` was
        enqueued: the batch buffer is overwritten only after the barrier every rank enters behind its own query kernels."""
        return self._exchange_for(b, k).gather_queries(host_slice, stream)

    def batch_query_device(self, queries, k: int, stream: int = 0, deliver: str = "all", exact: bool = False):
        """Collective: every rank passes the same [b, dim] batch (on its own device).

        deliver="all"   -> every rank receives the whole merged batch (ids [b,k], dist [b,k], count [b]);
        deliver="slice" -> every rank receives only ITS rows [q_begin, q_begin + q_count) of the merged batch (fused route;
                           `slice_range(b)` gives the range): the answer stays partitioned, nothing is pushed between ranks
                           after the merge and the step ends without a trailing barrier.
        The call is asynchronous on `stream` and the returned tensors are reused by the next call with the same (b, k).
        CONTRACT for degenerate inputs: a query that some shard's bounded selector could not answer (thousands of exact
        ties inside the margin, a NaN / zero-norm Cosine query, pool overflow) comes back with count = -1 -- the row is
        INVALID, never silently incomplete.  `batch_query` (the synchronous form) re-answers such batches with
        `exact=True`, which switches every shard to its exact fallback; asynchronous callers check `count < 0` themselves
        after synchronising."""
        import torch
        import torch.distributed as dist

        b = int(queries.shape[0])
        if deliver not in ("all", "slice"):
            raise ValueError("deliver must be 'all' or 'slice'")
        if exact and hasattr(self.local, "set_option"):
            self.local.set_option("device_fallback", 1)
        try:
            return self._batch_query_device(queries, k, stream, deliver, exact, b, torch, dist)
        finally:
            if exact and hasattr(self.local, "set_option"):
                self.local.set_option("device_fallback", 0)

    def slice_range(self, b: int) -> Tuple[int, int]:
        """Rows of a b-query batch that `deliver="slice"` returns on this rank."""
        from .exchange import slice_of

        return slice_of(self.rank, self.world, b)

    def batch_query(self, queries, k: int, stream: int = 0):
        """Synchronous, always-exact form: the whole merged batch on every rank as (ids, dist, count) CUDA tensors.  Batches
        in which any rank saw a flagged query are answered again through every shard's exact fallback (collective: all
        ranks take the same decision because they all hold the same merged counts)."""
        import torch

        ids, dist_, cnt = self.batch_query_device(queries, k, stream, deliver="all")
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        if bool((cnt < 0).any()):
            if hasattr(self.local, "raise_pending_error"):
                try:
                    self.local.raise_pending_error()      # clears the sticky device word of the first attempt
                except Exception:
                    pass
            ids, dist_, cnt = self.batch_query_device(queries, k, stream, deliver="all", exact=True)
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
        return ids, dist_, cnt

    def _batch_query_device(self, queries, k, stream, deliver, exact, b, torch, dist):
        if self.route == "fused":
            try:
                px = self._exchange_for(b, k)
            except Exception as e:   # the ranks cannot map each other's memory: keep the collective route
                self.route, self.route_note = "allgather", f"peer mapping unavailable: {type(e).__name__}: {e}"
            else:
                if exact or not (self.share_seeds and hasattr(self.local, "query_seed_device")):
                    self.local.query_batch_device(queries, k, px.local.ids, px.local.dist, px.local.count, stream)
                elif self.two_round and self.push and self.sliced_seeds and hasattr(self.local, "query_seed_slice_push_device"):
                    # SLICED seeding: this rank seeds only ITS slice of the batch, over `world` times the rows, and writes one
                    # bound per query into every rank's bound array; then the k best bounds after the last chunk, by push
                    q0, q1 = self.slice_range(b)
                    self.local.query_seed_slice_push_device(queries, k, q0, q1 - q0, self.world, px.bound_dst, stream)
                    px.seed_barrier()
                    self.local.query_filter_bounds_push_device(queries, k, px.bound_src, self.world, px.kth_push_dst, stream)
                    px.kth_barrier()
                    self.local.query_rescore_device(queries, k, px.kth_recv_src, px.local.ids, px.local.dist, px.local.count, stream)
                elif self.two_round and self.push and hasattr(self.local, "query_filter_push_device"):
                    # three phases around two small exchanges, bounds delivered by PUSH (P2P stores into every peer's receive
                    # buffer; each consumer then reads local memory): seed bounds, then the k best bounds after the last
                    # chunk, so that every shard rescores only its share of the global survivors
                    self.local.query_seed_push_device(queries, k, px.seed_push_dst, stream)
                    px.seed_barrier()
                    self.local.query_filter_push_device(queries, k, px.seed_recv_src, px.kth_push_dst, stream)
                    px.kth_barrier()
                    self.local.query_rescore_device(queries, k, px.kth_recv_src, px.local.ids, px.local.dist, px.local.count, stream)
                elif self.two_round and hasattr(self.local, "query_filter_device"):
                    # the same with PULL delivery (every consumer reads the peers' arrays over NVLink)
                    self.local.query_seed_device(queries, k, px.seed_keys, stream)
                    px.seed_barrier()
                    self.local.query_filter_device(queries, k, px.seed_ptrs, px.kth_keys, stream)
                    px.kth_barrier()
                    self.local.query_rescore_device(queries, k, px.kth_ptrs, px.local.ids, px.local.dist, px.local.count, stream)
                else:
                    # the shards pool what a short prefix of each taught them: one global threshold instead of `world` local ones
                    self.local.query_seed_device(queries, k, px.seed_keys, stream)
                    px.seed_barrier()
                    self.local.query_finish_device(queries, k, px.seed_ptrs, px.local.ids, px.local.dist, px.local.count, stream)
                return px.merge_slice(stream) if deliver == "slice" else px.exchange_merge(stream)
        key = (b, k)
        if key not in self._own:
            dev = self.device
            self._own[key] = (torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
                              torch.empty((b,), dtype=torch.int32, device=dev))
            self._gather[key] = (torch.empty((self.world, b, k), dtype=torch.int64, device=dev),
                                 torch.empty((self.world, b, k), dtype=torch.float32, device=dev),
                                 torch.empty((self.world, b), dtype=torch.int32, device=dev))
        own, gathered = self._own[key], self._gather[key]
        if not exact and self.share_seeds and hasattr(self.local, "query_seed_device"):
            if key not in self._seeds:
                self._seeds[key] = (torch.empty((b, k), dtype=torch.int32, device=self.device),
                                    torch.empty((self.world, b, k), dtype=torch.int32, device=self.device))
            mine, every = self._seeds[key]
            self.local.query_seed_device(queries, k, mine, stream)
            if self.device.type == "cuda":
                dist.all_gather_into_tensor(every, mine, group=self.group)
            else:
                dist.all_gather(list(every.unbind(0)), mine, group=self.group)
            self.local.query_finish_device(queries, k, [every[s].data_ptr() for s in range(self.world)], own[0], own[1], own[2], stream)
        else:
            self.local.query_batch_device(queries, k, own[0], own[1], own[2], stream)
        for g, o in zip(gathered, own):
            if self.device.type == "cuda":
                dist.all_gather_into_tensor(g, o, group=self.group)
            else:   # gloo has no all_gather_into_tensor
                dist.all_gather(list(g.unbind(0)), o, group=self.group)
        merge = self._merge
        if merge is None:
            from .brute_force import merge_topk_device as merge
        out = merge(gathered[0], gathered[1], gathered[2], k, stream)
        if deliver == "slice":
            q0, q1 = self.slice_range(b)
            return out[0][q0:q1], out[1][q0:q1], out[2][q0:q1]
        return out
