"""Multi-GPU sharding of one batch -- one process per GPU, torch.distributed for plumbing.

The reference is single-device; a batch is a disjoint union of graphs with contiguous node-id ranges
(tasks/ppi_task.py:228,233), which gives two ways to shard the hot path:

  * ``split_batch_by_graphs``  -- cut at graph boundaries: independent units, NO communication
    (PPI / QM9 / packed VarMisuse batches).  This is what ``bench.py --gpus N`` scales with.
  * ``NodeRangePartition``     -- one big graph: rank r owns a contiguous node range and every edge whose
    TARGET it owns, so scatter / softmax / layer-norm / GRU stay local and atomic-free; before each layer the
    source rows owned by other ranks ("halo") are fetched with ONE all-to-all-v of node states
    (``exchange``; NCCL over NVLink on GPUs, gloo on CPU for the logic tests).  Weights are replicated.

Index construction is numpy on the host (deterministic, identical on every rank); the exchange moves torch
tensors on whatever device they live on.
"""
from typing import List, NamedTuple, Optional, Sequence

import numpy as np

from .batching import Batch


def balanced_cuts(weights: np.ndarray, parts: int) -> np.ndarray:
    """Cut points [parts+1] over len(weights) items so every part carries ~equal total weight."""
    csum = np.concatenate([[0], np.cumsum(weights, dtype=np.float64)])
    targets = csum[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate([[0], cuts, [len(weights)]]).astype(np.int64)


def split_batch_by_graphs(batch: Batch, parts: int) -> List[Batch]:
    """Graph-boundary sharding: contiguous runs of graphs with balanced message counts, node ids renumbered
    per shard.  Shards are independent (zero halo); concatenating the shard outputs in order gives the
    batch output."""
    off = batch.graph_node_offsets
    tgt_graph_edges = np.zeros(batch.num_graphs, dtype=np.int64)
    for a in batch.adjacency_lists:
        if a.shape[0]:
            g = np.searchsorted(off, a[:, 1], side="right") - 1
            tgt_graph_edges += np.bincount(g, minlength=batch.num_graphs)
    cuts = balanced_cuts(tgt_graph_edges + 1, parts)
    shards = []
    for r in range(parts):
        g0, g1 = int(cuts[r]), int(cuts[r + 1])
        lo, hi = int(off[g0]), int(off[g1])
        adj = []
        for a in batch.adjacency_lists:
            keep = (a[:, 1] >= lo) & (a[:, 1] < hi) if a.shape[0] else np.zeros(0, bool)
            adj.append((a[keep] - lo).astype(np.int32).reshape(-1, 2))
        shards.append(Batch(node_features=batch.node_features[lo:hi],
                            adjacency_lists=adj,
                            type_to_num_incoming_edges=np.ascontiguousarray(batch.type_to_num_incoming_edges[:, lo:hi]),
                            num_graphs=g1 - g0, num_nodes=hi - lo,
                            num_edges=int(sum(a.shape[0] for a in adj)),
                            graph_node_offsets=off[g0:g1 + 1] - lo))
    return shards


def _make_halo_exchange():
    import torch

    class HaloExchange(torch.autograd.Function):
        @staticmethod
        def forward(ctx, h_own, part, group):
            ctx.part, ctx.group = part, group
            return torch.cat([h_own, part._exchange_rows(h_own, group)], dim=0)

        @staticmethod
        def backward(ctx, g_local):
            part = ctx.part
            g_own = g_local[:part.n_own] + part._return_rows(g_local[part.n_own:], ctx.group)
            return g_own, None, None
    return HaloExchange


class _LazyHaloExchange:
    """torch is imported lazily in this module (the index construction is numpy-only)."""
    _cls = None

    @classmethod
    def apply(cls, *args):
        if cls._cls is None:
            cls._cls = _make_halo_exchange()
        return cls._cls.apply(*args)


_HaloExchange = _LazyHaloExchange


class NodeRangePartition:
    """Rank-local view of a node-range partition (targets owned, halo sources fetched).

    Local node numbering: owned nodes first ([0, n_own), global id = lo + i), then the halo nodes sorted by
    global id.  ``local_adjacency_lists`` / ``local_num_incoming`` feed the layer functions unchanged; only
    the first ``n_own`` output rows are meaningful (halo rows have no incoming edges).
    """

    def __init__(self, adjacency_lists: Sequence[np.ndarray], type_to_num_incoming_edges: Optional[np.ndarray],
                 num_nodes: int, rank: int, world_size: int):
        self.rank, self.world_size, self.num_nodes = rank, world_size, num_nodes
        adj = [np.asarray(a).reshape(-1, 2).astype(np.int64) for a in adjacency_lists]
        indeg = np.zeros(num_nodes, dtype=np.int64)
        for a in adj:
            if a.shape[0]:
                indeg += np.bincount(a[:, 1], minlength=num_nodes)
        # degree-balanced split points: equal sum of (in-degree + 1) per rank
        self.cuts = balanced_cuts(indeg + 1, world_size)
        self.lo, self.hi = int(self.cuts[rank]), int(self.cuts[rank + 1])
        self.n_own = self.hi - self.lo

        mine = [a[(a[:, 1] >= self.lo) & (a[:, 1] < self.hi)] for a in adj]
        srcs = np.concatenate([a[:, 0] for a in mine]) if mine else np.zeros(0, np.int64)
        remote = np.unique(srcs[(srcs < self.lo) | (srcs >= self.hi)])
        self.halo_global = remote                                     # sorted global ids of halo nodes
        self.n_halo = int(remote.shape[0])
        self.n_local = self.n_own + self.n_halo

        def to_local(ids):
            out = ids - self.lo
            is_remote = (ids < self.lo) | (ids >= self.hi)
            out[is_remote] = self.n_own + np.searchsorted(remote, ids[is_remote])
            return out

        self.local_adjacency_lists = []
        for a in mine:
            la = np.stack([to_local(a[:, 0].copy()), a[:, 1] - self.lo], axis=1) if a.shape[0] else np.zeros((0, 2))
            self.local_adjacency_lists.append(la.astype(np.int32).reshape(-1, 2))
        self.num_local_edges = int(sum(a.shape[0] for a in self.local_adjacency_lists))
        if type_to_num_incoming_edges is not None:
            c = np.zeros((len(adj), self.n_local), dtype=np.float32)
            c[:, :self.n_own] = np.asarray(type_to_num_incoming_edges)[:, self.lo:self.hi]
            self.local_num_incoming = c
        else:
            self.local_num_incoming = None

        # exchange lists.  recv: my halo nodes grouped by owner (already sorted by global id => grouped);
        # send: what every peer's halo needs from my range -- computed from the same global data on each
        # rank, so no handshake is needed.
        owner = np.searchsorted(self.cuts, remote, side="right") - 1
        self.recv_counts = np.bincount(owner, minlength=world_size).astype(np.int64)
        # What every peer's halo needs from my range: ONE pass over the edges whose source I own (round 1 looped over the
        # peers and ran np.unique over all edges for each: O(world * M)).  Key = (owner of the target, source): unique keys,
        # sorted, are exactly the peers' need lists in peer order, each sorted by node id -- the order the receivers expect.
        self.send_local_idx: List[np.ndarray] = [np.zeros(0, np.int64) for _ in range(world_size)]
        if adj:
            src_all = np.concatenate([a[:, 0] for a in adj])
            tgt_all = np.concatenate([a[:, 1] for a in adj])
            mine_src = (src_all >= self.lo) & (src_all < self.hi)
            peer = np.searchsorted(self.cuts, tgt_all[mine_src], side="right") - 1
            srcs = src_all[mine_src]
            away = peer != rank
            keys = np.unique(peer[away].astype(np.int64) * np.int64(max(num_nodes, 1)) + srcs[away])
            peers_of_keys = keys // np.int64(max(num_nodes, 1))
            bounds = np.searchsorted(peers_of_keys, np.arange(world_size + 1))
            for p in range(world_size):
                if p != rank:
                    self.send_local_idx[p] = keys[bounds[p]:bounds[p + 1]] % np.int64(max(num_nodes, 1)) - self.lo
        self.send_counts = np.array([x.shape[0] for x in self.send_local_idx], dtype=np.int64)
        self._idx_cache = {}
        self._any_halo = None

    def _send_index(self, device):
        import torch
        key = str(device)
        idx = self._idx_cache.get(key)
        if idx is None:   # index list is fixed per batch: build it on the device once
            idx = torch.as_tensor(np.concatenate(self.send_local_idx) if self.send_counts.sum() else np.zeros(0, np.int64),
                                  device=device)
            self._idx_cache[key] = idx
        return idx

    def any_halo(self, group=None) -> bool:
        """True when ANY rank of the partition has halo rows.  Decided once per partition with one tiny all-reduce (every
        rank must call it, like the exchange itself); block-diagonal batches cut at graph boundaries have none, and then
        ``exchange`` skips the collective altogether (an empty all-to-all-v still costs ~65 us on 8 GPUs)."""
        if self._any_halo is None:
            import torch
            import torch.distributed as dist
            if self.world_size > 1 and dist.is_available() and dist.is_initialized():
                dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
                t = torch.tensor([self.n_halo + int(self.send_counts.sum())], dtype=torch.int64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
                self._any_halo = bool(int(t.item()) > 0)
            else:
                self._any_halo = self.n_halo > 0
        return self._any_halo

    def _exchange_rows(self, h_own, group=None):
        """[n_own, D] owned rows -> [n_halo, D] halo rows (one all-to-all-v)."""
        import torch.distributed as dist
        D = h_own.shape[1]
        if not self.any_halo(group):
            return h_own.new_zeros((0, D))
        idx = self._send_index(h_own.device)
        send = h_own.index_select(0, idx) if idx.numel() else h_own.new_zeros((0, D))
        recv = h_own.new_empty((self.n_halo, D))
        if self.world_size > 1:
            dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=[int(x) for x in self.recv_counts],
                                   input_split_sizes=[int(x) for x in self.send_counts], group=group)
        return recv

    def _return_rows(self, g_halo, group=None):
        """The transpose of ``_exchange_rows``: gradients of the halo rows travel back to their owners (the same
        all-to-all-v with the split sizes swapped) and are ADDED to the owners' rows (a node can be in several peers' halos)."""
        import torch
        import torch.distributed as dist
        D = g_halo.shape[1]
        if not self.any_halo(group):
            return torch.zeros((self.n_own, D), dtype=g_halo.dtype, device=g_halo.device)
        back = g_halo.new_empty((int(self.send_counts.sum()), D))
        if self.world_size > 1:
            dist.all_to_all_single(back, g_halo.contiguous(), output_split_sizes=[int(x) for x in self.send_counts],
                                   input_split_sizes=[int(x) for x in self.recv_counts], group=group)
        g_own = torch.zeros((self.n_own, D), dtype=g_halo.dtype, device=g_halo.device)
        idx = self._send_index(g_halo.device)
        if idx.numel():
            g_own.index_add_(0, idx, back)
        return g_own

    def exchange(self, h_own, group=None):
        """[n_own, D] owned states -> [n_local, D] = owned rows followed by the halo rows, via one
        all-to-all-v (torch.distributed.all_to_all_single with uneven splits).  Differentiable: under autograd the
        backward sends the halo rows' gradients back to their owners with the transposed exchange, so a layer
        stack over a node-range partition can be trained (weights: scaffold.all_reduce_gradients_)."""
        import torch
        assert h_own.shape[0] == self.n_own
        if torch.is_grad_enabled() and h_own.requires_grad:
            return _HaloExchange.apply(h_own, self, group)
        return torch.cat([h_own, self._exchange_rows(h_own, group)], dim=0)

    def halo_bytes(self, D: int) -> int:
        """Bytes this rank receives per layer (fp32 rows of width D)."""
        return self.n_halo * D * 4
