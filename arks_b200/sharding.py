"""Multi-GPU host logic: tenant partitioning and the quota-delta fold for quotas shared across GPUs.

Partitioning: every rate-limit key (namespace, user, model) and quota key (namespace, quotaName) lives inside one
namespace (pkg/gateway/util.go:79-99, pkg/gateway/qosconfig/types.go:51-65), so `gpu = fnv1a64(namespace) mod G` gives
every key exactly one owner and the data path needs no collective.

Shared quotas: when one ArksQuota must be visible on several GPUs (a split hot tenant), each replica applies its own
increments at once and the others' at fold epochs: all-reduce(sum) of the per-GPU delta vectors over NCCL (NVLink /
NVSwitch) — `QuotaDeltaExchange.fold()`. On CPU (gloo) the same code runs over host buffers, which is how the logic is
tested without GPUs.
"""
from __future__ import annotations

import numpy as np


def fnv1a64(b: bytes) -> int:
    h = 0xCBF29CE484222325
    for c in b:
        h = ((h ^ c) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def shard_of(namespace: str, world: int) -> int:
    return fnv1a64(namespace.encode()) % world


def partition_objects(tokens, quotas, endpoints, rank: int, world: int, replicate_namespaces=()):
    """Objects of the namespaces owned by `rank` (plus namespaces replicated on every GPU, e.g. a split hot tenant)."""
    rep = set(replicate_namespaces)
    mine = lambda o: (o["metadata"].get("namespace", "default") in rep or
                      shard_of(o["metadata"].get("namespace", "default"), world) == rank)
    return [t for t in tokens if mine(t)], [q for q in quotas if mine(q)], [e for e in endpoints if mine(e)]


class QuotaDeltaExchange:
    """Fold epoch for shared quotas over torch.distributed. `engine` exposes take_quota_delta()/apply_quota_delta()
    (host form) and, on GPUs, export_quota_delta_dev()/fold_quota_delta_dev() (device form, NCCL).

    `shared_local_idx`: local indices of the quotas that are replicated on every GPU, listed in the same canonical
    order on every rank (local quota numbering differs per rank). Only those rows cross the fabric: the message is
    len(shared) x 3 x 8 bytes."""

    def __init__(self, engine, n_quotas: int, shared_local_idx, device=None):
        self.engine, self.n, self.device = engine, n_quotas, device
        self.idx = np.ascontiguousarray(shared_local_idx, np.int64)
        self._comm = False

    def set_shared(self, n_quotas: int, shared_local_idx):
        """after a table swap: the shared quotas' local indices in the NEW generation (the library refuses a fold whose
        indices belong to another generation, arks_comm_set_shared)"""
        self.n = n_quotas
        self.idx = np.ascontiguousarray(shared_local_idx, np.int64)
        if self._comm:
            self.engine.comm_set_shared(self.idx)

    def fold(self):
        import torch
        import torch.distributed as dist
        if self.device is not None and dist.get_backend() == "nccl":
            # the library's own fold: gather of the shared rows, ncclAllReduce and the apply, all on its compute stream
            # (arks_fold_quota_allreduce). torch.distributed only carries the 128-byte communicator id, once.
            if not self._comm:
                box = [self.engine.comm_unique_id() if dist.get_rank() == 0 else None]
                dist.broadcast_object_list(box, src=0)
                self.engine.comm_init(dist.get_rank(), dist.get_world_size(), box[0], self.idx)
                self._comm = True
            self.engine.fold_quota_allreduce(wait=True)
            return
        own = self.engine.take_quota_delta()
        t = torch.from_numpy(own[self.idx].copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        remote = np.zeros_like(own)
        remote[self.idx] = t.numpy() - own[self.idx]
        self.engine.apply_quota_delta(remote)
