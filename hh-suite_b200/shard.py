"""Target sharding across GPUs and the top-K hit exchange.

Targets are independent units, so the database shards by target with NO data-path collective; the only
exchange per query is one all_gather of K fixed-size hit records per rank (torch.distributed: NCCL on
GPUs, gloo in the CPU tests).  The reference has no equivalent (its MPI layer distributes queries,
lib/ffindex/src/mpq/mpq.c:42-105); semantics preserved: every target is aligned exactly once and the
merged list is ordered like a single-process search (score descending, target id ascending on ties)."""
from __future__ import annotations

import numpy as np

REC_DTYPE = np.dtype([("target", np.int32), ("score", np.float32), ("i2", np.int32), ("j2", np.int32),
                      ("i1", np.int32), ("j1", np.int32), ("nsteps", np.int32), ("matched_cols", np.int32)])


def balanced_shards(L: np.ndarray, world: int) -> list[np.ndarray]:
    """Longest-first greedy partition of target ids into `world` shards with balanced sum of lengths
    (the DP cost of a target is Lq*Lt).  Deterministic; each shard is returned sorted by id."""
    L = np.asarray(L)
    order = np.argsort(-L, kind="stable")
    load = np.zeros(world, np.int64)
    out: list[list[int]] = [[] for _ in range(world)]
    for t in order:
        r = int(np.argmin(load))
        out[r].append(int(t))
        load[r] += int(L[t])
    return [np.array(sorted(s), np.int32) for s in out]


def local_topk(hits: np.ndarray, ids: np.ndarray, k: int) -> np.ndarray:
    """Top-k records of one shard, padded to exactly k rows (score=-inf, target=-1)."""
    rec = np.zeros(k, REC_DTYPE)
    rec["target"] = -1
    rec["score"] = -np.inf
    n = len(hits)
    order = np.lexsort((ids, -hits["score"]))[:k]
    m = len(order)
    rec["target"][:m] = ids[order]
    for f in ("score", "i2", "j2", "i1", "j1", "nsteps", "matched_cols"):
        rec[f][:m] = hits[f][order]
    return rec


def merge_topk(recs: np.ndarray, k: int) -> np.ndarray:
    """Merge gathered records (any shape [..., k]) into the global top-k."""
    flat = recs.reshape(-1)
    flat = flat[flat["target"] >= 0]
    order = np.lexsort((flat["target"], -flat["score"]))[:k]
    return flat[order]


def allgather_topk(rec: np.ndarray, k: int, device=None) -> np.ndarray:
    """One all_gather of this rank's k records; returns the merged global top-k on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(rec.view(np.int32).reshape(k, 8).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * k, 8), dtype=torch.int32, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return merge_topk(out.cpu().numpy().view(REC_DTYPE).reshape(-1), k)
