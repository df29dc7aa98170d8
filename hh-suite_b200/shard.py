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


# ------------------------------------------------------------------------------------------------
# Path exchange for the merged top-K ("a second all-gather of padded paths for the merged top-K only")
# ------------------------------------------------------------------------------------------------
def allgather_paths(merged: np.ndarray, owned_ids: np.ndarray, owned_paths: dict, device=None) -> np.ndarray:
    """merged: the global top-k records (every rank holds the same array).  owned_ids: the global target ids this
    rank aligned; owned_paths[target] = uint8 state string of that target's alignment (nsteps entries).
    Returns a [k, max_nsteps] uint8 array, row r = path of merged[r], zero padded -- one all_reduce of k*max_nsteps
    bytes (each row has exactly one owner, so a sum is a gather)."""
    import torch
    import torch.distributed as dist
    k = len(merged)
    width = int(merged["nsteps"].max()) if k else 0
    mine = set(int(t) for t in owned_ids)
    buf = np.zeros((k, max(width, 1)), np.uint8)
    for r in range(k):
        t = int(merged["target"][r])
        if t in mine:
            p = np.asarray(owned_paths[t], np.uint8)
            assert len(p) == int(merged["nsteps"][r]), "path length disagrees with the gathered record"
            buf[r, :len(p)] = p
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------------
# Sharded two-stage prefilter (SURVEY 8e last sentence): every rank scores its own cs219 shard; the reference's
# selection rules are GLOBAL (count thresholds, E-values with the global database size), so each rank contributes
# its local candidates and every rank applies the rule to the union.  Exactness: the global "first min_hits by
# (score, id)" are among the ranks' local first min_hits, and "score > thresh" is a per-entry predicate.
# ------------------------------------------------------------------------------------------------
def _allgather_var(arr: np.ndarray, device=None) -> np.ndarray:
    """all_gather of int64 rows with a different count per rank (two collectives: counts, padded payload)."""
    import torch
    import torch.distributed as dist
    arr = np.ascontiguousarray(arr, np.int64)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return arr
    world = dist.get_world_size()
    cols = arr.shape[1]
    cnt = torch.tensor([arr.shape[0]], dtype=torch.int64, device=device)
    cnts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, cnt)
    m = int(cnts.max().item())
    pad = torch.zeros((m, cols), dtype=torch.int64, device=device)
    if arr.shape[0]:
        pad[:arr.shape[0]] = torch.from_numpy(arr).to(pad.device)
    out = torch.empty((world * m, cols), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, pad)
    out = out.cpu().numpy().reshape(world, m, cols)
    return np.concatenate([out[r, :int(cnts[r].item())] for r in range(world)], axis=0)


def stage1_merge(cand: np.ndarray, min_hits: int, smax_thresh: int) -> np.ndarray:
    """cand: int64 [m, 2] rows (global id, corrected score) from all ranks.  The reference's rule on the union:
    sort descending by (score, id), keep while count < min_hits or score > smax_thresh
    (src/hhprefilter.cpp:489-506).  Returns the kept rows in that order."""
    if len(cand) == 0:
        return cand.reshape(0, 2)
    order = np.lexsort((cand[:, 0], cand[:, 1]))[::-1]
    c = cand[order]
    stop = np.nonzero(c[min_hits:, 1] <= smax_thresh)[0]
    ncut = min_hits + int(stop[0]) if len(stop) else len(c)
    return c[:ncut]


def stage2_merge(cand: np.ndarray, ev: np.ndarray, min_hits: int, evalue_thresh: float, evalue_coarse: float,
                 maxnumdb: int) -> np.ndarray:
    """cand: global ids [m]; ev: their E-values (computed with the GLOBAL database size).  Keep ev < coarse
    (:530), sort ascending by ((int)ev, id) (:545; the reference's comparator truncates the E-value to int), keep while count < min_hits or ev <= evalue_thresh (:547-558),
    cap at maxnumdb (:590).  Returns global ids in the reference's output order."""
    keep = np.nonzero(ev < evalue_coarse)[0]
    order = keep[np.lexsort((cand[keep], ev[keep].astype(np.int64)))]   # (int)evalue: comparePair takes pair<int,int>
    # keep while count < min_hits or ev <= evalue_thresh: the first position >= min_hits with ev > thresh ends the list
    tail = np.nonzero(ev[order[min_hits:]] > evalue_thresh)[0]
    ncut = min_hits + int(tail[0]) if len(tail) else len(order)
    return np.asarray(cand, np.int64)[order[:min(ncut, maxnumdb)]]


def sharded_prefilter(score_stage1, score_stage2, local_ids: np.ndarray, local_len: np.ndarray, n_global: int,
                      Lq: int, bit_factor=4, smax_thresh=10, min_hits=100, evalue_thresh=1000.0,
                      evalue_coarse=100000.0, maxnumdb=20000, device=None):
    """Prefilter::prefilter_db over a database sharded by sequence.

    score_stage1() -> (local indices, corrected scores) of this rank's stage-1 candidates, i.e. the LOCAL
        application of the keep rule (capi.CsDB.run + select); local order must be monotone in the global id
        (contiguous, round-robin or balanced_shards' id-sorted shards) so that ties break the same way;
    score_stage2(local indices) -> gapped byte-SW scores (capi.CsDB.sw);
    local_ids[x] / local_len[x]: global id and length of local sequence x; n_global: total sequences (the E-value's
    database size).  Returns the global ids kept, identical on every rank and identical to the single-process result."""
    from . import capi
    import ctypes as C
    local_ids = np.asarray(local_ids, np.int64)
    local_len = np.asarray(local_len, np.int64)
    li, sc = score_stage1()
    li = np.asarray(li, np.int64)
    cand = np.stack([local_ids[li], np.asarray(sc, np.int64)], axis=1) if len(li) else np.zeros((0, 2), np.int64)
    first = stage1_merge(_allgather_var(cand, device), min_hits, smax_thresh)
    # stage 2 on the members this rank owns: a global id is mine iff it occurs in my (id-sorted) shard; every member
    # of `first` that is mine was one of my own stage-1 candidates (first is a subset of the union of the candidates)
    g = first[:, 0]
    x = np.searchsorted(local_ids, g)
    x[x >= len(local_ids)] = 0
    mine = local_ids[x] == g if len(local_ids) else np.zeros(len(g), bool)
    own_local = x[mine].astype(np.int32)
    rows = np.zeros((0, 3), np.int64)
    if len(own_local):
        sw = np.ascontiguousarray(score_stage2(own_local), np.int32)
        rows = np.stack([g[mine], sw.astype(np.int64), local_len[own_local]], axis=1)     # (id, score, length)
    allrows = _allgather_var(rows, device)
    # E-values with the GLOBAL database size
    ev = np.zeros(len(allrows), np.float64)
    if len(allrows):
        sw32 = np.ascontiguousarray(allrows[:, 1], np.int32)
        lens = np.ascontiguousarray(allrows[:, 2], np.int32)
        capi._ck(capi.load().hhg_prefilter_evalues(len(allrows), capi._p(sw32, capi.c_i32p), capi._p(lens, capi.c_i32p),
                                                   n_global, Lq, bit_factor, ev.ctypes.data_as(C.POINTER(C.c_double))))
    return stage2_merge(allrows[:, 0], ev, min_hits, evalue_thresh, evalue_coarse, maxnumdb)
