"""Multi-rank host logic on CPU: world_size-2 gloo.  Each rank scores its shard with the C oracle (the GPU
kernel's stand-in here), the ranks all_gather their top-K records, and the merged list must equal the
single-process top-K."""
import os

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, n, k, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hhsuite_b200 import shard, synth
    from hhsuite_b200.capi import HIT_DTYPE
    from oracle.binding import Oracle
    O = Oracle()
    qp, qtr, qss, qpav, qcols = synth.query_profile(60, 1)
    db = synth.prepared_db(n, seed=9, query_cols=qcols, planted=6, lo=5, hi=80, median=30)
    ids = shard.balanced_shards(db["L"], world)[rank]
    hits = np.zeros(len(ids), HIT_DTYPE)
    for x, t in enumerate(ids):
        L = int(db["L"][t])
        sc, i2, j2, bt = O.viterbi(qp, qtr, db["p"][db["p_off"][t]:db["p_off"][t] + L + 2],
                                   db["tr"][db["tr_off"][t]:db["tr_off"][t] + L + 1])
        ns, i_s, j_s, st, mc = O.backtrace(bt, i2, j2)
        hits[x] = (sc, i2, j2, i_s[ns], j_s[ns], ns, mc, 0, sc, 0.0)
    merged = shard.allgather_topk(shard.local_topk(hits, ids, k), k)
    if rank == 0:
        ret["merged"] = merged
        ret["shard0"] = ids
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_equals_global_topk():
    from hhsuite_b200 import shard, synth
    from oracle.binding import Oracle
    n, k, world = 60, 10, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29517, n, k, ret), nprocs=world, join=True)
    merged = ret["merged"]
    O = Oracle()
    qp, qtr, qss, qpav, qcols = synth.query_profile(60, 1)
    db = synth.prepared_db(n, seed=9, query_cols=qcols, planted=6, lo=5, hi=80, median=30)
    sc = np.array([O.viterbi(qp, qtr, db["p"][db["p_off"][t]:db["p_off"][t] + db["L"][t] + 2],
                             db["tr"][db["tr_off"][t]:db["tr_off"][t] + db["L"][t] + 1], want_bt=False)[0]
                   for t in range(n)], np.float32)
    order = np.lexsort((np.arange(n), -sc))[:k]
    assert np.array_equal(merged["target"], order)
    assert np.array_equal(merged["score"].view(np.uint32), sc[order].view(np.uint32))
    # shards are balanced and disjoint
    shards = shard.balanced_shards(db["L"], world)
    assert sorted(np.concatenate(shards).tolist()) == list(range(n))
    loads = [int(db["L"][s].sum()) for s in shards]
    assert abs(loads[0] - loads[1]) <= int(db["L"].max())


# ---------------------------------------------------------------------------------------------------------------
# sharded two-stage prefilter + path exchange (host logic; the oracle's byte kernels stand in for the GPU's)
# ---------------------------------------------------------------------------------------------------------------
def _pf_inputs():
    from tests.util import golden
    G = golden()
    prof = G["pf_prof"]
    rng = np.random.default_rng(12)
    best = prof[:219].argmax(axis=0).astype(np.uint8)
    seqs = [rng.integers(0, 219, int(L), dtype=np.uint8) for L in rng.integers(20, 120, 90)]
    for k in range(14):                       # planted partial matches of varying quality
        a = int(rng.integers(0, 300)); seg = best[a:a + int(rng.integers(40, 110))].copy()
        noise = rng.random(len(seg)) < 0.1 + 0.04 * k
        seg[noise] = rng.integers(0, 219, int(noise.sum()), dtype=np.uint8)
        seqs.append(seg)
    seqs += [seqs[3].copy() for _ in range(5)]      # exact score ties across shards
    perm = rng.permutation(len(seqs))
    return prof, [seqs[i] for i in perm]


def _pf_worker(rank, world, port, kw, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hhsuite_b200 import shard
    from oracle.binding import Oracle
    O = Oracle()
    prof, seqs = _pf_inputs()
    Lq = prof.shape[1]
    n = len(seqs)
    ids = np.arange(rank, n, world)                 # round-robin shard
    lens = np.array([len(seqs[g]) for g in ids], np.int32)

    def stage1():
        corr = np.array([O.lib.hho_ungapped_corrected(O.ungapped(prof, seqs[g], 50), Lq, len(seqs[g]), 4) for g in ids])
        order = np.lexsort((ids, corr))[::-1]
        stop = np.nonzero(corr[order[kw["min_hits"]:]] <= kw["smax_thresh"])[0]
        ncut = kw["min_hits"] + int(stop[0]) if len(stop) else len(ids)
        return order[:ncut].astype(np.int32), corr[order[:ncut]]

    def stage2(local):
        return np.array([O.sw_byte(prof, seqs[ids[x]], 24, 4, 50) for x in local], np.int32)

    out = shard.sharded_prefilter(stage1, stage2, ids, lens, n, Lq, **kw)
    # path exchange: pretend each kept sequence has an alignment path derived from its id
    rec = np.zeros(len(out), shard.REC_DTYPE)
    rec["target"] = out
    rec["nsteps"] = 5 + (out % 7)
    paths = {int(g): np.full(5 + int(g) % 7, 2 + int(g) % 5, np.uint8) for g in ids}
    P = shard.allgather_paths(rec, ids, paths)
    ret[rank] = (out, P)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_prefilter_equals_single_process():
    from hhsuite_b200 import shard
    from oracle.binding import Oracle
    kw = dict(min_hits=8, smax_thresh=10, evalue_thresh=1000.0, evalue_coarse=100000.0, maxnumdb=25)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pf_worker, args=(2, 29531, kw, ret), nprocs=2, join=True)
    out0, P0 = ret[0]
    out1, P1 = ret[1]
    assert np.array_equal(out0, out1) and np.array_equal(P0, P1)
    # single-process reference rule on oracle scores (same as tests/test_prefilter_gpu.py)
    O = Oracle()
    prof, seqs = _pf_inputs()
    Lq, n = prof.shape[1], len(seqs)
    corr = [O.lib.hho_ungapped_corrected(O.ungapped(prof, s, 50), Lq, len(s), 4) for s in seqs]
    order = sorted(range(n), key=lambda k: (corr[k], k), reverse=True)
    first = []
    for k in order:
        if len(first) >= kw["min_hits"] and corr[k] <= kw["smax_thresh"]:
            break
        first.append(k)
    sw = [O.sw_byte(prof, seqs[k], 24, 4, 50) for k in first]
    ev = [float(n) * Lq * len(seqs[k]) * O.fpow2(float(int(-s / 4))) for k, s in zip(first, sw)]
    sel = sorted([x for x in range(len(first)) if ev[x] < kw["evalue_coarse"]], key=lambda x: (int(ev[x]), first[x]))   # the reference sorts with the E-value truncated to int
    want = []
    for x in sel:
        if len(want) >= kw["min_hits"] and ev[x] > kw["evalue_thresh"]:
            break
        want.append(first[x])
    want = want[:kw["maxnumdb"]]
    assert out0.tolist() == want and len(want) >= kw["min_hits"]
    # every row of the exchanged path matrix is its owner's path, zero padded
    for r, g in enumerate(out0):
        L = 5 + int(g) % 7
        assert np.all(P0[r, :L] == 2 + int(g) % 5) and not P0[r, L:].any()
    # stage-1 merge helper against the plain rule, with ties straddling the cut
    cand = np.array([[5, 30], [9, 30], [2, 30], [7, 12], [1, 9], [3, 9], [8, 11]], np.int64)
    kept = shard.stage1_merge(cand, 4, 10)
    assert kept[:, 0].tolist() == [9, 5, 2, 7, 8]
