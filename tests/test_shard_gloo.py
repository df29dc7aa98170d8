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
