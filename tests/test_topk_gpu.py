"""Top-K selection on the device and the NCCL exchange behind the C-ABI (hhg_plan_topk / hhg_plan_topk_paths /
hhg_comm_*): the merged list must be the single-process list ordered by (score descending, global id ascending)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected(hits, gids, K, field):
    order = np.lexsort((gids, -hits[field].astype(np.float64)))[:K]
    return order


def _make(n, seed, lq=120):
    from hhsuite_b200 import synth
    rng = np.random.default_rng(seed)
    qp, qtr, qss, qpav, qcols = synth.query_profile(lq, 3)
    tg = [synth.prepared_profile(int(L), rng, qcols if k % 7 == 0 else None, noise=0.3)
          for k, L in enumerate(rng.integers(20, 260, n))]
    for k in range(0, n, 10):          # exact duplicates: equal scores, the id decides
        tg[(k + 5) % n] = tg[k]
    return (qp, qtr), tg


@pytest.mark.parametrize("K", [1, 37, 500, 5000])
def test_single_gpu_topk_matches_host_sort(hhg, gpu_ctx, K):
    (qp, qtr), tg = _make(900, 4)
    gpu_ctx.set_query(qp, qtr)
    db = hhg.TargetDB.from_profiles(gpu_ctx, tg)
    plan = hhg.Plan(gpu_ctx, db)
    plan.run()
    hits, paths = plan.fetch()
    for field, flag in (("score", False), ("hit_score", True)):
        gids = np.arange(900, dtype=np.int32) * 3 + 7            # arbitrary global ids
        rec = plan.topk(K, by_hit_score=flag, global_ids=gids)
        exp = _expected(hits, gids, K, field)
        assert len(rec) == min(K, 900)
        assert np.array_equal(rec["target"], gids[exp])
        assert np.array_equal(rec["hit"][field].view(np.uint32), hits[field][exp].view(np.uint32))
        assert np.all(rec["owner"] == 0)
        rec2 = plan.topk(K, by_hit_score=flag, id_base=1000)     # id_base form
        assert np.array_equal(rec2["target"], 1000 + _expected(hits, np.arange(900), K, field))
    # path rows of the merged list
    rec = plan.topk(64, by_hit_score=True)
    rows = plan.topk_paths(rec)
    for r, t in enumerate(rec["target"]):
        h = hits[t]
        assert np.array_equal(rows[r, :h["nsteps"]], paths[h["path_off"]:h["path_off"] + h["nsteps"]])
        assert not rows[r, h["nsteps"]:].any()
    plan.close(); db.close()


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("world", [2])
def test_sharded_search_merges_to_the_single_gpu_list(hhg, world):
    """One host thread per GPU, shards from shard.balanced_shards, NCCL rendezvous through hhg_comm_unique_id."""
    from hhsuite_b200 import shard
    (qp, qtr), tg = _make(700, 9)
    L = np.array([t[0].shape[0] - 2 for t in tg])
    K = 200
    ctx0 = hhg.Context(device=0)
    ctx0.set_query(qp, qtr)
    db0 = hhg.TargetDB.from_profiles(ctx0, tg)
    p0 = hhg.Plan(ctx0, db0)
    p0.run()
    ref = p0.topk(K, by_hit_score=True)
    ref_rows = p0.topk_paths(ref)
    p0.close(); db0.close(); ctx0.close()

    parts = shard.balanced_shards(L, world)
    uid = hhg.Comm.unique_id()
    out, errors = [None] * world, []

    def worker(r):
        try:
            ctx = hhg.Context(device=r)
            comm = hhg.Comm(ctx, r, world, uid)
            ctx.set_query(qp, qtr)
            db = hhg.TargetDB.from_profiles(ctx, [tg[t] for t in parts[r]])
            plan = hhg.Plan(ctx, db)
            plan.run()
            rec = plan.topk(K, comm=comm, by_hit_score=True, global_ids=parts[r])
            rows = plan.topk_paths(rec, comm=comm)
            out[r] = (rec.copy(), rows.copy())
            plan.close(); db.close(); comm.close(); ctx.close()
        except BaseException as e:   # noqa: BLE001
            errors.append((r, repr(e)))

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errors, errors
    for r in range(world):
        rec, rows = out[r]
        assert np.array_equal(rec["target"], ref["target"])
        assert np.array_equal(rec["key"], ref["key"])
        for f in ("score", "hit_score"):
            assert np.array_equal(rec["hit"][f].view(np.uint32), ref["hit"][f].view(np.uint32))
        for f in ("i1", "i2", "j1", "j2", "nsteps", "matched_cols"):
            assert np.array_equal(rec["hit"][f], ref["hit"][f])
        owners = np.array([next(k for k in range(world) if t in set(parts[k].tolist())) for t in rec["target"]])
        assert np.array_equal(rec["owner"], owners)
        assert np.array_equal(rows, ref_rows)


def test_topk_by_score_aass_is_the_reference_list_order(hhg, gpu_ctx):
    """Rank by the reference's sort key: score_aass from hhg_hitlist_pvalues (HitList::CalculatePvalues), ascending;
    the device selection must reproduce the host order (ties by global id)."""
    (qp, qtr), tg = _make(600, 12)
    gpu_ctx.set_query(qp, qtr)
    db = hhg.TargetDB.from_profiles(gpu_ctx, tg)
    plan = hhg.Plan(gpu_ctx, db)
    plan.run()
    hits, _ = plan.fetch(want_paths=False)
    rng = np.random.default_rng(1)
    neff = rng.uniform(1, 12, 600).astype(np.float32)
    Lt = np.array([t[0].shape[0] - 2 for t in tg], np.int32)
    st = hhg.capi.hitlist_pvalues(hits["hit_score"], hits["score_ss"], Lt, neff, 120, 5.5, 600)
    gids = np.arange(600, dtype=np.int32)[::-1].copy()
    rec = hhg.capi.plan_topk_by_key(gpu_ctx, plan.h, None, 150, st["score_aass"], global_ids=gids)
    exp = np.lexsort((gids, st["score_aass"]))[:150]
    assert np.array_equal(rec["target"], gids[exp])
    assert np.array_equal(rec["hit"]["hit_score"].view(np.uint32), hits["hit_score"][exp].view(np.uint32))
    plan.close(); db.close()
