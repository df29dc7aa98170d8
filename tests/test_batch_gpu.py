"""Query-batch mode (hhg_query_set_batch + hhg_viterbi_search_batch): several queries of different
lengths in ONE plan / launch must give, request by request, exactly what the single-query calls give."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hits_equal(a, b):
    for f in a.dtype.names:
        if f == "path_off":
            continue
        x, y = a[f], b[f]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        assert np.array_equal(x, y), f


def _setup(nq_lens, n, seed):
    from hhsuite_b200 import synth
    rng = np.random.default_rng(seed)
    queries = [synth.query_profile(L, 40 + k) for k, L in enumerate(nq_lens)]
    tg = [synth.prepared_profile(int(L), rng, queries[k % len(queries)][4] if k % 3 == 0 else None, noise=0.3)
          for k, L in enumerate(rng.integers(10, 420, n))]
    return queries, tg, rng


@pytest.mark.parametrize("strip_rows", [None, 16])
def test_batch_equals_single_queries_prepared_shard(hhg, oracle, strip_rows, monkeypatch):
    if strip_rows:
        monkeypatch.setenv("HHG_STRIP_ROWS", str(strip_rows))
    queries, tg, rng = _setup([97, 400, 33, 211], 300, 5)
    ctx = hhg.Context()
    db = hhg.TargetDB.from_profiles(ctx, tg)
    # request list: every query gets its own random subset of targets (different sizes, one with a single target)
    req_q, ids = [], []
    for q, m in enumerate([120, 300, 1, 77]):
        sel = rng.permutation(300)[:m]
        req_q += [q] * m
        ids += sel.tolist()
    perm = rng.permutation(len(ids))                       # requests of different queries interleaved arbitrarily
    req_q = np.array(req_q, np.int32)[perm]; ids = np.array(ids, np.int32)[perm]
    hhg.capi.query_set_batch(ctx, [(q[0], q[1], q[2]) for q in queries])
    hb, pb_ = hhg.capi.viterbi_search_batch(ctx, db, req_q, ids)
    for q, qq in enumerate(queries):
        ctx.set_query(qq[0], qq[1])
        m = np.nonzero(req_q == q)[0]
        hs, ps = hhg.viterbi_search(ctx, db, ids=ids[m])
        _hits_equal(hb[m], hs)
        for k, r in enumerate(m):
            n = int(hs[k]["nsteps"])
            assert np.array_equal(pb_[hb[r]["path_off"]:hb[r]["path_off"] + n], ps[hs[k]["path_off"]:hs[k]["path_off"] + n])
    # and one direct oracle check per query
    for q, qq in enumerate(queries):
        r = int(np.nonzero(req_q == q)[0][0])
        tp, ttr, _ = tg[ids[r]]
        sc, i2, j2, bt = oracle.viterbi(qq[0], qq[1], tp, ttr)
        assert np.float32(sc).view(np.uint32) == hb[r]["score"].view(np.uint32) and (i2, j2) == (hb[r]["i2"], hb[r]["j2"])
    db.close(); ctx.close()


@pytest.mark.parametrize("columnscore", [1, 0, 2, 3])
def test_batch_over_raw_shard_fuses_the_null_model(hhg, columnscore):
    """Raw shard (emissions before the null model + pav): the batch search factors HMM::IncludeNullModelInHMM in per
    query while it builds its operand stream; reference = hhg_db_apply_null_model + single search per query."""
    from hhsuite_b200 import synth
    queries, tg, rng = _setup([150, 64, 333], 200, 9)
    n = len(tg)
    pav = rng.dirichlet(np.ones(20) * 5, n).astype(np.float32)
    pb = rng.dirichlet(np.ones(20) * 5).astype(np.float32)
    q_pav = np.stack([q[3] for q in queries]).astype(np.float32)
    L = np.array([t[0].shape[0] - 2 for t in tg], np.int32)
    P = np.concatenate([t[0] for t in tg]); T = np.concatenate([t[1] for t in tg])
    p_off = np.concatenate([[0], np.cumsum(L + 2)[:-1]]); tr_off = np.concatenate([[0], np.cumsum(L + 1)[:-1]])
    ctx = hhg.Context()
    db = hhg.TargetDB(ctx, L, P, T, p_off, tr_off, pav=pav)
    req_q = rng.integers(0, 3, 350).astype(np.int32)
    ids = rng.integers(0, n, 350).astype(np.int32)
    hhg.capi.query_set_batch(ctx, [(q[0], q[1], q[2]) for q in queries], q_pav=q_pav)
    hb, pth = hhg.capi.viterbi_search_batch(ctx, db, req_q, ids, columnscore=columnscore, pb=pb)
    for q, qq in enumerate(queries):
        ctx.set_query(qq[0], qq[1])
        db.apply_null_model(q_pav[q], pb, columnscore)
        m = np.nonzero(req_q == q)[0]
        hs, ps = hhg.viterbi_search(ctx, db, ids=ids[m])
        _hits_equal(hb[m], hs)
    db.close(); ctx.close()
