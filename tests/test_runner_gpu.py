"""ViterbiRunner mirror (alternative alignments loop) against an emulation of
ViterbiRunner::alignment built from the oracle's primitives."""
import numpy as np
import pytest

from tests.util import bits, rasterize_exclusion

pytestmark = pytest.mark.gpu


def test_alternative_alignments_loop(hhg, gpu_ctx, oracle):
    from hhsuite_b200 import synth
    rng = np.random.default_rng(77)
    qp, qtr, qss, qpav, qcols = synth.query_profile(120, 51)
    targets = []
    for k in range(12):
        if k % 3 == 0:      # two noisy copies of the query in one target -> a genuine second alignment
            a = synth.prepared_profile(120, rng, qcols, noise=0.2)
            b = synth.prepared_profile(120, rng, qcols, noise=0.3)
            p = np.concatenate([a[0][:-1], b[0][1:]])
            tr = np.concatenate([a[1][:-1], b[1]])
            targets.append((np.ascontiguousarray(p), np.ascontiguousarray(tr), None))
        else:
            targets.append(synth.prepared_profile(int(rng.integers(30, 200)), rng, qcols if k % 3 == 1 else None, 0.3))
    gpu_ctx.set_query(qp, qtr)
    db = hhg.TargetDB.from_profiles(gpu_ctx, targets)
    runner = hhg.runner.ViterbiRunner(gpu_ctx, db, altali=4, smin=20.0)
    hits = runner.alignment()
    got = {(h.target, h.irep): h for h in hits}

    # Hit.score decides re-queueing; take it from the GPU hit of the same (target, pass) after checking
    # its raw score/path against the oracle (Hit.score itself is pinned by the golden tests).
    want = {}
    Lq = 120
    todo = list(range(len(targets)))
    masks = {t: np.zeros((Lq + 1, targets[t][0].shape[0] - 1), np.uint8) for t in todo}
    for rep in range(4):
        nxt = []
        for t in todo:
            tp, ttr, _ = targets[t]
            sc, i2, j2, bt = oracle.viterbi(qp, qtr, tp, ttr, celloff=masks[t] if rep > 0 else None)
            n, i_s, j_s, st, mc = oracle.backtrace(bt, i2, j2)
            h = got[(t, rep + 1)]
            assert bits(np.float32(h.vit_score)) == bits(sc), (t, rep)
            assert (h.i2, h.j2, h.nsteps, h.matched_cols) == (i2, j2, n, mc), (t, rep)
            assert np.array_equal(h.i[1:], i_s[1:]) and np.array_equal(h.j[1:], j_s[1:])
            assert np.array_equal(h.states[1:], st[1:])
            assert h.lastrep == (1 if h.score <= 20.0 else 0)
            want[(t, rep + 1)] = True
            if h.score > 20.0:
                nxt.append(t)
                masks[t] |= rasterize_exclusion(Lq, tp.shape[0] - 2, i_s, j_s, n)
        todo = nxt
        if not todo:
            break
    assert set(want) == set(got)
    # the doubled targets must produce a second alignment above smin
    assert all(got[(t, 2)].score > 20.0 for t in range(0, 12, 3))
    assert max(h.irep for h in hits) >= 3
    db.close()


def test_early_stopping_and_ss_consensus_mirror(hhg, gpu_ctx):
    """The Python mirror of the two order-dependent runner behaviours (the exact check against the reference's own
    runner is tests/test_dropin_gpu.py): chunks of 2000 with hhg_early_stop_sum, ss term by batch consensus."""
    from hhsuite_b200 import synth
    from tests.util import golden
    rng = np.random.default_rng(3)
    qp, qtr, qss, qpav, qcols = synth.query_profile(80, 5)
    n = 2100
    tg = [synth.prepared_profile(int(L), rng, None, 0.3) for L in rng.integers(20, 60, n)]
    gpu_ctx.set_query(qp, qtr, qss, golden()["S33"], use_ss=True)
    db = hhg.TargetDB.from_profiles(gpu_ctx, tg)
    neff = rng.uniform(2, 9, n).astype(np.float32)
    t_pred = rng.random(n) < 0.7
    r = hhg.runner.ViterbiRunner(gpu_ctx, db, altali=2, early_stopping=dict(dbsize=n, q_neff=6.0, t_neff=neff),
                                 ss=dict(q_pred=True, t_pred=t_pred, seqlen=db.Lh))
    hits = r.alignment()
    assert r.early_stopped_at == 2000                       # random targets: the first chunk ends the round
    assert sorted({h.target for h in hits}) == list(range(2000))
    # every hit equals a direct search of its target with the ss choice of its batch
    order = np.argsort(-db.Lh[:2000], kind="stable")
    use = np.zeros(2000, bool)
    for b in range(0, 2000, 8):
        use[order[b:b + 8]] = t_pred[order[b:b + 8]].all()
    for flag in (0, 1):
        ids = np.nonzero(use == bool(flag))[0].astype(np.int32)
        hhg.capi._ck(gpu_ctx.L.hhg_set_use_ss(gpu_ctx.h, flag))
        ref, _ = hhg.viterbi_search(gpu_ctx, db, ids=ids)
        got = {h.target: h for h in hits if h.irep == 1}
        for k, t in enumerate(ids):
            assert bits(np.float32(got[int(t)].score)) == bits(ref[k]["hit_score"])
    db.close()
