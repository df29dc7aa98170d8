"""BASELINE config 3 shape at test scale: prefilter (both stages) + Viterbi on the survivors."""
import numpy as np
import pytest

from tests.util import golden

pytestmark = pytest.mark.gpu


def test_prefilter_then_viterbi_finds_planted_homologs(hhg, gpu_ctx):
    from hhsuite_b200 import synth
    G = golden()
    lib = G["cs219_lin"]
    n, planted = 6000, 30
    qp, qtr, qss, qpav, qcols = synth.query_profile(300, 5)
    db_h = synth.prepared_db(n, seed=8, query_cols=qcols, planted=planted, fast=True, hi=600)
    # scatter the planted targets through the shard so their ids are not special
    perm = np.random.default_rng(1).permutation(n)
    prof_list = []
    for t in perm:
        L = int(db_h["L"][t])
        prof_list.append((db_h["p"][db_h["p_off"][t]:db_h["p_off"][t] + L + 2],
                          db_h["tr"][db_h["tr_off"][t]:db_h["tr_off"][t] + L + 1], None))
    planted_ids = set(np.nonzero(perm < planted)[0].tolist())
    db = hhg.TargetDB.from_profiles(gpu_ctx, prof_list)
    # cs219 sequences: nearest column state of every target column (null-model ratio * background = prob)
    bg = synth._PB.astype(np.float32)
    seqs = [hhg.pipeline.translate_cs219(p[1:-1] * bg[None, :], bg, lib) for (p, tr, ss) in prof_list]
    L = np.array([len(s) for s in seqs], np.int32)
    off = np.concatenate([[0], np.cumsum(L.astype(np.int64))[:-1]])
    csdb = hhg.CsDB(gpu_ctx, L, off, np.concatenate(seqs))
    ids, hits = hhg.pipeline.search(gpu_ctx, db, csdb, qp, qtr, qpav, lib, min_prefilter_hits=50, maxnumdb=500)
    assert 50 <= len(ids) <= 500
    found = planted_ids & set(ids.tolist())
    assert len(found) >= 0.9 * planted, (len(found), planted)
    first = [h for h in hits if h.irep == 1]
    assert len(first) == len(ids)
    ranked = sorted(first, key=lambda h: -h.score)
    top = {h.target for h in ranked[:len(found)]}
    assert len(top & found) >= 0.9 * len(found)
    # the prefilter is a filter, not a scorer: every survivor was aligned exactly once per pass
    assert sorted(h.target for h in first) == sorted(ids.tolist())
    db.close(); csdb.close()
