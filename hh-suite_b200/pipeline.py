"""One HHblits-style search iteration on the hot path (src/hhblits.cpp:1118-1221): two-stage cs219
prefilter over the whole shard, then Viterbi (with alternative alignments) on the survivors.
Everything between the stages is the reference's selection logic (prefilter.py, runner.py)."""
from __future__ import annotations

import numpy as np

from . import capi, prefilter, runner


def search(ctx: capi.Context, db: capi.TargetDB, csdb: capi.CsDB, q_p, q_tr, q_pav, lib219, q_prefilter_p=None,
           altali=4, smin=20.0, cs_names=None, db_names=None, **pf_kwargs):
    """q_p/q_tr: prepared query (Viterbi); q_prefilter_p: HMM::p of the prefilter-pseudocount copy of the query
    (q_tmp, src/hhblits.cpp:1149-1163; defaults to q_p).  Returns (survivor ids in the TARGET shard, list of runner.Hit).

    The prefilter's survivors are entries of the cs219 index; the reference carries them to the Viterbi stage BY NAME
    (src/hhprefilter.cpp:561-590, then a lookup in the hhm/a3m index), because the two indices of a real database
    need not list the same entries in the same order.  Pass both name lists (cs_names[k] = name of cs219 sequence k,
    db_names[t] = name of target t) to map by name; without them the two shards must be index-aligned, which is
    checked as far as it can be (same number of entries, same lengths)."""
    prof = capi.build_prefilter_profile(q_p if q_prefilter_p is None else q_prefilter_p, q_pav, lib219,
                                        pf_kwargs.get("score_offset", 50), pf_kwargs.get("bit_factor", 4))
    ids = prefilter.prefilter_db(csdb, prof, **pf_kwargs)
    if cs_names is not None or db_names is not None:
        if cs_names is None or db_names is None:
            raise ValueError("pass both cs_names and db_names, or neither")
        where = {nm: t for t, nm in enumerate(db_names)}
        missing = [cs_names[k] for k in ids if cs_names[k] not in where]
        if missing:
            raise KeyError(f"{len(missing)} prefilter hits have no entry in the profile shard, e.g. {missing[0]!r}")
        ids = np.array([where[cs_names[k]] for k in ids], np.int32)
    elif csdb.n != db.n or not np.array_equal(np.asarray(csdb.Lh), np.asarray(db.Lh)):
        raise ValueError("cs219 shard and profile shard are not index-aligned (different sizes or lengths): "
                         "pass cs_names / db_names so the survivors are mapped by name like the reference does")
    ctx.set_query(q_p, q_tr)
    hits = runner.ViterbiRunner(ctx, db, altali=altali, smin=smin).alignment(ids) if len(ids) else []
    return ids, hits


def translate_cs219(p_cols: np.ndarray, pav_like: np.ndarray, lib219: np.ndarray) -> np.ndarray:
    """Nearest column state for synthetic data: argmax_k sum_a p[a] * lib[k][a] / bg[a] (the same score
    the prefilter profile is built from).  A stand-in for the reference's cstranslate on synthetic shards."""
    w = (lib219 / pav_like[None, :]).astype(np.float32)        # [219, 20]
    return np.argmax(p_cols.astype(np.float32) @ w.T, axis=1).astype(np.uint8)
