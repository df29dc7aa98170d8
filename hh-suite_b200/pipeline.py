"""One HHblits-style search iteration on the hot path (src/hhblits.cpp:1118-1221): two-stage cs219
prefilter over the whole shard, then Viterbi (with alternative alignments) on the survivors.
Everything between the stages is the reference's selection logic (prefilter.py, runner.py)."""
from __future__ import annotations

import numpy as np

from . import capi, prefilter, runner


def search(ctx: capi.Context, db: capi.TargetDB, csdb: capi.CsDB, q_p, q_tr, q_pav, lib219, q_prefilter_p=None,
           altali=4, smin=20.0, **pf_kwargs):
    """q_p/q_tr: prepared query (Viterbi); q_prefilter_p: HMM::p of the prefilter-pseudocount copy of the query
    (q_tmp, src/hhblits.cpp:1149-1163; defaults to q_p).  Returns (survivor ids, list of runner.Hit)."""
    prof = capi.build_prefilter_profile(q_p if q_prefilter_p is None else q_prefilter_p, q_pav, lib219,
                                        pf_kwargs.get("score_offset", 50), pf_kwargs.get("bit_factor", 4))
    ids = prefilter.prefilter_db(csdb, prof, **pf_kwargs)
    ctx.set_query(q_p, q_tr)
    hits = runner.ViterbiRunner(ctx, db, altali=altali, smin=smin).alignment(ids) if len(ids) else []
    return ids, hits


def translate_cs219(p_cols: np.ndarray, pav_like: np.ndarray, lib219: np.ndarray) -> np.ndarray:
    """Nearest column state for synthetic data: argmax_k sum_a p[a] * lib[k][a] / bg[a] (the same score
    the prefilter profile is built from).  A stand-in for the reference's cstranslate on synthetic shards."""
    w = (lib219 / pav_like[None, :]).astype(np.float32)        # [219, 20]
    return np.argmax(p_cols.astype(np.float32) @ w.T, axis=1).astype(np.uint8)
