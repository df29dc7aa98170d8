"""Host-side mirror of the reference's ViterbiRunner (src/hhviterbirunner.{h,cpp}) on top of the C-ABI.

``ViterbiRunner.alignment`` reproduces the control flow of ViterbiRunner::alignment
(src/hhviterbirunner.cpp:75-210) for the part that lives on the hot path:

  * pass 0 aligns every requested target; pass k>0 ("alternative alignments", par.altali, default 4)
    re-aligns only the targets whose previous hit scored above par.smin (default 20), with ALL of that
    target's earlier paths masked by the +-40 cross of Viterbi::ExcludeAlignment
    (merge_thread_results :249-271, exclude_alignments :273-289);
  * every (target, pass) yields one hit with irep = pass+1 (:257) and lastrep = (score <= smin) (:36);
  * hit.score is Hit.score (ScoreForBacktrace), computed on the device.

The reference batches 8 targets per SIMD call and sorts each chunk by length purely for speed; results do
not depend on it, and neither do ours.  Not mirrored here (documented in DESIGN.md): the hhblits
early-stopping filter (:178-188) and the SS-mode consensus over the 8 lanes of a batch (:14-22), which
make the reference's results depend on batch composition."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class Hit:
    target: int          # index into the shard (the reference carries an HHEntry*)
    irep: int            # 1-based index of the alternative alignment (Hit::irep)
    lastrep: int         # 1 if score <= smin (Hit::lastrep)
    score: float         # Hit.score (with correlation term)
    score_ss: float
    vit_score: float     # raw Viterbi score (ViterbiResult::score)
    i1: int
    i2: int
    j1: int
    j2: int
    nsteps: int
    matched_cols: int
    i: np.ndarray = field(repr=False, default=None)       # i_steps[1..nsteps]
    j: np.ndarray = field(repr=False, default=None)
    states: np.ndarray = field(repr=False, default=None)


class ViterbiRunner:
    def __init__(self, ctx: capi.Context, db: capi.TargetDB, altali: int = 4, smin: float = 20.0):
        self.ctx, self.db = ctx, db
        self.altali, self.smin = altali, smin

    def alignment(self, ids=None) -> list[Hit]:
        ids = np.arange(self.db.n, dtype=np.int32) if ids is None else np.ascontiguousarray(ids, np.int32)
        todo = ids
        excl: dict[int, list[tuple[np.ndarray, np.ndarray]]] = {}
        out: list[Hit] = []
        for rep in range(self.altali):
            if len(todo) == 0:
                break
            exclusions = None
            if rep > 0:
                exclusions = []
                for t in todo:
                    ii = np.concatenate([e[0] for e in excl[int(t)]])
                    jj = np.concatenate([e[1] for e in excl[int(t)]])
                    exclusions.append((ii, jj))
            hits, paths = capi.viterbi_search(self.ctx, self.db, ids=todo, exclusions=exclusions)
            nxt = []
            for k, t in enumerate(todo):
                h = hits[k]
                gi, gj, gs = capi.expand_path(h, paths)
                n = int(h["nsteps"])
                out.append(Hit(int(t), rep + 1, 1 if h["hit_score"] <= self.smin else 0, float(h["hit_score"]),
                               float(h["score_ss"]), float(h["score"]), int(h["i1"]), int(h["i2"]), int(h["j1"]),
                               int(h["j2"]), n, int(h["matched_cols"]), gi, gj, gs))
                if h["hit_score"] > self.smin:
                    nxt.append(int(t))
                    # ExcludeAlignment masks steps 1 <= step < nsteps (src/hhviterbi.cpp:66)
                    excl.setdefault(int(t), []).append((gi[1:n].copy(), gj[1:n].copy()))
            todo = np.array(nxt, np.int32)
        return out
