"""Throughput of hhg_mac_realign on a realistic batch: query L=400, the 500 best hits of a 20k-target shard
(BASELINE configs[1] shape), next to the compiled reference's PosteriorDecoder::realign on one host thread.
    python tools/mac_probe.py [n_hits]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hhsuite_b200 as hh  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def main():
    nh = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    lq = 400
    qp, qtr, qss, qpav, qcols = synth.query_profile(lq, 1)
    db_h = synth.prepared_db(20000, seed=1000, query_cols=qcols, planted=600, fast=True)
    ctx = hh.Context()
    ctx.set_query(qp, qtr)
    db = hh.TargetDB(ctx, db_h["L"], db_h["p"], db_h["tr"], db_h["p_off"], db_h["tr_off"])
    hits, paths = hh.viterbi_search(ctx, db)
    order = np.argsort(-hits["hit_score"])[:nh]
    order = np.array([t for t in order if hits["nsteps"][t] > 0], np.int32)
    vits = []
    for t in order:
        i_s, j_s, st = hh.expand_path(hits[t], paths)
        vits.append((int(hits["i1"][t]), int(hits["i2"][t]), int(hits["j1"][t]), int(hits["j2"][t]), int(hits["nsteps"][t]), i_s, j_s))
    hh.capi.mac_query_set(ctx, qp, hh.capi.log2lin(qtr))
    hh.capi.mac_realign(ctx, db, order, vits)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        mh, mp = hh.capi.mac_realign(ctx, db, order, vits)
        best = min(best, time.perf_counter() - t0)
    cells = float(lq) * float(db.Lh[order].sum())
    print(f"hhg_mac_realign: {len(order)} hits, {cells / 1e6:.1f} M cells (Lq x Lt), {best * 1e3:.2f} ms wall "
          f"({cells / best / 1e9:.2f} Gcells/s); aligned pairs: {int(mh['matched_cols'].sum())}, "
          f"mean sum_of_probs {float(mh['sum_of_probs'].mean()):.2f}")
    try:
        from oracle.binding import RefShim
        R = RefShim(nocontxt=True, maxres=4096)
        R.set_query(qp, qtr, qpav, None)
        m = min(len(order), 40)
        t0 = time.perf_counter()
        same = 0
        for r in range(m):
            t = int(order[r])
            L = int(db_h["L"][t])
            tp = db_h["p"][db_h["p_off"][t]:db_h["p_off"][t] + L + 2]
            ttr = db_h["tr"][db_h["tr_off"][t]:db_h["tr_off"][t] + L + 1]
            ref = R.mac_realign(tp, ttr, vits[r], want_post=False)
            same += int(ref["nsteps"] == int(mh["nsteps"][r]) and ref["Pforward"] == float(mh["pforward"][r])
                        and np.array_equal(ref["i"][1:], mp[r]["i"][1:]))
        dt = time.perf_counter() - t0
        c = float(lq) * float(db.Lh[order[:m]].sum())
        print(f"reference PosteriorDecoder::realign, 1 thread: {m} hits {dt * 1e3:.1f} ms ({c / dt / 1e9:.3f} Gcells/s); "
              f"{same}/{m} hits identical (path, Pforward)")
    except (FileNotFoundError, OSError) as e:
        print("reference not available:", e)


if __name__ == "__main__":
    main()
