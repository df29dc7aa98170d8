"""Query-batch mode vs one query at a time in the latency-bound regime: Q queries (L=400), each against its own 3000
survivors of a 200k-target shard.   python tools/batch_probe.py [Q]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hhsuite_b200 as hh  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rng = np.random.default_rng(1)
    queries = [synth.query_profile(400, 10 + k) for k in range(Q)]
    db_h = synth.prepared_db(200000, seed=1000, fast=True)
    ctx = hh.Context()
    db = hh.TargetDB(ctx, db_h["L"], db_h["p"], db_h["tr"], db_h["p_off"], db_h["tr_off"])
    surv = [rng.choice(200000, 3000, replace=False).astype(np.int32) for _ in range(Q)]
    cells = sum(400.0 * float(db_h["L"][s].sum()) for s in surv)

    def sequential():
        for q in range(Q):
            ctx.set_query(queries[q][0], queries[q][1])
            hh.viterbi_search(ctx, db, ids=surv[q])

    req_q = np.concatenate([np.full(3000, q, np.int32) for q in range(Q)])
    ids = np.concatenate(surv)

    def batched():
        hh.capi.query_set_batch(ctx, [(q[0], q[1], q[2]) for q in queries])
        hh.capi.viterbi_search_batch(ctx, db, req_q, ids)

    for name, fn in (("one query at a time", sequential), ("one batch", batched)):
        fn(); fn()
        t = time.perf_counter()
        for _ in range(3):
            fn()
        dt = (time.perf_counter() - t) / 3
        print(f"{name:22s}: {Q} queries x 3000 targets: {dt * 1e3:7.2f} ms wall  {cells / dt / 1e9:7.1f} GCUPS", flush=True)


if __name__ == "__main__":
    main()
