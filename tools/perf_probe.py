"""Developer probe: forward-kernel time of the bench workload under different env knobs.
usage: python tools/perf_probe.py [--targets N] [--lq L] KEY=VAL,KEY=VAL ...   (one run per argument)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hhsuite_b200 as hh  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def main():
    args = sys.argv[1:]
    n, lq = 100000, 400
    cfgs = []
    while args:
        a = args.pop(0)
        if a == "--targets":
            n = int(args.pop(0))
        elif a == "--lq":
            lq = int(args.pop(0))
        else:
            cfgs.append(a)
    qp, qtr, qss, qpav, qcols = synth.query_profile(lq, 1)
    db_h = synth.prepared_db(n, seed=1000, query_cols=qcols, planted=64, fast=True)
    ref = None
    for cfg in cfgs or [""]:
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
        if os.environ.get("HHG_LIB"):
            hh.capi._lib = None          # reload: a different kernel variant
        ctx = hh.Context(device=0)
        ctx.set_query(qp, qtr)
        db = hh.TargetDB(ctx, db_h["L"], db_h["p"], db_h["tr"], db_h["p_off"], db_h["tr_off"])
        plan = hh.Plan(ctx, db)
        for _ in range(2):
            plan.run()
        ctx.sync()
        t = [plan.run_timed() for _ in range(3)]
        ms = float(np.mean([a for a, b in t]))
        hits, _ = plan.fetch(want_paths=False)
        chk = int(hits["score"].view(np.uint32).astype(np.uint64).sum() + hits["i2"].sum() * 7 + hits["j2"].sum() * 13)
        if ref is None:
            ref = chk
        print(f"{cfg:40s} viterbi {ms:8.3f} ms  {plan.cells / ms / 1e6:8.1f} GCUPS  bt {np.mean([b for a, b in t]):.3f} ms "
              f"checksum {'same' if chk == ref else 'DIFFERENT'}", flush=True)
        plan.close(); db.close(); ctx.close()
        for kv in filter(None, cfg.split(",")):
            os.environ.pop(kv.split("=")[0], None)


if __name__ == "__main__":
    main()
