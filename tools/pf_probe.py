"""Developer probe: ungapped prefilter throughput (cells = Lq * sum L) on a synthetic cs219 shard."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hhsuite_b200 as hh  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def _host_select(hh, db, lq):
    raw = db.fetch()
    L = hh.capi.load()
    corr = np.zeros(db.n, np.int32)
    hh.capi._ck(L.hhg_prefilter_corrected_scores(db.n, hh.capi._p(np.ascontiguousarray(raw, np.int32), hh.capi.c_i32p),
                                                hh.capi._p(np.ascontiguousarray(db.Lh, np.int32), hh.capi.c_i32p), lq, 4,
                                                hh.capi._p(corr, hh.capi.c_i32p)))
    order = np.lexsort((np.arange(db.n), corr))[::-1]
    stop = np.nonzero(corr[order[100:]] <= 10)[0]
    ncut = 100 + int(stop[0]) if len(stop) else db.n
    first = order[:ncut]
    return first, corr[first]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    lq = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    import torch
    cs = synth.cs219_db(n, seed=3)
    rng = np.random.default_rng(1)
    prof = rng.integers(35, 70, (220, lq), dtype=np.uint8)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    ctx = hh.Context(device=0, stream=st.cuda_stream)
    db = hh.CsDB(ctx, cs["L"], cs["off"], cs["seq"])
    db.run(prof, 50, upload=True)
    ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        db.run(prof, 50, upload=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    cells = float(lq) * float(cs["L"].sum())
    sc = db.fetch()
    print(f"prefilter ungapped: n={n} Lq={lq} {ms:.3f} ms  {cells / ms / 1e9:.2f} Tcells/s  "
          f"({float(cs['L'].sum()) / ms / 1e6:.1f} GB/s of cs219 bytes) max score {sc.max()} mean {sc.mean():.1f}")
    # a real query profile (golden: data/query.hhm, Lq=431) so that scores spread like a real search
    G = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_v1.npz"))
    rprof = np.ascontiguousarray(G["pf_prof"])
    lq_r = rprof.shape[1]
    db.run(rprof, 50, upload=True)
    ctx.sync()
    lq_saved, lq = lq, lq_r
    # stage-1 selection: device (histogram + compaction) vs host (fetch N scores, correct, sort), wall clock
    for name, fn in (("device select", lambda: db.select(lq, 4, 10, 100)),
                     ("host fetch+sort", lambda: _host_select(hh, db, lq))):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            r = fn()
        dt = (time.perf_counter() - t0) / 3
        print(f"stage-1 selection (real profile Lq={lq}), {name}: {dt * 1e3:.2f} ms, {len(r[0])} survivors")
    lq = lq_saved
    db.run(prof, 50, upload=True)
    ctx.sync()
    try:
        from oracle.binding import RefShim
        R = RefShim()
        qp, qtr, qss, qpav, _ = synth.query_profile(lq, 1)
        R.set_query(qp, qtr, qpav, None)
        # striped profile in the reference's layout from our linear profile
        W = (lq + 31) // 32
        qc = np.full(220 * (lq + 64), 50, np.uint8)
        pos = np.arange(lq)
        for k in range(220):
            qc[k * W * 32 + (pos % W) * 32 + pos // W] = prof[k]
        m = min(n, 200000)
        sub = dict(L=cs["L"][:m], off=cs["off"][:m], seq=cs["seq"])
        thr = os.cpu_count()
        sec, ref = R.ungapped_bench(qc, sub, thr)
        assert np.array_equal(ref, sc[:m]), "GPU != reference scores"
        print(f"reference ungapped_sse_score on {thr} threads, {m} seqs: {sec*1e3:.1f} ms  "
              f"{lq * float(sub['L'].sum()) / sec / 1e12:.3f} Tcells/s; scores identical")
    except (FileNotFoundError, OSError) as e:
        print("reference not available:", e)


if __name__ == "__main__":
    main()
