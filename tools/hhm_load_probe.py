"""Throughput of the HHM-text database loader (hhg_db_create_hhm) next to the reference's per-template
HMM::Read + PrepareTemplateHMM (compiled reference, 1 thread) on the same records.
    python tools/hhm_load_probe.py [n_records]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hhsuite_b200 as hhg  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    rng = np.random.default_rng(1)
    lens = np.clip(np.round(np.exp(rng.normal(np.log(200), 0.5, 64))), 30, 2000).astype(int)
    uniq = [synth.hhm_text(int(L), 100 + k, f"u{k}").encode() for k, L in enumerate(lens)]
    pick = rng.integers(0, len(uniq), n)
    texts = [uniq[i] for i in pick]
    data = b"".join(t + b"\0" for t in texts)
    ln = np.array([len(t) + 1 for t in texts], np.int64)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
    G = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    ctx = hhg.Context()
    out = {}
    for rep in range(3):
        t0 = time.perf_counter()
        db = hhg.TargetDB.from_hhm(ctx, data, off, ln, G["R"])
        dt = time.perf_counter() - t0
        cols = int(db.Lh.sum())
        db.close()
        out = dict(records=n, columns=cols, text_MB=len(data) / 1e6, seconds=dt, records_per_s=n / dt,
                   text_MB_per_s=len(data) / 1e6 / dt, columns_per_s=cols / dt)
    print("hhg_db_create_hhm:", out)
    try:
        from oracle.binding import RefShim
        r = RefShim(nocontxt=True, maxres=4096)
        import tempfile
        d = tempfile.mkdtemp()
        paths = []
        for k, t in enumerate(uniq):
            p = os.path.join(d, f"u{k}.hhm")
            open(p, "wb").write(t)
            paths.append(p)
        qp = os.path.join(ROOT, "oracle", "_ref", "data", "query.hhm")
        r.load_query_hhm(qp if os.path.exists(qp) else paths[0])
        m = min(n, 1000)
        t0 = time.perf_counter()
        for i in pick[:m]:
            r.lib.hhref_prepare_template_hhm  # noqa: B018  (resolve once)
            r.prepare_template_hhm(paths[i])
        dt = time.perf_counter() - t0
        print("reference HMM::Read + PrepareTemplateHMM, 1 thread:", dict(records=m, seconds=dt, records_per_s=m / dt))
    except Exception as e:  # the compiled reference is optional
        print("reference arm unavailable:", e)


if __name__ == "__main__":
    main()
