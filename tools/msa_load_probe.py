"""Throughput of the A3M database loader (hhg_db_create_a3m) next to the reference's per-template alignment branch
(Alignment::Read + Compress + Filter + FrequenciesAndTransitions + PrepareTemplateHMM; compiled reference, 1 thread)
on the same records.
    python tools/msa_load_probe.py [n_records]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hhsuite_b200 as hhg  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rng = np.random.default_rng(1)
    lens = np.clip(np.round(np.exp(rng.normal(np.log(200), 0.5, 32))), 30, 800).astype(int)
    nseq = np.clip(np.round(np.exp(rng.normal(np.log(120), 0.8, 32))), 5, 1500).astype(int)
    uniq = [synth.a3m_text(int(L), int(N), 100 + k, f"u{k}", ident=float(rng.uniform(0.3, 0.8))).encode()
            for k, (L, N) in enumerate(zip(lens, nseq))]
    pick = rng.integers(0, len(uniq), n)
    texts = [uniq[i] for i in pick]
    data = b"".join(t + b"\0" for t in texts)
    ln = np.array([len(t) + 1 for t in texts], np.int64)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
    from oracle.binding import RefShim
    r = RefShim(nocontxt=True, maxres=4096)
    R, pb = r.R(), r.pb()
    ctx = hhg.Context()
    res = int(sum(int(lens[i]) * int(nseq[i]) for i in pick))
    for rep in range(3):
        t0 = time.perf_counter()
        db = hhg.TargetDB.from_a3m(ctx, data, off, ln, R, pb)
        dt = time.perf_counter() - t0
        cols = int(db.Lh.sum())
        db.close()
        print("hhg_db_create_a3m:", dict(records=n, columns=cols, text_MB=round(len(data) / 1e6, 1), residues=res,
                                         seconds=round(dt, 3), records_per_s=round(n / dt, 1),
                                         text_MB_per_s=round(len(data) / 1e6 / dt, 1)))
    if os.environ.get("HHG_TIMING"):
        return
    d = tempfile.mkdtemp()
    paths = []
    for k, t in enumerate(uniq):
        p = os.path.join(d, f"u{k}.a3m")
        open(p, "wb").write(t)
        paths.append(p)
    m = min(n, 300)
    t0 = time.perf_counter()
    for i in pick[:m]:
        r.msa_to_hmm(paths[i], prep=True, capL=1000, capN=2000)
    dt = time.perf_counter() - t0
    print("reference alignment branch + PrepareTemplateHMM, 1 thread:", dict(records=m, seconds=round(dt, 3), records_per_s=round(m / dt, 1)))


if __name__ == "__main__":
    main()
