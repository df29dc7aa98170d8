"""Query preparation of one hhblits iteration, timed stage by stage next to the compiled reference (1 thread):
alignment -> HMM (hhg_msa_to_hmm), context-specific pseudocounts for the HMM and for the prefilter profile
(hhg_query_context_pseudocounts), prefilter byte profile (hhg_prefilter_build_profile).
    python tools/query_prep_probe.py"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hhsuite_b200 as hhg  # noqa: E402
from hhsuite_b200 import capi, synth  # noqa: E402
from oracle.binding import RefShim  # noqa: E402


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


def main():
    r = RefShim(nocontxt=True, maxres=4096)
    pb, R = r.pb(), r.R()
    ctx = hhg.Context()
    crf = capi.Crf(ctx, r.crf_text())
    d = tempfile.mkdtemp()
    qa = os.path.join(ROOT, "oracle", "_ref", "data", "query.a3m")
    cases = [("synthetic L=400 N=300", synth.a3m_text(400, 300, 5, ident=0.5).encode())]
    if os.path.exists(qa):
        cases.append(("data/query.a3m L=431 N=59", open(qa, "rb").read()))
    for name, a3m in cases:
        path = os.path.join(d, "q.a3m")
        open(path, "wb").write(a3m)
        t_msa, raw = best(lambda: capi.msa_to_hmm(ctx, a3m, pb))
        t_hmm, (p, pav) = best(lambda: crf.pseudocounts(raw["f"], raw["neff_m"], raw["neff_hmm"], pb, capi.Admix.hhm()))
        t_pf, (ppf, pavpf) = best(lambda: crf.pseudocounts(raw["f"], raw["neff_m"], raw["neff_hmm"], pb, capi.Admix.prefilter()))
        t_tr, q = best(lambda: capi.query_from_a3m(ctx, a3m, R, pb))
        rt_msa, ref = best(lambda: r.msa_to_hmm(path, capL=1000, capN=2000), 3)
        rt_pc, (rp, rpav) = best(lambda: r.context_pc(ref["f"], ref["neff_m"], ref["neff_hmm"], engine=0), 3)
        same = np.array_equal(p.view(np.uint32), rp.view(np.uint32)) and np.array_equal(raw["f"].view(np.uint32), ref["f"].view(np.uint32))
        print(f"{name}: library  alignment->HMM {t_msa:.2f} ms | context pc (HMM) {t_hmm:.2f} ms | context pc (prefilter) {t_pf:.2f} ms"
              f" | transitions+nocontxt path {t_tr:.2f} ms || reference (OpenMP as built, {os.cpu_count()} cpus visible)"
              f" alignment->HMM {rt_msa:.2f} ms | context pc {rt_pc:.2f} ms || identical: {same}")
    crf.close(); ctx.close()


if __name__ == "__main__":
    main()
