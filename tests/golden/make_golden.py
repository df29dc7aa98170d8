"""Generate tests/golden/golden_v1.npz from the UNMODIFIED compiled reference (oracle/_ref).

Run in the authoring container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
Every array below is produced by reference code paths:
  * HMM::Read + PrepareQueryHMM / PrepareTemplateHMM   -> prepared fp32 profiles
  * Viterbi::Align (AVX2, no FMA)                      -> score, i2, j2, backtrace bytes
  * Viterbi::Backtrace / ScoreForBacktrace             -> path, Hit.score
  * Prefilter::stripe_query_profile / ungapped_sse_score / swStripedByte
The synthetic inputs come from hh-suite_b200/synth.py with fixed seeds.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import RefShim  # noqa: E402
import hhsuite_b200  # noqa: E402
from hhsuite_b200 import synth  # noqa: E402

REFDATA = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    R = RefShim(nocontxt=True, maxres=4096)
    G = {}
    # ---- config 1: data/query.hhm vs one synthetic HMM (L=150), plus self and reversed-self
    q = R.load_query_hhm(os.path.join(REFDATA, "query.hhm"))
    G["q_p"], G["q_tr"], G["q_pav"], G["q_ss"] = q["p"], q["tr"], q["pav"], q["ss"]
    txt = synth.hhm_text(150, 7, "synth150")
    open("/tmp/synth150.hhm", "w").write(txt)
    G["synth150_hhm_sha"] = np.frombuffer(hashlib.sha256(txt.encode()).digest(), np.uint8)
    t = R.prepare_template_hhm("/tmp/synth150.hhm")
    ts = R.prepare_template_hhm(os.path.join(REFDATA, "query.hhm"))
    rev = dict(p=np.ascontiguousarray(np.concatenate([ts["p"][:1], ts["p"][1:-1][::-1], ts["p"][-1:]])),
               tr=ts["tr"], ss=ts["ss"])
    for name, tt in (("t150", t), ("tself", ts), ("trev", rev)):
        G[name + "_p"], G[name + "_tr"] = tt["p"], tt["tr"]
    res = R.viterbi([(t["p"], t["tr"], None), (ts["p"], ts["tr"], None), (rev["p"], rev["tr"], None)])
    for k, name in enumerate(("t150", "tself", "trev")):
        sc, i2, j2, bt = res[k]
        n, i_s, j_s, st, mc = R.backtrace(k)
        hs, hss = R.hit_score(k)
        G[name + "_res"] = np.array([i2, j2, n, mc, i_s[n], j_s[n]], np.int32)
        G[name + "_score"] = np.array([sc, hs], np.float32)
        G[name + "_states"] = st[1:]
        G[name + "_bt_sha"] = sha(bt[1:, 1:])
        if name == "t150":
            G[name + "_bt"] = bt
    # ---- variants on small synthetic prepared profiles: SS, cell-off, global, ragged 8-lane batch
    qp, qtr, qss, qpav, _ = synth.query_profile(120, 5)
    G["v_q_p"], G["v_q_tr"], G["v_q_ss"], G["v_q_pav"] = qp, qtr, qss, qpav
    R.set_query(qp, qtr, qpav, qss)
    G["S33"] = R.S33()
    rng = np.random.default_rng(11)
    lens = [100, 77, 64, 33, 150, 31, 8, 1]
    tg = [synth.prepared_profile(L, rng) for L in lens]
    for k, (p, tr, ss) in enumerate(tg):
        G[f"v_t{k}_p"], G[f"v_t{k}_tr"], G[f"v_t{k}_ss"] = p, tr, ss
    co = []
    for (p, tr, ss) in tg:
        m = np.zeros((121, p.shape[0] - 1), np.uint8)
        m[40:60, :] = 1
        m[:, 20:25] = 1
        m[(np.add.outer(np.arange(121), np.arange(p.shape[0] - 1)) % 17) == 0] = 1
        co.append(m)
    G["v_celloff_rows"] = np.array([40, 60, 20, 25, 17], np.int32)
    variants = dict(plain={}, ss=dict(use_ss=True), co=dict(celloff=co), co_ss=dict(celloff=co, use_ss=True),
                    glob=dict(local=False, egq=0.1, egt=0.2))
    for vn, kw in variants.items():
        hs = []
        if vn == "glob":
            # global mode is only well defined by the reference for equal-length lanes: one target per call
            out = []
            for k in range(len(tg)):
                out.append(R.viterbi([tg[k]], **kw)[0])
                hs.append(R.hit_score(0))
        else:
            out = R.viterbi(tg, **kw)
            hs = [R.hit_score(k) for k in range(len(tg))]
        for k, (sc, i2, j2, bt) in enumerate(out):
            G[f"v_{vn}_{k}_res"] = np.array([i2, j2], np.int32)
            G[f"v_{vn}_{k}_score"] = np.array([sc, hs[k][0], hs[k][1]], np.float32)
            G[f"v_{vn}_{k}_bt"] = bt
    # ---- fast_log2 table and the query-dependent template preparation (null model)
    G["fastlog2_lg2"] = R.fast_log2_table()
    q = R.load_query_hhm(os.path.join(REFDATA, "query.hhm"))
    for cs in (0, 1, 2, 3):
        for name, path in (("t150", "/tmp/synth150.hhm"), ("tself", os.path.join(REFDATA, "query.hhm"))):
            t = R.prepare_template_hhm_raw(path, cs)
            if cs == 1:
                G[f"nm_{name}_praw"], G[f"nm_{name}_tr"], G[f"nm_{name}_pav"] = t["p_raw"], t["tr"], t["pav"]
            G[f"nm_{name}_p_cs{cs}"] = t["p"]
    pb = np.zeros(20, np.float32)
    R.lib.hhref_get_pb(pb.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)))
    G["pb"] = pb
    # ---- HHM text -> prepared records (HMM::Read + the query-independent part of PrepareTemplateHMM)
    G["R"] = R.R()
    pp = R.prep_params()
    G["prep_params"] = np.array([getattr(pp, f[0]) for f in pp._fields_], np.float32)
    txt_ss = synth.hhm_text(60, 11, "ss60", with_ss=True)
    open("/tmp/ss60.hhm", "w").write(txt_ss)
    G["hhm_t150_text"] = np.frombuffer(txt.encode(), np.uint8)
    G["hhm_ss60_text"] = np.frombuffer(txt_ss.encode(), np.uint8)
    t = R.prepare_template_hhm_raw("/tmp/ss60.hhm", 1)
    t2 = R.prepare_template_hhm("/tmp/ss60.hhm")
    G["hhm_ss60_praw"], G["hhm_ss60_tr"], G["hhm_ss60_pav"], G["hhm_ss60_ss"] = t["p_raw"], t["tr"], t["pav"], t2["ss"]
    # ---- MAC realignment (PosteriorDecoder::realign) of the config-1 hits around their Viterbi alignments
    q = R.load_query_hhm(os.path.join(REFDATA, "query.hhm"))
    for name, path, mact in (("t150", "/tmp/synth150.hhm", 0.0), ("tself", os.path.join(REFDATA, "query.hhm"), 0.35)):
        tt = R.prepare_template_hhm(path)
        res = R.viterbi([(tt["p"], tt["tr"], None)])
        sc, i2, j2, bt = res[0]
        n, i_s, j_s, st, mc = R.backtrace(0)
        m = R.mac_realign(tt["p"], tt["tr"], (int(i_s[n]), i2, int(j_s[n]), j2, n, i_s, j_s), mact=mact)
        G[f"mac_{name}_vit"] = np.array([int(i_s[n]), i2, int(j_s[n]), j2, n], np.int32)
        G[f"mac_{name}_vit_i"], G[f"mac_{name}_vit_j"] = i_s[:n + 1].astype(np.int32), j_s[:n + 1].astype(np.int32)
        G[f"mac_{name}_res"] = np.array([m["i1"], m["i2"], m["j1"], m["j2"], m["nsteps"], m["matched_cols"]], np.int32)
        G[f"mac_{name}_f"] = np.array([m["sum_of_probs"], mact], np.float32)
        G[f"mac_{name}_pforward"] = np.array([m["Pforward"]], np.float64)
        G[f"mac_{name}_i"], G[f"mac_{name}_j"], G[f"mac_{name}_states"] = m["i"], m["j"], m["states"]
        G[f"mac_{name}_ppost"] = m["P_posterior"]
        G[f"mac_{name}_post_sha"] = sha(m["post"][1:, 1:])
        if name == "tself":      # second MAC alignment of the same template: the first one is excluded
            m2 = R.mac_realign(tt["p"], tt["tr"], (int(i_s[n]), i2, int(j_s[n]), j2, n, i_s, j_s), mact=mact,
                               excl=[(m["i"][1:], m["j"][1:])])
            G["mac_tself2_res"] = np.array([m2["i1"], m2["i2"], m2["j1"], m2["j2"], m2["nsteps"], m2["matched_cols"]], np.int32)
            G["mac_tself2_pforward"] = np.array([m2["Pforward"]], np.float64)
            G["mac_tself2_i"], G["mac_tself2_j"], G["mac_tself2_ppost"] = m2["i"], m2["j"], m2["P_posterior"]
        if name == "t150":
            G["mac_q_tr_lin"] = m["q_tr_lin"]
    # ---- prefilter
    lib = R.cs219()
    G["cs219_lin"] = lib
    qc, W = R.stripe_query_profile(50, 4)
    Lq = q["L"]
    # un-stripe the reference profile: qc[k*W*32 + (pos % W)*32 + pos // W]  
    prof = np.zeros((220, Lq), np.uint8)
    pos = np.arange(Lq)
    for k in range(220):
        prof[k] = qc[k * W * 32 + (pos % W) * 32 + pos // W]
    G["pf_prof"] = prof
    rng = np.random.default_rng(3)
    seqs = [rng.integers(0, 219, L, dtype=np.uint8) for L in (30, 64, 200, 431, 777, 1)]
    # a planted near-match: the argmax state per query column
    best = prof[:219].argmax(axis=0).astype(np.uint8)
    seqs.append(best[50:250].copy())
    seqs.append(best.copy())
    G["pf_nseq"] = np.array([len(seqs)], np.int32)
    for k, s in enumerate(seqs):
        G[f"pf_seq{k}"] = s
        G[f"pf_ref{k}"] = np.array([R.ungapped(qc, s, 50), R.sw_byte(qc, s, 24, 4, 50)], np.int32)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(G), "arrays")
    hitlist_goldens(R)


def hitlist_goldens(R):
    """hitlist_v1.npz: HitList::CalculatePvalues / CalculateHHblitsEvalues / SortList of the compiled reference on
    the seeded synthetic hit lists of tests/test_hitlist_cpu.py (same `cases()`)."""
    from tests.test_hitlist_cpu import cases
    H = {}
    for i, c in enumerate(cases()):
        hb = c["hb"] or {}
        ref = R.hitlist_stats(c["score"], c["score_ss"], c["L"], c["neff"], c["qL"], c["qneff"], c["N"], c["loc"], c["ssm"],
                              c["ssw"], c["ssm2"], c["files"], hhblits=c["hb"] is not None, dbsize=hb.get("dbsize", 1),
                              pf_evalue_thresh=hb.get("thresh", 1.0))
        for f, v in ref.items():
            H[f"c{i}_{f}"] = v
    out = os.path.join(os.path.dirname(OUT), "hitlist_v1.npz")
    np.savez_compressed(out, **H)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
