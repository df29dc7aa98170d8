"""CPU tests: the C oracle (oracle/hh_oracle.c) against the committed reference goldens
(tests/golden/golden_v1.npz, produced by the compiled reference via tests/golden/make_golden.py)."""
import hashlib

import numpy as np

from tests.util import bits, golden


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_config1_query_hhm_goldens(oracle):
    G = golden()
    for n in ("t150", "tself", "trev"):
        sc, i2, j2, bt = oracle.viterbi(G["q_p"], G["q_tr"], G[n + "_p"], G[n + "_tr"])
        gi2, gj2, nsteps, mc, i1, j1 = G[n + "_res"]
        assert bits(sc) == bits(G[n + "_score"][0])
        assert (i2, j2) == (gi2, gj2)
        assert np.array_equal(_sha(bt[1:, 1:]), G[n + "_bt_sha"])
        k, i_s, j_s, st, m = oracle.backtrace(bt, i2, j2)
        assert (k, m, i_s[k], j_s[k]) == (nsteps, mc, i1, j1)
        assert np.array_equal(st[1:], G[n + "_states"])
    assert np.array_equal(oracle.viterbi(G["q_p"], G["q_tr"], G["t150_p"], G["t150_tr"])[3][1:, 1:], G["t150_bt"][1:, 1:])


def test_variant_goldens(oracle):
    G = golden()
    co_rows = G["v_celloff_rows"]
    for k in range(8):
        tp, ttr, tss = G[f"v_t{k}_p"], G[f"v_t{k}_tr"], G[f"v_t{k}_ss"]
        Lt = tp.shape[0] - 2
        m = np.zeros((121, Lt + 1), np.uint8)
        m[co_rows[0]:co_rows[1], :] = 1
        m[:, co_rows[2]:co_rows[3]] = 1
        m[(np.add.outer(np.arange(121), np.arange(Lt + 1)) % co_rows[4]) == 0] = 1
        variants = dict(plain={}, ss=dict(q_ss=G["v_q_ss"], t_ss=tss, S33=G["S33"]), co=dict(celloff=m),
                        co_ss=dict(celloff=m, q_ss=G["v_q_ss"], t_ss=tss, S33=G["S33"]),
                        glob=dict(local=False, egq=0.1, egt=0.2))
        for vn, kw in variants.items():
            sc, i2, j2, bt = oracle.viterbi(G["v_q_p"], G["v_q_tr"], tp, ttr, **kw)
            assert bits(sc) == bits(G[f"v_{vn}_{k}_score"][0]), (vn, k)
            assert (i2, j2) == tuple(G[f"v_{vn}_{k}_res"]), (vn, k)
            assert np.array_equal(bt[1:, 1:], G[f"v_{vn}_{k}_bt"][1:, 1:]), (vn, k)


def test_prefilter_goldens(oracle):
    """Query profile (un-striped) and ungapped scores against Prefilter::stripe_query_profile /
    ungapped_sse_score outputs of the reference."""
    G = golden()
    prof = oracle.prefilter_query_profile(G["q_p"], G["q_pav"], G["cs219_lin"], 50, 4)
    assert np.array_equal(prof, G["pf_prof"])
    for k in range(int(G["pf_nseq"][0])):
        assert oracle.ungapped(G["pf_prof"], G[f"pf_seq{k}"], 50) == G[f"pf_ref{k}"][0], k


def test_exclude_alignment_cross(oracle):
    from tests.util import rasterize_exclusion
    m = np.zeros((101, 81), np.uint8)
    i_s = np.array([0, 60, 59, 58, 58, 57], np.int32)
    j_s = np.array([0, 50, 49, 48, 47, 46], np.int32)
    oracle.exclude_alignment(m, i_s, j_s, 5)
    assert np.array_equal(m, rasterize_exclusion(100, 80, i_s, j_s, 5))
    assert m[58, 47] == 1 and m[57, 46] == 0 and m[100, 50] == 1 and m[60, 10] == 1 and m[60, 9] == 0
