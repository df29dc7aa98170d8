import numpy as np


def test_synth_is_deterministic_and_well_formed():
    from hhsuite_b200 import synth
    a = synth.prepared_db(300, seed=5, fast=True)
    b = synth.prepared_db(300, seed=5, fast=True)
    for k in a:
        assert np.array_equal(a[k], b[k])
    L = a["L"]
    assert L.min() >= 30 and L.max() <= 2000
    assert a["p"].shape[0] == int((L + 2).sum()) and a["tr"].shape[0] == int((L + 1).sum())
    assert np.isfinite(a["p"]).all() and (a["tr"] <= 0).all()
    # topology rows (src/hhviterbialgorithm.cpp:47-55)
    first = a["tr"][a["tr_off"]]
    assert np.all(first[:, 0] == 0) and np.all(first[:, 1] < -90)
    last = a["tr"][a["tr_off"] + L]
    assert np.all(last[:, 2] < -90) and np.all(last[:, 6] < -90)
    big = synth.lengths(200000, np.random.default_rng(1))
    assert 195 <= np.median(big) <= 205


def test_hhm_text_shape():
    from hhsuite_b200 import synth
    t = synth.hhm_text(50, 3, "x")
    assert t.startswith("HHsearch 1.6") and t.rstrip().endswith("//")
    assert sum(1 for ln in t.splitlines() if ln[:2] in [c + " " for c in synth.AA_SORTED]) == 50
