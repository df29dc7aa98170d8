"""The drop-in check: the reference's own HHalign front half feeding (1) the reference's
ViterbiRunner::alignment and (2) the C-ABI adapter of INTEGRATION.md; every Hit of every alternative
alignment must be identical.  Uses the binary oracle/_ref/hh_dropin_check (built from
oracle/ref_gpu_adapter.cpp against the unmodified reference; shipped prebuilt to the GPU box)."""
import os
import subprocess

import numpy as np

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "hh_dropin_check")
QUERY = os.path.join(ROOT, "oracle", "_ref", "data", "query.hhm")


def _run(args):
    r = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    return r


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_reference_runner_vs_gpu_adapter(tmp_path):
    from hhsuite_b200 import synth
    # config 1: data/query.hhm vs one synthetic HMM (L=150) + itself (a strong hit with alternative alignments)
    files = []
    for k, L in enumerate([150, 60, 431, 300, 33, 200, 97, 120, 250, 75]):
        f = tmp_path / f"t{k}.hhm"
        f.write_text(synth.hhm_text(L, 100 + k, f"t{k}"))
        files.append(str(f))
    r = _run([QUERY, files[0]])
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr
    r = _run([QUERY, QUERY] + files)          # 11 templates: more than one 8-lane batch, self-hit re-queued 4x
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr
    assert "up to irep 4" in r.stdout or "up to irep 3" in r.stdout or "up to irep 2" in r.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_reference_runner_vs_gpu_adapter_with_ss(tmp_path):
    """Query and all templates carry predicted secondary structure -> the reference takes the *AndSS kernels."""
    from hhsuite_b200 import synth
    q = tmp_path / "q.hhm"
    q.write_text(synth.hhm_text(180, 7, "qss", with_ss=True))
    files = []
    for k, L in enumerate([180, 90, 140, 260, 45, 180, 75, 200, 66]):
        f = tmp_path / f"s{k}.hhm"
        f.write_text(synth.hhm_text(L, 7 if k in (0, 5) else 300 + k, f"s{k}", with_ss=True))   # two self-like hits
        files.append(str(f))
    r = _run([str(q)] + files)
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_mac_realignment_in_the_dropin_check(tmp_path):
    """--mac: PosteriorDecoderRunner::executeComputation vs hhg_mac_realign on every Viterbi hit."""
    from hhsuite_b200 import synth
    files = []
    for k, L in enumerate([150, 431, 300, 97, 200]):
        f = tmp_path / f"m{k}.hhm"
        f.write_text(synth.hhm_text(L, 500 + k, f"m{k}"))
        files.append(str(f))
    r = _run(["--mac", QUERY, QUERY] + files)
    assert r.returncode == 0 and "all MAC alignments identical" in r.stdout and "all hits identical" in r.stdout, \
        r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_mac_realignment_with_predicted_secondary_structure(tmp_path):
    """Query and templates carry ss_pred: Viterbi runs the *AndSS kernels (hit.ssm2 = 3), and the reference's MAC calls
    Viterbi::ScoreSS(..., hit.ssm2, ...) whose switch knows HMM::PRED_PRED = 4 but not 3 (src/hhhit.cpp:303-308 vs
    src/hhviterbi.h:199-209, src/hhhmm.h:58-61): the SS term of the realignment is identically 0 for predicted-vs-
    predicted structure, so hhg_mac_realign (no SS term) must still match bit for bit."""
    from hhsuite_b200 import synth
    q = tmp_path / "q.hhm"
    q.write_text(synth.hhm_text(170, 7, "qss", with_ss=True))
    files = []
    for k, L in enumerate([170, 90, 140, 260, 170, 75]):
        f = tmp_path / f"s{k}.hhm"
        f.write_text(synth.hhm_text(L, 7 if k in (0, 4) else 600 + k, f"s{k}", with_ss=True))
        files.append(str(f))
    r = _run(["--mac", str(q)] + files)
    assert r.returncode == 0 and "all MAC alignments identical" in r.stdout and "all hits identical" in r.stdout, \
        r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_behind_the_c_abi(tmp_path):
    """--gpus 2: C++ only (no torch): templates sharded over two GPUs, one host thread + hhg_ctx + hhg_comm each;
    all Hits equal the reference's and the NCCL-merged list (hhg_plan_topk + hhg_plan_topk_paths) equals the
    reference's first-round hits sorted by (Hit.score desc, template index asc)."""
    from hhsuite_b200 import synth
    files = []
    for k, L in enumerate([150, 60, 431, 300, 33, 200, 97, 120, 250, 75, 150]):     # 150 twice: seeds differ
        f = tmp_path / f"g{k}.hhm"
        f.write_text(synth.hhm_text(L, 100 + k, f"g{k}"))
        files.append(str(f))
    r = _run(["--gpus", "2", QUERY, QUERY] + files + [files[3]])                    # files[3] twice: an exact score tie
    assert r.returncode == 0 and "all hits identical" in r.stdout and "identical to the reference's sorted" in r.stdout, \
        r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_ss_mode_is_the_batch_consensus_of_the_reference(tmp_path):
    """ViterbiConsumerThread::align takes the *AndSS kernels only when ALL 8 lanes of a batch carry a predicted
    secondary structure (src/hhviterbirunner.cpp:14-22); batches are formed after the length sort of the chunk.  Mixed
    template list, entries with their real lengths: the adapter has to reproduce the reference's per-batch choice."""
    from hhsuite_b200 import synth
    q = tmp_path / "q.hhm"
    q.write_text(synth.hhm_text(160, 7, "qss", with_ss=True))
    files = []
    lens = [160, 90, 140, 260, 45, 160, 75, 200, 66, 120, 33, 300, 180, 95, 210, 150, 58, 170, 240]
    for k, L in enumerate(lens):
        f = tmp_path / f"x{k}.hhm"
        f.write_text(synth.hhm_text(L, 7 if k in (0, 5) else 400 + k, f"x{k}", with_ss=(k % 3 != 1)))
        files.append(str(f))
    r = _run(["--real-lengths", str(q)] + files)
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
@pytest.mark.parametrize("homologs", [0, 40], ids=["stops", "continues"])
def test_hhblits_early_stopping(tmp_path, homologs):
    """hhblits aligns the prefilter's list in chunks of 2000 and stops after a chunk whose hits sum to less than
    2000 * 0.01 in 1/(1+Eval) (src/hhviterbirunner.cpp:109-111,178-188,213-247).  2100 templates: without homologs
    the first chunk ends the round (the last 100 are never aligned), with 40 copies of the query it goes on."""
    from hhsuite_b200 import synth
    rng = np.random.default_rng(5)
    qtext = open(QUERY).read()
    files = []
    for k in range(2100):
        f = tmp_path / f"e{k}.hhm"
        if k < homologs:
            f.write_text(qtext)
        else:
            f.write_text(synth.hhm_text(int(rng.integers(30, 70)), 9000 + k, f"e{k}"))
        files.append(str(f))
    r = _run(["--hhblits", "2100", "--real-lengths", QUERY] + files)
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    if homologs:      # the first chunk passes the filter; the 100-entry rest is aligned too (and then fails it, at the end)
        assert "early stop after 2100 of 2100" in r.stdout and "reference aligned 2100 first-round hits" in r.stdout
    else:
        assert "early stop after 2000 of 2100" in r.stdout and "reference aligned 2000 first-round hits" in r.stdout


def _write_cs219_ffindex(base, names, seqs):
    """<base>.ffdata / .ffindex like cstranslate writes them: entry = state bytes + NUL, index sorted by name."""
    order = sorted(range(len(names)), key=lambda k: names[k])
    off = 0
    with open(base + ".ffdata", "wb") as fd, open(base + ".ffindex", "w") as fi:
        for k in order:
            b = bytes(seqs[k]) + b"\0"
            fd.write(b)
            fi.write(f"{names[k]}\t{off}\t{len(b)}\n")
            off += len(b)


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
@pytest.mark.parametrize("maxnumdb", [20000, 40])
def test_prefilter_db_seam(tmp_path, refshim, maxnumdb):
    """Seam 2: the reference's Prefilter::prefilter_db (AVX2, OpenMP) vs GpuPrefilter::prefilter_db (same signature,
    C-ABI underneath) on one synthetic cs219 ffindex: new_prefilter_hits / old_prefilter_hits must be equal element by
    element (lengths, names, order), including the previous_hits split and the maxnumdb cap."""
    from tests.util import golden
    G = golden()
    q = refshim.load_query_hhm(QUERY)
    import hhsuite_b200 as hh
    prof = hh.capi.build_prefilter_profile(q["p"], q["pav"], G["cs219_lin"], 50, 4)
    best = prof[:219].argmax(axis=0).astype(np.uint8)
    rng = np.random.default_rng(17)
    n = 3000
    seqs = [rng.integers(0, 219, int(L), dtype=np.uint8) for L in rng.integers(30, 500, n)]
    for k in range(0, n, 12):                      # planted homologs with substitutions and an indel
        a = int(rng.integers(0, 200)); ln = int(rng.integers(60, 230))
        seg = best[a:a + ln].copy()
        seg[rng.random(ln) < 0.2] = rng.integers(0, 219)
        cut = int(rng.integers(10, ln - 10))
        seqs[k] = np.concatenate([seg[:cut], rng.integers(0, 219, int(rng.integers(0, 6)), dtype=np.uint8), seg[cut:]])
    names = [f"T{k:05d}.a3m" for k in range(n)]
    base = str(tmp_path / "db_cs219")
    _write_cs219_ffindex(base, names, seqs)
    prev = ",".join(f"T{k:05d}" for k in range(0, n, 36))
    r = _run(["--prefilter", base, "--maxnumdb", str(maxnumdb), "--previous", prev, QUERY])
    assert r.returncode == 0 and "identical lists" in r.stdout, r.stdout + r.stderr
    assert " 0 new + 0 old" not in r.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_excluded_regions_in_the_dropin_check(tmp_path):
    """par.exclstr / par.template_exclstr: the reference's runner masks the regions itself; the adapter passes the
    same ranges to hhg_set_excluded_regions; all Hits of all alternative alignments identical."""
    from hhsuite_b200 import synth
    files = []
    for k, L in enumerate([150, 431, 300, 97, 200, 60]):
        f = tmp_path / f"x{k}.hhm"
        f.write_text(synth.hhm_text(L, 800 + k, f"x{k}"))
        files.append(str(f))
    r = _run(["--excl", "1-33,200-260", "--template-excl", "10-20,400-500", QUERY, QUERY] + files)
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_alignment_templates_in_the_dropin_check(tmp_path):
    """Templates given as A3M alignments: the reference's runner takes the alignment branch of getTemplateHMM per
    template (Read, Compress, Filter, FrequenciesAndTransitions with global weights, src/hhviterbirunner.cpp:143),
    the adapter builds the shard with hhg_db_create_a3m; every Hit must be identical."""
    from hhsuite_b200 import synth
    qa = os.path.join(ROOT, "oracle", "_ref", "data", "query.a3m")
    files = []
    for k, (L, n) in enumerate([(150, 40), (60, 5), (300, 120), (33, 0), (200, 25), (97, 60), (120, 250), (250, 12), (75, 33)]):
        f = tmp_path / f"t{k}.a3m"
        f.write_text(synth.a3m_text(L, n, 400 + k, f"t{k}", with_ss=(k % 4 == 0)))
        files.append(str(f))
    if os.path.exists(qa):
        files.insert(1, qa)                                  # the query's own alignment: a strong hit, alternative alignments
    r = _run(["--hhm-loader", QUERY] + files)
    assert r.returncode == 0 and "all hits identical" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/hh_dropin_check not built/shipped")
def test_mac_realignment_with_excluded_regions(tmp_path):
    """--mac together with -excl / -template_excl: PosteriorDecoderRunner masks the regions in the realignment as well
    (src/hhposteriordecoder.cpp:100-108); the library applies the context's regions in k_mac_band."""
    from hhsuite_b200 import synth
    files = []
    for k, L in enumerate([150, 431, 300, 97]):
        f = tmp_path / f"y{k}.hhm"
        f.write_text(synth.hhm_text(L, 900 + k, f"y{k}"))
        files.append(str(f))
    r = _run(["--mac", "--excl", "1-33,200-260", "--template-excl", "10-20,400-500", QUERY, QUERY] + files)
    assert r.returncode == 0 and "all MAC alignments identical" in r.stdout and "all hits identical" in r.stdout, \
        r.stdout + r.stderr
