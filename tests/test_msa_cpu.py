"""CPU tests of the A3M scanner (hhg_a3m_parse: Alignment::Read + Compress + the first steps of Filter2 restated on the
host side of the library) against the compiled reference: residue codes, insert counts, residue counts and the
length order the identity filter walks."""
import numpy as np
import pytest

from tests import msa_cases


def test_scanner_equals_reference(refshim, tmp_path):
    from hhsuite_b200 import capi
    for k, t in enumerate(msa_cases.texts() + msa_cases.TINY):
        path = tmp_path / f"m{k}.a3m"
        path.write_bytes(t)
        ref = refshim.msa_to_hmm(str(path))
        got = capi.a3m_parse(t)
        assert (got["L"], got["N_in"], got["kfirst"], got["kss_pred"], got["kss_conf"]) == \
               (ref["L"], ref["N_in"], ref["kfirst"], ref["kss_pred"], ref["kss_conf"]), k
        assert np.array_equal(got["X"][:, 1:-1], ref["X"][:, 1:-1]), k      # columns 0 / L+1 are set later by the reference
        rows = ref["keep"] > 0                                              # the shim exports inserts of profile rows
        assert np.array_equal(got["I"][rows][:, :-1], ref["I"][rows][:, :-1]), k
        assert np.array_equal(got["nres"], ref["nres"]) and np.array_equal(got["ksort"], ref["ksort"]), k


def test_scanner_errors_and_limits():
    from hhsuite_b200 import capi
    with pytest.raises(capi.HhgError):
        capi.a3m_parse(b"no header line\nACDE\n")
    with pytest.raises(capi.HhgError):
        capi.a3m_parse(b">a\nACDEFGHIKL\n>b\nACDEFGH\n")                    # unequal number of match columns
    with pytest.raises(capi.HhgError):
        capi.a3m_parse(b">a\n>b\nACDE\n")                                   # a sequence without residues
    with pytest.raises(capi.HhgError):
        capi.a3m_parse(b">a\nACDEFGHIKL\n", capi.MsaParams.defaults(M=4))   # 1 (case), 2 (gap rule), 3 (first sequence)
    # maxseq: rows beyond the limit are ignored like the reference does (with a warning)
    t = b">m\nACDEFGHIKL\n" + b"".join(b">s%d\nACDEFGHIKL\n" % i for i in range(10))
    assert capi.a3m_parse(t, capi.MsaParams.defaults(maxseq=4))["N_in"] == 4
    # lower case = insert, '.' dropped, inserts before the first match column belong to column 0
    a = capi.a3m_parse(b">m\nACDE\n>s\nab.A-cdDE\n")
    assert a["L"] == 4 and a["I"][1].tolist() == [2, 0, 2, 0, 0, 0] and a["X"][1, 1:5].tolist() == [0, 22, 3, 6][:0] + [0, 21, 3, 6]


def test_ca3m_scanner_equals_reference(refshim, tmp_path):
    """Compressed alignments: hhg_ca3m_parse (Alignment::ReadCompressed + Compress restated) against the compiled
    reference reading the same ffindex triple."""
    from hhsuite_b200 import capi, ffindex
    prefix = msa_cases.ca3m_database(tmp_path)
    sq = ffindex.FFIndex(prefix + "_sequence.ffdata")
    ca = ffindex.FFIndex(prefix + "_ca3m.ffdata")
    seqs = capi.SeqDb.make(bytes(sq.data), sq.offsets, sq.lengths)
    for k, name in enumerate(ca.names):
        ref = refshim.ca3m_to_hmm(prefix, name)
        got = capi.ca3m_parse(bytes(ca.record(k)), seqs)
        assert (got["L"], got["N_in"], got["kfirst"]) == (ref["L"], ref["N_in"], ref["kfirst"]) and got["keep"][0] == 0, name
        assert np.array_equal(got["X"][:, 1:-1], ref["X"][:, 1:-1]), name
        rows = ref["keep"] > 0
        assert np.array_equal(got["I"][rows][:, :-1], ref["I"][rows][:, :-1]), name
        assert np.array_equal(got["nres"], ref["nres"]) and np.array_equal(got["ksort"], ref["ksort"]), name
    # a record with the consensus row only has no master sequence (the reference exits)
    from hhsuite_b200 import synth
    lone, _, _ = synth.a3m_to_ca3m(synth.a3m_text(30, 0, 24))
    with pytest.raises(capi.HhgError):
        capi.ca3m_parse(lone + b"\0", seqs)
    # a sequence record pointing outside the sequence database is refused
    bad = bytearray(ca.record(0)); pos = bad.index(b";") + 1; bad[pos:pos + 4] = (10 ** 6).to_bytes(4, "little")
    with pytest.raises(capi.HhgError):
        capi.ca3m_parse(bytes(bad), seqs)
    sq.close(); ca.close()
