"""Synthetic A3M alignments shared by the alignment -> HMM tests (CPU scanner test and GPU parity tests)."""
import os

CASES = [  # (match columns, sequences after the master, seed, generator options)
    (50, 20, 1, {}),
    (8, 5, 2, {}),                                    # fewer than NCOLMIN columns: global weights everywhere
    (120, 300, 3, dict(with_ss=True)),
    (200, 60, 4, dict(with_comment=True)),
    (80, 40, 5, dict(consensus_first=True)),          # compressed databases: consensus row outside the profile
    (30, 0, 6, {}),                                   # single sequence
    (300, 150, 7, dict(ident=0.9, dup_frac=0.6)),     # most rows removed by the 90 % identity filter
    (15, 3, 8, dict(with_ss=True, with_comment=True)),
    (60, 1, 9, {}),
    (431, 58, 10, dict(ident=0.3)),
    (700, 90, 11, dict(ident=0.6, with_ss=True)),
]


def texts():
    from hhsuite_b200 import synth
    out = [synth.a3m_text(L, n, seed, **kw).encode() for (L, n, seed, kw) in CASES]
    q = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "data", "query.a3m")
    if os.path.exists(q):
        out.append(open(q, "rb").read())
    return out
