"""Seeded synthetic inputs for the Viterbi / prefilter hot path.

Two kinds of output:
  * HHM text (``hhm_text``) that the reference's own reader accepts
    (/root/reference/src/hhhmm.cpp:202-691) -- used for the small parity configs where the
    prepared fp32 profiles must come out of the reference's PrepareTemplateHMM.
  * "prepared profiles" (``prepared_profile`` / ``prepared_db``): the fp32 arrays the DP consumes,
    in the layout of include/hhg.h:  p[(L+2)*20], tr[(L+1)*7] (reference enum order
    M2M,M2I,M2D,I2M,I2I,D2M,D2D, src/hhdecl.h:68; log2 values).  Both the CPU oracle / reference
    and the GPU read exactly these bits.
"""
from __future__ import annotations

import numpy as np

AA_SORTED = "ACDEFGHIKLMNPQRSTVWY"      # column order in HHM files (src/hhhmm.cpp:73 of query.hhm)
NEG = -99999.0 / 1000.0                  # what the reference stores for '*' (strinta -> -99999)/HMMSCALE

# Background = the NULL line every HH-suite HHM carries (e.g. data/query.hhm:72), in 1/1000 bits, column
# order ACDEFGHIKLMNPQRSTVWY.  Using the standard line matters: HMM::Read overwrites the global pb array
# from each file's NULL line (src/hhhmm.cpp:540-543), so templates with different NULL lines would make the
# reference's results depend on the (thread-racy) order in which templates are read.
_NULL_MILLIBITS = np.array([3706, 5728, 4211, 4064, 4839, 3729, 4763, 4308, 4069, 3323, 5509, 4640, 4464, 4937,
                            4285, 4423, 3815, 3783, 6325, 4665], dtype=np.float64)
_PB = np.exp2(-_NULL_MILLIBITS / 1000.0)
_PB /= _PB.sum()


def lengths(n: int, rng: np.random.Generator, median: int = 200, sigma: float = 0.5,
            lo: int = 30, hi: int = 2000) -> np.ndarray:
    """Target lengths ~ round(lognormal(ln median, sigma)) clipped to [lo, hi]."""
    L = np.rint(np.exp(rng.normal(np.log(median), sigma, size=n))).astype(np.int64)
    return np.clip(L, lo, hi).astype(np.int32)


def _columns(L: int, rng: np.random.Generator) -> np.ndarray:
    """L Dirichlet emission columns with a mix of sharpness."""
    alpha = rng.choice([0.3, 1.0, 3.0], size=L)[:, None] * (_PB[None, :] * 20.0)
    f = rng.gamma(alpha, 1.0) + 1e-12
    return f / f.sum(axis=1, keepdims=True)


def _transitions(L: int, rng: np.random.Generator) -> np.ndarray:
    """tr[i][7] linear probabilities for i=0..L (reference order M2M,M2I,M2D,I2M,I2I,D2M,D2D)."""
    tr = np.zeros((L + 1, 7))
    m2i = rng.uniform(0.005, 0.05, L + 1)
    m2d = rng.uniform(0.005, 0.05, L + 1)
    tr[:, 0] = 1 - m2i - m2d
    tr[:, 1] = m2i
    tr[:, 2] = m2d
    i2i = rng.uniform(0.3, 0.7, L + 1)
    tr[:, 3] = 1 - i2i
    tr[:, 4] = i2i
    d2d = rng.uniform(0.3, 0.7, L + 1)
    tr[:, 5] = 1 - d2d
    tr[:, 6] = d2d
    # topology of the reference HMM (hhviterbialgorithm.cpp:47-55): start state only M->M; column L
    # has no transitions into delete states
    tr[0] = [1, 0, 0, 1, 0, 1, 0]
    tr[L, 0] = 1 - tr[L, 1]
    tr[L, 2] = 0
    tr[L, 5] = 1
    tr[L, 6] = 0
    return tr


def hhm_text(L: int, seed: int, name: str = "synth", with_ss: bool = False) -> str:
    """A syntactically complete HHM 1.6 record of length L."""
    rng = np.random.default_rng(seed)
    f = _columns(L, rng)
    tr = _transitions(L, rng)
    neff = rng.uniform(1.0, 10.0, L + 1)
    cons = "".join(AA_SORTED[i] for i in f.argmax(axis=1))

    def enc(p):
        return "*" if p <= 0 or -1000.0 * np.log2(p) > 99998 else str(int(round(-1000.0 * np.log2(p))))

    out = ["HHsearch 1.6", f"NAME  {name}", "FAM   ", f"FILE  {name}", "COM   synthetic",
           "DATE  Tue Sep 22 00:00:00 2026", f"LENG  {L} match states, {L} columns in multiple alignment",
           "FILT  1 out of 1 sequences passed filter", f"NEFF  {neff.mean():.1f} ", "SEQ"]
    if with_ss:
        ss = rng.choice(list("HEC"), size=L)
        out += [">ss_pred", "".join(ss), ">ss_conf", "".join(str(d) for d in rng.integers(0, 10, L))]
    out += [">Consensus", cons.lower(), f">{name}", cons, "#"]
    out.append("NULL   " + "\t".join(str(int(v)) for v in _NULL_MILLIBITS) + "\t")
    out.append("HMM    " + "\t".join(AA_SORTED) + "\t")
    out.append("       M->M\tM->I\tM->D\tI->M\tI->I\tD->M\tD->D\tNeff\tNeff_I\tNeff_D")
    out.append("       " + "\t".join(enc(p) for p in tr[0]) + "\t*\t*\t*\t")
    for i in range(1, L + 1):
        out.append(f"{cons[i-1]} {i:<4d} " + "\t".join(enc(p) for p in f[i - 1]) + f"\t{i}")
        nm = int(round(1000 * neff[i]))
        out.append("       " + "\t".join(enc(p) for p in tr[i]) + f"\t{nm}\t{nm // 3}\t{nm // 4}\t")
        out.append("")
    out.append("//")
    return "\n".join(out) + "\n"


def _log2_tr(tr: np.ndarray) -> np.ndarray:
    with np.errstate(divide="ignore"):
        lg = np.log2(tr)
    lg[~np.isfinite(lg)] = NEG
    return lg.astype(np.float32)


def prepared_profile(L: int, rng: np.random.Generator, base: np.ndarray | None = None,
                     noise: float = 0.0):
    """(p[(L+2),20] f32, tr[(L+1),7] f32, ss[(L+2)] u8).  p is a probability ratio like the output
    of IncludeNullModelInHMM (src/hhhmm.cpp:2074-2081); rows 0 and L+1 hold the background."""
    f = _columns(L, rng)
    if base is not None:                                   # planted homolog: noisy copy of `base`
        n = min(L, base.shape[0])
        off = int(rng.integers(0, base.shape[0] - n + 1))
        f[:n] = (1 - noise) * base[off:off + n] + noise * f[:n]
    mix = 0.85 * f + 0.15 * _PB[None, :]
    p = np.empty((L + 2, 20), dtype=np.float32)
    p[1:L + 1] = (mix / _PB[None, :]).astype(np.float32)
    p[0] = p[L + 1] = 1.0
    tr = _log2_tr(_transitions(L, rng))
    ss = np.zeros(L + 2, dtype=np.uint8)
    ss[1:L + 1] = (rng.integers(1, 4, L) * 11 + rng.integers(1, 11, L)).astype(np.uint8)
    return p, tr, ss


def query_profile(L: int, seed: int):
    """Prepared query (linear probabilities, no null-model division -- src/hhfunc.cpp:121-160)."""
    rng = np.random.default_rng(seed)
    f = _columns(L, rng)
    mix = 0.9 * f + 0.1 * _PB[None, :]
    p = np.empty((L + 2, 20), dtype=np.float32)
    p[1:L + 1] = mix.astype(np.float32)
    p[0] = p[L + 1] = _PB.astype(np.float32)
    tr = _log2_tr(_transitions(L, rng))
    ss = np.zeros(L + 2, dtype=np.uint8)
    ss[1:L + 1] = (rng.integers(1, 4, L) * 11 + rng.integers(1, 11, L)).astype(np.uint8)
    pav = p[1:L + 1].mean(axis=0).astype(np.float32)
    return p, tr, ss, pav, mix


def prepared_db(n: int, seed: int, median: int = 200, sigma: float = 0.5, lo: int = 30,
                hi: int = 2000, query_cols: np.ndarray | None = None, planted: int = 0,
                lens: np.ndarray | None = None, fast: bool = False):
    """Concatenated prepared profiles of n targets.

    Returns dict(L int32[n], p f32[sum(L+2),20], tr f32[sum(L+1),7], ss u8[sum(L+2)],
                 p_off int64[n] (row offsets into p/ss), tr_off int64[n] (row offsets into tr)).
    ``fast`` draws all columns in one vectorised call (for 100k..1M targets)."""
    rng = np.random.default_rng(seed)
    L = lengths(n, rng, median, sigma, lo, hi) if lens is None else np.asarray(lens, dtype=np.int32)
    p_rows = (L.astype(np.int64) + 2)
    t_rows = (L.astype(np.int64) + 1)
    p_off = np.concatenate([[0], np.cumsum(p_rows)[:-1]]).astype(np.int64)
    tr_off = np.concatenate([[0], np.cumsum(t_rows)[:-1]]).astype(np.int64)
    P = np.empty((int(p_rows.sum()), 20), dtype=np.float32)
    T = np.empty((int(t_rows.sum()), 7), dtype=np.float32)
    S = np.zeros(int(p_rows.sum()), dtype=np.uint8)
    if fast:
        # columns and transition rows are drawn (with replacement) from pools of 65536 distinct
        # Dirichlet columns / transition rows: i.i.d. columns as above at a fraction of the sampling cost
        tot = int(p_rows.sum())
        NP = 65536
        pool_f = _columns(NP, rng).astype(np.float32)
        pool_p = ((0.85 * pool_f + 0.15 * _PB[None, :].astype(np.float32)) /
                  _PB[None, :].astype(np.float32)).astype(np.float32)
        idx = rng.integers(0, NP, tot, dtype=np.int32)
        np.take(pool_p, idx, axis=0, out=P)
        S[:] = (rng.integers(1, 4, tot, dtype=np.uint8) * 11 + rng.integers(1, 11, tot, dtype=np.uint8))
        tt = int(t_rows.sum())
        pool_t = _log2_tr(_transitions(NP - 1, rng))
        idx = rng.integers(1, NP - 1, tt, dtype=np.int32)
        np.take(pool_t, idx, axis=0, out=T)
        first = tr_off
        last = tr_off + L
        T[first] = _log2_tr(np.array([[1, 0, 0, 1, 0, 1, 0]], dtype=np.float64))[0]
        lin = np.exp2(T[last].astype(np.float64))
        lin[:, 0] = 1 - lin[:, 1]; lin[:, 2] = 0; lin[:, 5] = 1; lin[:, 6] = 0
        T[last] = _log2_tr(lin)
        P[p_off] = 1.0
        P[p_off + L + 1] = 1.0
        S[p_off] = 0
        S[p_off + L + 1] = 0
        if planted and query_cols is not None:
            for k in range(min(planted, n)):
                p, t, s = prepared_profile(int(L[k]), rng, query_cols, noise=0.3 + 0.05 * (k % 8))
                P[p_off[k]:p_off[k] + L[k] + 2] = p
    else:
        for k in range(n):
            base = query_cols if (query_cols is not None and k < planted) else None
            p, t, s = prepared_profile(int(L[k]), rng, base, noise=0.3 + 0.05 * (k % 8))
            P[p_off[k]:p_off[k] + L[k] + 2] = p
            T[tr_off[k]:tr_off[k] + L[k] + 1] = t
            S[p_off[k]:p_off[k] + L[k] + 2] = s
    return dict(L=L, p=P, tr=T, ss=S, p_off=p_off, tr_off=tr_off)


def cs219_db(n: int, seed: int, lens: np.ndarray | None = None, median: int = 200):
    """Synthetic column-state sequences: bytes in [0,218] (throughput is data-independent)."""
    rng = np.random.default_rng(seed)
    L = lengths(n, rng, median) if lens is None else np.asarray(lens, dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(L.astype(np.int64))[:-1]]).astype(np.int64)
    seq = rng.integers(0, 219, int(L.sum()), dtype=np.uint8)
    return dict(L=L, seq=seq, off=off)


AA = "ARNDCQEGHILKMFPSTWYV"


def a3m_text(L: int, nseq: int, seed: int, name: str = "msa", with_ss: bool = False, with_comment: bool = False,
             consensus_first: bool = False, ident: float = 0.5, dup_frac: float = 0.3, x_frac: float = 0.01,
             ss_conf: bool = True) -> str:
    """A synthetic A3M alignment with L match columns (upper case / '-') and `nseq` sequences after the master:
    point mutations at rate 1-ident, a fraction of near-duplicates (> 90 % identical: removed by the filter), runs
    of deletions, lower-case insert runs, N-/C-terminal truncation (end gaps), a few 'X', optional >ss_pred / >ss_conf
    rows, a '#' name line and a '_consensus' first sequence (what compressed databases produce)."""
    rng = np.random.default_rng(seed)
    master = rng.integers(0, 20, L)
    out = []
    if with_comment:
        out.append(f"#{name} synthetic alignment")
    if with_ss:
        out.append(">ss_pred PSIPRED predicted secondary structure")
        out.append("".join("CHE"[int(v)] for v in rng.integers(0, 3, L)))
        conf = "".join(str(int(v)) for v in rng.integers(0, 10, L))
        if ss_conf:
            out.append(">ss_conf PSIPRED confidence values")
            out.append(conf)
    mseq = "".join(AA[a] for a in master)
    out.append(f">{name}_consensus" if consensus_first else f">{name} master")
    out.append(mseq)
    prev = master
    for k in range(nseq):
        base = prev if (k > 0 and rng.random() < dup_frac) else master
        rate = 0.03 if base is prev else (1.0 - ident) * rng.uniform(0.3, 1.4)
        seq = base.copy()
        mut = rng.random(L) < rate
        seq[mut] = rng.integers(0, 20, int(mut.sum()))
        prev = seq
        chars = [AA[a] for a in seq]
        for i in np.nonzero(rng.random(L) < x_frac)[0]:
            chars[i] = "X"
        # deletions
        for _ in range(int(rng.integers(0, 4))):
            a = int(rng.integers(0, L)); b = min(L, a + int(rng.integers(1, 8)))
            for i in range(a, b):
                chars[i] = "-"
        # end gaps
        if rng.random() < 0.5:
            a = int(rng.integers(0, max(1, L // 3)))
            for i in range(a):
                chars[i] = "-"
        if rng.random() < 0.5:
            b = int(rng.integers(0, max(1, L // 3)))
            for i in range(L - b, L):
                chars[i] = "-"
        if all(c == "-" for c in chars):
            chars[L // 2] = "A"
        # inserts (lower case) between match columns
        pieces = []
        for i, c in enumerate(chars):
            pieces.append(c)
            if rng.random() < 0.03:
                pieces.append("".join(AA[a].lower() for a in rng.integers(0, 20, int(rng.integers(1, 6)))))
        if rng.random() < 0.2:
            pieces.insert(0, "".join(AA[a].lower() for a in rng.integers(0, 20, int(rng.integers(1, 4)))))
        out.append(f">seq{k} synthetic")
        s = "".join(pieces)
        # a3m files wrap nothing, but the reader joins lines: split some sequences over two lines
        if rng.random() < 0.3 and len(s) > 10:
            cut = int(rng.integers(1, len(s) - 1))
            out.append(s[:cut]); out.append(s[cut:])
        else:
            out.append(s)
    return "\n".join(out) + "\n"


def a3m_to_ca3m(a3m: str, seq_index_base: int = 0):
    """Compress an A3M alignment (master first, no ss rows) the way a compressed HH-suite database stores it
    (what Alignment::ReadCompressed, src/hhalignment.cpp:546-815, decodes): returns (ca3m record bytes,
    [unaligned sequences for the sequence database], [headers for the header database]).  Sequence k (k >= 1) of the
    alignment refers to entry seq_index_base + k - 1 of the sequence database."""
    import struct
    names, rows = [], []
    for ln in a3m.splitlines():
        if ln.startswith("#"):
            continue
        if ln.startswith(">"):
            names.append(ln[1:]); rows.append("")
        elif names:
            rows[-1] += ln.strip()
    cons = "".join(c for c in rows[0] if not c.islower())
    out = bytearray()
    out += (">" + names[0].split()[0] + "_consensus\n").encode()
    out += (cons + "\n;").encode()
    seqs, heads = [], []
    for k in range(1, len(rows)):
        row = rows[k]
        full = "".join(c.upper() for c in row if c not in "-.")
        seqs.append(full.encode() + b"\n")
        heads.append((">" + names[k] + "\n").encode())
        blocks = []
        i, n = 0, len(row)
        m = 0
        while i < n:
            c = row[i]
            if c == "-":
                j = i
                while j < n and row[j] == "-":
                    j += 1
                g = j - i
                if j == n:                       # trailing gaps are implied by the consensus length
                    i = j
                    break
                while g > 0:
                    t = min(g, 127)
                    blocks.append((m, -t)); m = 0; g -= t
                i = j
            elif c.islower():
                j = i
                while j < n and row[j].islower():
                    j += 1
                g = j - i
                while g > 0:
                    t = min(g, 127)
                    blocks.append((m, t)); m = 0; g -= t
                i = j
            else:
                m += 1
                if m == 255:
                    blocks.append((255, 0)); m = 0
                i += 1
        if m:
            blocks.append((m, 0))
        out += struct.pack("<IHH", seq_index_base + k - 1, 1, len(blocks))
        for (nm, nid) in blocks:
            out += struct.pack("<Bb", nm, nid)
    return bytes(out), seqs, heads
