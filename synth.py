"""Seeded synthetic amplicon generator for bench.py and the size-independent GPU tests.

Shapes follow SURVEY.md section 8(d): reads are built as
    [5' flank][SSU/5.8S-end motif sampled from a LEFT-boundary profile][random ITS spacer]
    [5.8S/LSU-start motif sampled from a RIGHT-boundary profile][3' flank]
so that the filter cascade sees realistic boundary motifs (random reads would make the search look
2-3x cheaper than it is).  Motifs are sampled position by position from the match-emission
distributions of profiles of the bundled ITSx_db (a tiny HMMER3 text reader lives here so that the
generator does not depend on the product or on the oracle).  Abundances follow a Zipf law with every
unique seen at least once; read order is a seeded permutation.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
HMM_DIR = os.path.join(ROOT, "itsxpress_b200", "ITSx_db", "HMMs")
SHARPEN = 1.0   # exponent on the match-emission distributions when sampling motifs (1 = as the profile says)


def read_match_emissions(path, prefixes):
    """[(name, probs[M,4])] for profiles whose NAME starts with one of `prefixes`, file order."""
    out = []
    name, M, rows = None, 0, None
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("NAME "):
            name = ln[5:].strip()
        elif ln.startswith("LENG "):
            M = int(ln[5:])
        elif ln.startswith("HMM "):
            i += 2
            if "COMPO" in lines[i]:
                i += 1
            i += 2          # node-0 insert emissions + transitions
            rows = np.zeros((M, 4))
            for k in range(M):
                tok = lines[i].split()
                rows[k] = [0.0 if t == "*" else np.exp(-float(t)) for t in tok[1:5]]
                i += 3
            if any(name.startswith(p) for p in prefixes):
                out.append((name, rows / rows.sum(1, keepdims=True)))
            continue
        i += 1
    return out


def _sample_motifs(rng, profs, n):
    """n motifs, each sampled from a random profile of `profs` -> list of uint8 arrays (ASCII)."""
    pick = rng.integers(len(profs), size=n)
    out = [None] * n
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for p in range(len(profs)):
        idx = np.flatnonzero(pick == p)
        if len(idx) == 0:
            continue
        pr = profs[p][1]
        # sharpen towards the consensus so that most motifs score like real boundaries
        pr = pr ** SHARPEN
        pr = pr / pr.sum(1, keepdims=True)
        cum = np.cumsum(pr, axis=1)
        u = rng.random((len(idx), pr.shape[0], 1))
        codes = (u > cum[None, :, :]).sum(2).clip(0, 3)
        letters = acgt[codes]
        for j, i in enumerate(idx):
            out[i] = letters[j]
    return out


def make_uniques(rng, n_unique, length, left_profs, right_profs, spacer=(110, 150), flank5=(5, 25),
                 n_frac=0.005):
    """n_unique distinct amplicons.  length: int (fixed, truncated/padded) or (lo, hi) (natural length)."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    left = _sample_motifs(rng, left_profs, n_unique)
    right = _sample_motifs(rng, right_profs, n_unique)
    f5 = rng.integers(flank5[0], flank5[1] + 1, n_unique)
    sp = rng.integers(spacer[0], spacer[1] + 1, n_unique)
    seqs = []
    seen = set()
    for i in range(n_unique):
        while True:
            if isinstance(length, tuple):
                tot = int(rng.integers(length[0], length[1] + 1))
            else:
                tot = int(length)
            core = np.concatenate([acgt[rng.integers(4, size=f5[i])], left[i], acgt[rng.integers(4, size=sp[i])],
                                   right[i]])
            if len(core) < tot:
                core = np.concatenate([core, acgt[rng.integers(4, size=tot - len(core))]])
            s = core[:tot].copy()
            if rng.random() < n_frac:
                s[int(rng.integers(tot))] = ord("N")
            b = s.tobytes()
            if b not in seen:
                seen.add(b)
                seqs.append(s)
                break
            sp[i] = rng.integers(spacer[0], spacer[1] + 1)
    return seqs


def zipf_counts(rng, n_reads, n_unique, s=1.0):
    w = 1.0 / np.arange(1, n_unique + 1) ** s
    w /= w.sum()
    extra = rng.multinomial(n_reads - n_unique, w) if n_reads > n_unique else np.zeros(n_unique, np.int64)
    return 1 + extra


def make_reads(seed, n_reads, n_unique, length, hmm_file, left_prefix, right_prefix, zipf_s=1.0,
               spacer=(110, 150), with_qual=False):
    """Returns (seq uint8[total], off int64[n_reads+1], qual or None, unique_of_read int32[n_reads])."""
    rng = np.random.default_rng(seed)
    path = os.path.join(HMM_DIR, hmm_file)
    lp = read_match_emissions(path, [left_prefix])
    rp = read_match_emissions(path, [right_prefix])
    uniq = make_uniques(rng, n_unique, length, lp, rp, spacer=spacer)
    counts = zipf_counts(rng, n_reads, n_unique, zipf_s)
    which = np.repeat(np.arange(n_unique, dtype=np.int32), counts)
    rng.shuffle(which)
    ulen = np.array([len(u) for u in uniq], np.int64)
    uoff = np.zeros(n_unique + 1, np.int64)
    uoff[1:] = np.cumsum(ulen)
    ucat = np.concatenate(uniq)
    lens = ulen[which]
    off = np.zeros(n_reads + 1, np.int64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    # gather: per-read source start - destination start, repeated
    delta = np.repeat(uoff[which] - off[:-1], lens)
    seq = ucat[delta + np.arange(total, dtype=np.int64)]
    qual = None
    if with_qual:
        # Illumina-like: Q in [2, 41] with 3' decay
        pos = np.arange(total, dtype=np.int64) - np.repeat(off[:-1], lens)
        q = 38.0 - 20.0 * (pos / np.repeat(lens, lens)) ** 2 + rng.normal(0, 3, total)
        qual = (np.clip(q, 2, 41).astype(np.uint8) + 33)
    return seq, off, qual, which


def make_quals(seed, off, lo=2, hi=41):
    """Illumina-like qualities for the reads of `off`: Q in [lo, hi] with 3' decay plus +-4 of noise, built with
    8-bit arithmetic so that a 1 M x 250 bp sample costs ~1 s and ~0.5 GB (make_reads(with_qual=True) draws a normal
    per base in float64).  Returns uint8 ASCII (Q + 33)."""
    off = np.asarray(off, np.int64)
    lens = (off[1:] - off[:-1])
    total = int(off[-1])
    rng = np.random.default_rng(seed)
    curve = np.clip(38.0 - 20.0 * (np.arange(256) / 256.0) ** 2, lo, hi).astype(np.int16)      # by 256ths of the read
    if len(lens) and lens.min() == lens.max():
        L = int(lens[0])
        base = np.tile(curve[(np.arange(L) * 256) // max(L, 1)], len(lens))
    else:
        pos = np.arange(total, dtype=np.int64) - np.repeat(off[:-1], lens)
        base = curve[(pos * 256) // np.maximum(np.repeat(lens, lens), 1)]
    q = base + rng.integers(-4, 5, total, dtype=np.int8)
    return (np.clip(q, lo, hi) + 33).astype(np.uint8)


CONFIGS = {
    # BASELINE.json configs[1]: 1 M single-end 250 bp fungal ITS1 reads, 30 % unique.  F.hmm (Fungi) is
    # missing from the reference mount, so the largest present analogue M.hmm (Metazoa) stands in.
    "c2": dict(n_reads=1_000_000, n_unique=300_000, length=250, hmm_file="M.hmm", left_prefix="1_",
               right_prefix="2_", zipf_s=1.0, region="ITS1", taxa="Metazoa (stand-in for Fungi: F.hmm missing)"),
    # scaled-down BASELINE configs[3] shape: merged reads of 330-441 bp, 90 % unique, ITS2 (hmmsearch-bound)
    "c4s": dict(n_reads=400_000, n_unique=360_000, length=(330, 441), hmm_file="M.hmm", left_prefix="3_",
                right_prefix="4_", zipf_s=1.0, spacer=(150, 230), region="ITS2",
                taxa="Metazoa (stand-in for Fungi: F.hmm missing)"),
    # scaled-down BASELINE configs[2] shape: merged reads ~450 bp, 40 % unique, --region ALL --taxa All
    # (motifs sampled from M.hmm, searched against the 1_/4_ profiles of EVERY present taxon file)
    "c3s": dict(n_reads=200_000, n_unique=80_000, length=(380, 520), hmm_file="M.hmm", left_prefix="1_",
                right_prefix="4_", zipf_s=1.0, spacer=(250, 330), region="ALL", taxa="All",
                search_files="ALL"),
    # reduced copy of the same shape for smoke tests
    "c2_small": dict(n_reads=20_000, n_unique=6_000, length=250, hmm_file="M.hmm", left_prefix="1_",
                     right_prefix="2_", zipf_s=1.0, region="ITS1", taxa="Metazoa"),
}


def make_config(name, seed=None, scale=1.0):
    cfg = dict(CONFIGS[name])
    meta = {k: cfg.pop(k) for k in ("region", "taxa")}
    search_files = cfg.pop("search_files", None)
    if scale != 1.0:
        cfg["n_reads"] = max(1000, int(cfg["n_reads"] * scale))
        cfg["n_unique"] = max(300, int(cfg["n_unique"] * scale))
    seed = 2 * 1_000_003 if seed is None else seed
    seq, off, qual, which = make_reads(seed, **cfg)
    if search_files == "ALL":
        files = sorted(f for f in os.listdir(HMM_DIR) if f.endswith(".hmm") and os.path.getsize(os.path.join(HMM_DIR, f)))
    else:
        files = [cfg["hmm_file"]]
    return seq, off, which, dict(cfg, search_files=files, **meta)


def make_pairs(seed, frag_seq, frag_off, read_len=250, trim=(0, 0), err_scale=1.0, n_rate=0.002, lower_rate=0.0,
               empty_rate=0.0):
    """Illumina-like read pairs off the given fragments (one pair per fragment, BASELINE configs[0]/[4] shape):
    R1 = the first read_len bases, R2 = reverse complement of the last read_len bases (a fragment shorter than the
    read gives a staggered pair with random read-through).  Qualities decay towards the 3' end; substitution errors
    are drawn at each base's own error probability (x err_scale); n_rate of the bases become N with Q2; `trim` =
    (lo, hi) random 3' shortening; lower_rate of the reads are lower-cased; empty_rate of the reads are empty.
    Returns (fseq, fqual, foff, rseq, rqual, roff), uint8 / int64 arrays."""
    rng = np.random.default_rng(seed)
    n = len(frag_off) - 1
    comp = np.zeros(256, np.uint8)
    comp[:] = ord("N")
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = []
    for side in (0, 1):
        flen = (frag_off[1:] - frag_off[:-1]).astype(np.int64)
        rl = np.full(n, read_len, np.int64)
        if trim[1] > 0:
            rl -= rng.integers(trim[0], trim[1] + 1, n)
        if empty_rate > 0:
            rl[rng.random(n) < empty_rate] = 0
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum(rl)
        total = int(off[-1])
        pos = np.arange(total, dtype=np.int64) - np.repeat(off[:-1], rl)        # position within the read
        fl = np.repeat(flen, rl)
        inside = pos < fl
        start = np.repeat(frag_off[:-1], rl)
        if side == 0:
            src = start + np.minimum(pos, fl - 1)
            seq = frag_seq[np.maximum(src, 0)]
        else:
            src = start + np.maximum(fl - 1 - pos, 0)
            seq = comp[frag_seq[np.maximum(src, 0)]]
        seq = np.where(inside, seq, acgt[rng.integers(0, 4, total)])              # read-through past the fragment
        q = 38.0 - 22.0 * (pos / max(read_len, 1)) ** 2 + rng.normal(0, 3, total)
        q = np.clip(q, 2, 41).astype(np.int64)
        qv = np.arange(64, dtype=np.int64)
        perr = (np.where(qv < 2, 0.75, 10.0 ** (-qv / 10.0)) * err_scale)[q]       # per-quality table, then gather
        hit = rng.random(total) < perr
        seq = np.where(hit, acgt[(np.searchsorted(acgt, np.minimum(seq, ord("T"))) + rng.integers(1, 4, total)) % 4], seq)
        isn = rng.random(total) < n_rate
        seq = np.where(isn, ord("N"), seq).astype(np.uint8)
        q = np.where(isn, 2, q)
        if lower_rate > 0:
            low = np.repeat(rng.random(n) < lower_rate, rl)
            seq = np.where(low, seq | 0x20, seq).astype(np.uint8)
        out += [seq, (q + 33).astype(np.uint8), off]
    return tuple(out)


def make_pair_config(seed, n_pairs, frag_len=(260, 480), read_len=250, **kw):
    """Random-sequence fragments of the given length range and their read pairs (merge-stage workloads)."""
    rng = np.random.default_rng(seed)
    flen = rng.integers(frag_len[0], frag_len[1] + 1, n_pairs).astype(np.int64)
    foff = np.zeros(n_pairs + 1, np.int64)
    foff[1:] = np.cumsum(flen)
    fseq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(foff[-1]))]
    return (fseq, foff) + make_pairs(seed + 1, fseq, foff, read_len=read_len, **kw)
