"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the CPU oracle
on the same inputs (bit-exact for integer/index work; scores within 0.01 bit; envelopes identical)."""
import os

import numpy as np
import pytest

from conftest import HMM_DIR, TD

pytestmark = pytest.mark.gpu


def _rand_reads(rng, n, lo, hi, dup_frac=0.5, rc_frac=0.2, n_frac=0.02, lower_frac=0.05):
    comp = bytes.maketrans(b"ACGTNRYacgtnry", b"TGCANYRtgcanyr")
    base = []
    reads = []
    for i in range(n):
        if base and rng.random() < dup_frac:
            s = base[rng.integers(len(base))]
            if rng.random() < rc_frac:
                s = s.translate(comp)[::-1]
            if rng.random() < lower_frac:
                s = s.lower()
        else:
            L = int(rng.integers(lo, hi + 1))
            s = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L).tobytes())
            if rng.random() < n_frac:
                p = int(rng.integers(L))
                s = s[:p] + bytes([rng.choice(np.frombuffer(b"NRY", np.uint8))]) + s[p + 1:]
            base.append(s)
        reads.append(s)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), np.uint8).copy(), off


def test_derep_fixture(gpu_ctx, oracle, fixture_reads):
    b, seq, off, _ = fixture_reads
    rep, strand, nu = gpu_ctx.derep(seq, off)
    orep, ostrand, onu = oracle.derep(seq, off)
    assert nu == onu == 137
    assert np.array_equal(rep, orep)
    assert np.array_equal(strand, ostrand)
    first, ab = gpu_ctx.derep_clusters(nu)
    assert np.array_equal(first, np.flatnonzero(orep == np.arange(b.n)))
    assert np.array_equal(ab, np.bincount(orep, minlength=b.n)[first])


@pytest.mark.parametrize("seed,n,lo,hi", [(1, 5000, 1, 40), (2, 20000, 200, 300), (3, 3000, 15, 17)])
def test_derep_random_both_strands(gpu_ctx, oracle, seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    seq, off = _rand_reads(rng, n, lo, hi)
    rep, strand, nu = gpu_ctx.derep(seq, off)
    orep, ostrand, onu = oracle.derep(seq, off)
    assert nu == onu
    assert np.array_equal(rep, orep)
    assert np.array_equal(strand, ostrand)
    assert strand.sum() > 0


def test_derep_forced_collisions(gpu_ctx, oracle):
    rng = np.random.default_rng(7)
    seq, off = _rand_reads(rng, 1500, 30, 60)
    gpu_ctx.set_key_bits(6)
    try:
        rep, strand, nu = gpu_ctx.derep(seq, off)
        st = gpu_ctx.derep_stats()
    finally:
        gpu_ctx.set_key_bits(64)
    orep, ostrand, onu = oracle.derep(seq, off)
    assert st.n_collided > 0
    assert nu == onu and np.array_equal(rep, orep) and np.array_equal(strand, ostrand)


def test_derep_empty(gpu_ctx):
    rep, strand, nu = gpu_ctx.derep(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert nu == 0 and len(rep) == 0


def _uniques(seq, off, rep):
    idx = np.flatnonzero(rep == np.arange(len(rep)))
    parts = [seq[off[i]:off[i + 1]] for i in idx]
    uoff = np.zeros(len(idx) + 1, np.int64)
    uoff[1:] = np.cumsum([len(p) for p in parts])
    return np.concatenate(parts), uoff, idx


def _search_both(gpu_ctx, oracle, files, prefixes, seq, off, left, right, resolve_multidomain=1):
    from itsxpress_b200 import _lib
    paths = [os.path.join(HMM_DIR, f) for f in files]
    n = gpu_ctx.load_profiles(paths, prefixes)
    side = gpu_ctx.set_sides_by_prefix(left, right)
    db = oracle.ProfileDB(paths, prefixes)
    assert db.n == n and db.names == gpu_ctx.names
    prm = _lib.default_params()
    prm.resolve_multidomain = resolve_multidomain
    gpu_ctx.search_seqs(seq, off, prm)
    rows = gpu_ctx.hits()
    st = gpu_ctx.search_stats()
    codes = oracle.digitize(seq.tobytes())
    orows, onrep, ost = db.search(codes, off, oracle.default_params(0, resolve_multidomain))
    return rows, st, orows, onrep, ost, side, db


def _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, seqlen, tie_budget=0):
    assert st.n_pairs == ost.npairs_total
    assert st.n_past_msv == ost.n_past_msv            # integer filter: exact
    assert abs(st.n_past_bias - ost.n_past_bias) <= tie_budget
    assert abs(st.n_past_fwd - ost.n_past_fwd) <= tie_budget
    assert abs(st.n_hits_reported - ost.n_reported_pairs) <= tie_budget
    if tie_budget == 0:
        assert np.array_equal(gpu_ctx.nreported(), onrep)
        assert len(rows) == len(orows)
        # same rows in the same (hmmsearch) order
        assert np.array_equal(rows["seq"], orows["seq"])
        assert np.array_equal(rows["prof"], orows["prof"])
        assert np.array_equal(rows["ienv"], orows["ienv"])       # envelope coordinates identical
        assert np.array_equal(rows["jenv"], orows["jenv"])
        if len(rows):
            assert np.max(np.abs(rows["bitscore"] - orows["bitscore"])) <= 0.01     # bits
        assert np.allclose(np.exp(rows["lnP"]), np.exp(orows["lnP"]), rtol=1e-4, atol=0)   # relative 1e-4 on E
    pos = gpu_ctx.positions(len(seqlen))
    opos = oracle.itspos(orows, side, seqlen)
    for k in ("start", "stop", "tlen", "left_from", "left_to", "right_from", "right_to"):
        assert np.array_equal(pos[k], opos[k]), k
    for k in ("left_score10", "right_score10"):
        assert np.max(np.abs(pos[k].astype(np.int64) - opos[k].astype(np.int64))) <= 1, k
    return pos


def test_search_fixture_metazoa_its2(gpu_ctx, oracle, fixture_reads):
    """137 fixture representatives x 256 Metazoa ITS2 profiles (F.hmm is missing from the mount)."""
    b, seq, off, _ = fixture_reads
    rep, _, _ = oracle.derep(seq, off)
    useq, uoff, _ = _uniques(seq, off, rep)
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, ["M.hmm"], ["3_", "4_"], useq, uoff,
                                                        "3_", "4_")
    assert len(orows) > 1000
    assert st.n_multidomain_regions == ost.n_multidomain_regions
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(uoff).astype(np.int32))


def test_search_fixture_all_taxa_with_iupac(gpu_ctx, oracle, fixture_reads):
    """first 40 representatives (+ the reads with N, + IUPAC / lower-case variants) x every present taxon's
    1_/3_/4_ profiles (G.hmm's 1_ set has the M=25 and M=11 models)."""
    b, seq, off, _ = fixture_reads
    rep, _, _ = oracle.derep(seq, off)
    useq, uoff, _ = _uniques(seq, off, rep)
    nfix = 40
    has_n = [i for i in range(len(uoff) - 1) if b"N" in useq[uoff[i]:uoff[i + 1]].tobytes()]
    keep = sorted(set(range(nfix)) | set(has_n))
    parts = [useq[uoff[i]:uoff[i + 1]] for i in keep]
    # add degenerate + lower-case variants
    v = parts[0].copy(); v[100] = ord("R"); v[150] = ord("n"); parts.append(v)
    parts.append(np.frombuffer(parts[1].tobytes().lower(), np.uint8))
    s2 = np.concatenate(parts)
    o2 = np.zeros(len(parts) + 1, np.int64)
    o2[1:] = np.cumsum([len(p) for p in parts])
    files = sorted(f for f in os.listdir(HMM_DIR) if f.endswith(".hmm"))
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, files, ["1_", "3_", "4_"], s2, o2,
                                                        "3_", "4_")
    assert min(db.M) < 45 and len(orows) > 500
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(o2).astype(np.int32))


@pytest.mark.parametrize("resolve", [1, 0])
def test_search_multidomain_regions(gpu_ctx, oracle, resolve):
    """Synthetic ITS1 amplicons: a few hundred regions are flagged multidomain.  resolve=1: stochastic-traceback
    ensemble + clustering (p7_domaindef) on both sides, identical cluster envelopes; resolve=0: one envelope."""
    import synth
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.1)
    rep, _, _ = oracle.derep(seq, off)
    useq, uoff, _ = _uniques(seq, off, rep)
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, [cfg["hmm_file"]],
                                                        [cfg["left_prefix"], cfg["right_prefix"]], useq, uoff,
                                                        cfg["left_prefix"], cfg["right_prefix"], resolve)
    assert ost.n_multidomain_regions > 100 and st.n_multidomain_regions == ost.n_multidomain_regions
    md = orows["is_multidomain"] != 0
    assert md.sum() > 50
    assert np.array_equal(rows["is_multidomain"] != 0, md)
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(uoff).astype(np.int32))
    # the envelopes that came out of flagged regions, explicitly
    assert np.array_equal(rows["ienv"][md], orows["ienv"][md]) and np.array_equal(rows["jenv"][md], orows["jenv"][md])
    assert np.max(np.abs(rows["domcorrection"][md] - orows["domcorrection"][md])) <= 2e-3


def test_search_multidomain_long_regions_and_iupac(gpu_ctx, oracle):
    """The resolver's less common paths: amplicons glued two and three times over (repeated boundary motifs => long
    flagged regions that need several position blocks / passes and hold several clusters), with IUPAC codes and N
    sprinkled into them (null2 odds averaged over the code's bases)."""
    import synth
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.05)
    rep, _, _ = oracle.derep(seq, off)
    useq, uoff, _ = _uniques(seq, off, rep)
    rng = np.random.default_rng(7)
    n = len(uoff) - 1
    parts = []
    for i in range(0, min(n, 240), 2):
        a, b = useq[uoff[i]:uoff[i + 1]], useq[uoff[i + 1]:uoff[i + 2]]
        glued = [a, b] if i % 3 else [a, b, a]
        v = np.concatenate(glued).copy()
        if i % 4 == 0:
            at = rng.choice(len(v), size=6, replace=False)
            v[at] = np.frombuffer(b"NRYKMS", np.uint8)
        parts.append(v)
    s2 = np.concatenate(parts)
    o2 = np.zeros(len(parts) + 1, np.int64)
    o2[1:] = np.cumsum([len(p) for p in parts])
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, [cfg["hmm_file"]],
                                                        [cfg["left_prefix"], cfg["right_prefix"]], s2, o2,
                                                        cfg["left_prefix"], cfg["right_prefix"], 1)
    assert ost.n_multidomain_regions > 100 and st.n_multidomain_regions == ost.n_multidomain_regions
    md = orows["is_multidomain"] != 0
    assert md.sum() > 100
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(o2).astype(np.int32))
    assert np.array_equal(rows["ienv"][md], orows["ienv"][md]) and np.array_equal(rows["jenv"][md], orows["jenv"][md])
    assert np.max(np.abs(rows["domcorrection"][md] - orows["domcorrection"][md])) <= 2e-3


@pytest.mark.parametrize("thr", [(0.02, 1e-3, 1e-5), (0.3, 1e-4, 1e-5)])
def test_viterbi_filter_stage(gpu_ctx, oracle, thr):
    """K6, the 16-bit Viterbi filter.  Under the reference's --F1 1e-6 --F2 1e-6 it never runs (p7_Pipeline: only for
    F2 < P <= F1); with HMMER's own default thresholds (0.02 / 1e-3 / 1e-5) and with a wide F1 it does.  The GPU stage
    (vit_kernel: integer max-plus in 1/500-bit words, saturating at -32768, overflow = pass) against the oracle's
    restatement of p7_ViterbiFilter, through the whole cascade: same rows, envelopes and positions, and the counters show
    the stage ran and removed pairs."""
    import synth
    from itsxpress_b200 import _lib
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.05)
    rng = np.random.default_rng(8)
    rnd, roff = _rand_reads(rng, 150, 240, 260, dup_frac=0.0, n_frac=0.0)          # unrelated sequences: filter fodder
    rep, _, _ = oracle.derep(seq, off)
    useq, uoff, _ = _uniques(seq, off, rep)
    s2 = np.concatenate([useq, rnd])
    o2 = np.concatenate([uoff, uoff[-1] + roff[1:]]).astype(np.int64)
    paths = [os.path.join(HMM_DIR, cfg["hmm_file"])]
    pre = [cfg["left_prefix"], cfg["right_prefix"]]
    n = gpu_ctx.load_profiles(paths, pre)
    side = gpu_ctx.set_sides_by_prefix(*pre)
    db = oracle.ProfileDB(paths, pre)
    prm = _lib.default_params()
    prm.F1, prm.F2, prm.F3 = thr
    gpu_ctx.search_seqs(s2, o2, prm)
    rows, st = gpu_ctx.hits(), gpu_ctx.search_stats()
    oprm = oracle.default_params(0, 1)
    oprm.F1, oprm.F2, oprm.F3 = thr
    orows, onrep, ost = db.search(oracle.digitize(s2.tobytes()), o2, oprm)
    assert st.n_vit_run > 100 and st.n_past_vit < st.n_past_bias and st.vit_cells > 0
    assert st.n_past_msv == ost.n_past_msv and st.n_past_bias == ost.n_past_bias
    assert st.n_past_fwd == ost.n_past_fwd              # downstream of the Viterbi filter: its pass set is the oracle's
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(o2).astype(np.int32))


def test_search_random_sequences_no_hits(gpu_ctx, oracle):
    rng = np.random.default_rng(5)
    seq, off = _rand_reads(rng, 64, 300, 420, dup_frac=0.0, n_frac=0.0)
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, ["A.hmm"], ["3_", "4_"], seq, off,
                                                        "3_", "4_")
    assert st.n_past_msv == ost.n_past_msv
    assert len(rows) == len(orows)


def _golden_positions(ids):
    tab = {}
    with open(os.path.join(os.path.dirname(__file__), "golden", "c1_positions.tsv")) as f:
        for line in f:
            if line.startswith("#"):
                continue
            k, a, b, c = line.split("\t")
            tab[k] = (int(a), int(b), int(c))
    return tab


def test_trim_against_reference_goldens(gpu_ctx, oracle, fixture_reads):
    """derep on the GPU + the reference's own (start, stop, tlen) table -> bytes of singleOut / t2_r1 / t2_r2."""
    import gzip
    from itsxpress_b200.fastq import read_fastq, format_records
    b, seq, off, qual = fixture_reads
    rep, strand, nu = gpu_ctx.derep(seq, off)
    first, _ = gpu_ctx.derep_clusters(nu)
    ids = b.ids()
    tab = _golden_positions(ids)
    start = np.full(nu, -1, np.int32); stop = np.full(nu, -1, np.int32); tlen = np.full(nu, -1, np.int32)
    for u, r in enumerate(first):
        if ids[r] in tab:
            start[u], stop[u], tlen[u] = tab[ids[r]]
    gpu_ctx.positions_set(start, stop, tlen)
    keep, lo, hi, nk = gpu_ctx.trim_bounds(b.n, mode=0)
    okeep, olo, ohi = oracle.trim_bounds(off, rep, *(np.repeat(-1, b.n).astype(np.int32) for _ in range(3)))
    # oracle indexes positions by representative READ index
    s_r = np.full(b.n, -1, np.int32); e_r = s_r.copy(); t_r = s_r.copy()
    s_r[first] = start; e_r[first] = stop; t_r[first] = tlen
    okeep, olo, ohi = oracle.trim_bounds(off, rep, s_r, e_r, t_r, mode=0)
    assert nk == 226 == int(okeep.sum())
    assert np.array_equal(keep, okeep) and np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    assert int((hi - lo)[keep == 1].sum()) == 42637
    # device gather == slices; bytes == the reference's QIIME2 single-end output
    ki, oo, os_, oq = gpu_ctx.trim_gather(b.n, mode=0, qual=qual)
    assert np.array_equal(ki, np.flatnonzero(keep))
    for t, i in enumerate(ki[:50]):
        assert os_[oo[t]:oo[t + 1]].tobytes() == seq[off[i] + lo[i]:off[i] + hi[i]].tobytes()
        assert oq[oo[t]:oo[t + 1]].tobytes() == qual[off[i] + lo[i]:off[i] + hi[i]].tobytes()
    text = format_records(b, ki, lo[ki], hi[ki])
    gold = os.path.join(TD, "singleOut", "75aea4f5-f10e-421e-91d2-feda9fe7b2e1", "data",
                        "4774-1-MSITS3_0_L001_R1_001.fastq.gz")
    assert text == gzip.open(gold, "rb").read()
    # paired, unmerged: R1 [start:stop], R2 [tlen-stop : tlen-start]
    r1 = read_fastq(os.path.join(TD, "4774-1-MSITS3_R1.fastq"))
    r2 = read_fastq(os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    id1 = {k: i for i, k in enumerate(r1.ids())}
    order = np.array([id1[k] for k in ids])           # merged reads are a subset of the pairs, same order
    assert np.all(np.diff(order) > 0)
    for mode, rb, goldf in ((2, r1, "t2_r1.fq"), (1, r2, "t2_r2.fq")):
        s_off = np.zeros(b.n + 1, np.int64)
        s_off[1:] = np.cumsum(rb.s_len[order])
        k2, l2, h2, nk2 = gpu_ctx.trim_bounds(b.n, mode=mode, off_other=s_off)
        ok2, ol2, oh2 = oracle.trim_bounds(s_off if mode == 2 else off, rep, s_r, e_r, t_r, mode=mode, off_r2=s_off)
        assert np.array_equal(k2, ok2) and np.array_equal(l2, ol2) and np.array_equal(h2, oh2)
        sel = np.flatnonzero(k2)
        text = format_records(rb, order[sel], l2[sel], h2[sel])
        assert text == open(os.path.join(TD, goldf), "rb").read()


def test_whole_path_fixture(gpu_ctx, oracle, fixture_reads):
    """itsx_run (host buffers in/out) == oracle pipeline on the bundled sample with Metazoa ITS2 profiles."""
    b, seq, off, _ = fixture_reads
    paths = [os.path.join(HMM_DIR, "M.hmm")]
    gpu_ctx.load_profiles(paths, ["3_", "4_"])
    side = gpu_ctx.set_sides_by_prefix("3_", "4_")
    out, st = gpu_ctx.run(seq, off)
    orep, _, onu = oracle.derep(seq, off)
    useq, uoff, idx = _uniques(seq, off, orep)
    db = oracle.ProfileDB(paths, ["3_", "4_"])
    orows, onrep, ost = db.search(oracle.digitize(useq.tobytes()), uoff)
    opos = oracle.itspos(orows, side, np.diff(uoff).astype(np.int32))
    s_r = np.full(b.n, -1, np.int32); e_r = s_r.copy(); t_r = s_r.copy()
    s_r[idx] = opos["start"]; e_r[idx] = opos["stop"]; t_r[idx] = opos["tlen"]
    okeep, olo, ohi = oracle.trim_bounds(off, orep, s_r, e_r, t_r, mode=0)
    assert st.n_unique == onu and st.n_kept == int(okeep.sum()) > 100
    assert np.array_equal(out["rep"], orep)
    assert np.array_equal(out["keep"], okeep)
    assert np.array_equal(out["lo"], olo) and np.array_equal(out["hi"], ohi)


@pytest.fixture(scope="module")
def owner_ctx():
    """second context on the same GPU: the owner side of a sharded run"""
    from itsxpress_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def test_run_sharded_single_rank_equals_itsx_run(gpu_ctx, owner_ctx, fixture_reads):
    """distributed.run_sharded with the GPU engine (csrc/shard.cu through device pointers) on one rank == itsx_run on
    the same reads: representatives, strands, bounds, and the re-expanded slices == itsx_run_trim."""
    from itsxpress_b200.distributed import Comm, GpuEngine, run_sharded
    b, seq, off, qual = fixture_reads
    # make some reads reverse complements of others so that strand '-' occurs
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    parts = [seq[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
    quals = [qual[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
    parts += [p.translate(comp)[::-1] for p in parts[:20]]
    quals += [q[::-1] for q in quals[:20]]
    seq2 = np.frombuffer(b"".join(parts), np.uint8).copy()
    qual2 = np.frombuffer(b"".join(quals), np.uint8).copy()
    off2 = np.zeros(len(parts) + 1, np.int64)
    off2[1:] = np.cumsum([len(p) for p in parts])
    paths = [os.path.join(HMM_DIR, "M.hmm")]
    for c in (gpu_ctx, owner_ctx):
        c.load_profiles(paths, ["3_", "4_"])
        c.set_sides_by_prefix("3_", "4_")
    want, st = gpu_ctx.run(seq2, off2)
    want = {k: v.copy() for k, v in want.items()}
    _, wstrand, _ = gpu_ctx.derep(seq2, off2)
    wt, st2 = gpu_ctx.run_trim(seq2, qual2, off2)
    wt = {k: v.copy() for k, v in wt.items()}
    eng = GpuEngine(gpu_ctx, owner_ctx)
    phases = {}
    got = run_sharded(eng, Comm(), seq2, off2, 0, phases=phases)
    assert got["n_unique_global"] == st.n_unique == got["n_owned"] == got["n_local_unique"]
    assert np.array_equal(got["rep"], want["rep"])
    assert np.array_equal(got["strand"], wstrand) and got["strand"].sum() >= 20
    assert np.array_equal(got["keep"], want["keep"]) and int(got["keep"].sum()) == st.n_kept
    assert np.array_equal(got["lo"], want["lo"]) and np.array_equal(got["hi"], want["hi"])
    assert set(phases) == set(__import__("itsxpress_b200.distributed", fromlist=["PHASES"]).PHASES)
    # block resident ahead of the call, re-expansion at the end
    eng.upload(seq2, off2, qual2)
    g2 = run_sharded(eng, Comm(), None, None, 0, want_rep=False, gather=True)
    for k in ("kept_index", "out_off", "out_seq", "out_qual"):
        assert np.array_equal(g2[k], wt[k]), k
    assert len(wt["kept_index"]) == st2.n_kept == st.n_kept and st2.out_bytes == len(wt["out_seq"]) > 10000
    # the bench's legs: slices left in HBM (value), then fetched into preallocated worst-case buffers (e2e)
    g3 = run_sharded(eng, Comm(), None, None, 0, want_rep=False, gather="device")
    assert g3["n_kept"] == st2.n_kept and g3["out_bytes"] == st2.out_bytes
    big = dict(kept_index=np.empty(len(off2) - 1, np.int32), out_off=np.empty(len(off2), np.int64),
               out_seq=np.empty(len(seq2), np.uint8), out_qual=np.empty(len(seq2), np.uint8))
    eng.resident = eng.resident_qual = False
    g4 = run_sharded(eng, Comm(), seq2, off2, 0, want_rep=False, gather=True, qual=qual2, gather_out=big)
    for k in ("kept_index", "out_off", "out_seq", "out_qual"):
        assert np.array_equal(g4[k], wt[k]), k
    # an empty block is legal (more ranks than reads)
    eng2 = GpuEngine(gpu_ctx, owner_ctx)
    empty = run_sharded(eng2, Comm(), np.zeros(0, np.uint8), np.zeros(1, np.int64), 0)
    assert len(empty["keep"]) == 0 and empty["n_unique_global"] == 0


def test_run_trim_equals_bounds_plus_slices(gpu_ctx, fixture_reads):
    """itsx_run_trim (the e2e call: qualities in, trimmed records out) == itsx_run's bounds applied on the host, and
    the resident path (reads_upload + quals_upload + run_resident + run_fetch) gives the same bytes."""
    b, seq, off, qual = fixture_reads
    gpu_ctx.load_profiles([os.path.join(HMM_DIR, "M.hmm")], ["3_", "4_"])
    gpu_ctx.set_sides_by_prefix("3_", "4_")
    want, st = gpu_ctx.run(seq, off)
    keep, lo, hi = want["keep"].copy(), want["lo"].copy(), want["hi"].copy()
    got, st2 = gpu_ctx.run_trim(seq, qual, off)
    got = {k: v.copy() for k, v in got.items()}
    ki = np.flatnonzero(keep)
    assert np.array_equal(got["kept_index"], ki) and st2.n_kept == len(ki) == st.n_kept
    exp_seq = np.concatenate([seq[off[i] + lo[i]:off[i] + hi[i]] for i in ki])
    exp_qual = np.concatenate([qual[off[i] + lo[i]:off[i] + hi[i]] for i in ki])
    assert np.array_equal(got["out_seq"], exp_seq) and np.array_equal(got["out_qual"], exp_qual)
    assert np.array_equal(np.diff(got["out_off"]), (hi - lo)[ki]) and st2.out_bytes == len(exp_seq)
    gpu_ctx.reads_upload(seq, off)
    gpu_ctx.quals_upload(qual)
    st3 = gpu_ctx.run_resident()
    ki3, oo3, os3, oq3 = gpu_ctx.run_fetch()
    assert st3.n_kept == len(ki) and np.array_equal(ki3, ki) and np.array_equal(os3, exp_seq) and np.array_equal(oq3, exp_qual)
    assert np.array_equal(oo3, got["out_off"])


def test_vector_gather_all_alignments(gpu_ctx):
    """the 16-byte-store copy of trim / shard kernels (warp_copy) against numpy for every source and destination
    alignment and for slice lengths around the 16- and 32-byte edges"""
    rng = np.random.default_rng(11)
    lens = np.concatenate([np.arange(0, 70), rng.integers(0, 400, 600)]).astype(np.int64)
    n = len(lens)
    rl = lens + rng.integers(0, 40, n)                 # read lengths; the slice starts at a random offset inside
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(rl)
    seq = rng.integers(33, 127, int(off[-1]), dtype=np.uint8)
    qual = rng.integers(33, 127, int(off[-1]), dtype=np.uint8)
    start = rng.integers(0, (rl - lens) + 1)
    stop = start + lens
    gpu_ctx.trim_set_map(np.arange(n, dtype=np.int32), n)
    gpu_ctx.positions_set(start.astype(np.int32), stop.astype(np.int32), rl.astype(np.int32))
    ki, oo, os_, oq = gpu_ctx.trim_gather(n, mode=0, seq=seq, qual=qual, off=off)
    want_k = np.flatnonzero(lens > 0)
    assert np.array_equal(ki, want_k)
    exp_s = np.concatenate([seq[off[i] + start[i]:off[i] + stop[i]] for i in want_k])
    exp_q = np.concatenate([qual[off[i] + start[i]:off[i] + stop[i]] for i in want_k])
    assert np.array_equal(os_, exp_s) and np.array_equal(oq, exp_q)
    assert np.array_equal(np.diff(oo), lens[want_k])


def _winners_from_multidomain(rows, side):
    """ItsPosition restated on the oracle's rows (table order, strict > on the printed 0.1-bit score: first row wins,
    SeqSample.py:400-429): how many selected left / right boundaries came out of a multidomain region."""
    sd = side[rows["prof"]]
    ok = sd >= 0
    r = rows[ok]
    sd = sd[ok]
    s10 = np.rint(r["bitscore"].astype(np.float64) * 10.0).astype(np.int64)
    order = np.lexsort((np.arange(len(r)), -s10, sd, r["seq"]))        # per (seq, side): best score, earliest row
    key = r["seq"][order].astype(np.int64) * 2 + sd[order]
    firsts = np.concatenate([[True], key[1:] != key[:-1]]) if len(order) else np.zeros(0, bool)
    return int((r["is_multidomain"][order][firsts] != 0).sum())


@pytest.mark.parametrize("config,scale,frac_mod", [("c2", 1.0, 12), ("c3s", 0.05, 1), ("c4s", 0.02, 1)])
def test_bench_workloads_against_oracle(gpu_ctx, oracle, config, scale, frac_mod):
    """the bench's own workloads against the oracle, whole path.  c2 = bench.cpu_sample of the FULL
    BASELINE configs[1] sample (all reads of every 12th unique: 75 k reads, 25 k uniques, 98 profiles -- the sample the
    cpu_baseline leg times); c3s = configs[2] shape (--region ALL --taxa All: 340 profiles of every taxon file, 380-520
    bp); c4s = configs[3] shape (ITS2, 90 % unique, 330-441 bp, 256 profiles).  Compared: derep classes, the MSV pass
    count, per-profile reported hits, every unique's ItsPosition fields, keep / lo / hi of every read, and how many
    selected boundaries came out of multidomain regions."""
    import synth
    import bench
    seq, off, which, cfg = synth.make_config(config, scale=scale)
    if frac_mod > 1:
        seq, off, _ = bench.cpu_sample(seq, off, which, frac_mod=frac_mod)
    paths = [os.path.join(HMM_DIR, f) for f in cfg["search_files"]]
    pre = [cfg["left_prefix"], cfg["right_prefix"]]
    nprof = gpu_ctx.load_profiles(paths, pre)
    side = gpu_ctx.set_sides_by_prefix(*pre)
    got, st = gpu_ctx.run(seq, off)
    got = {k: v.copy() for k, v in got.items()}
    ss = gpu_ctx.search_stats()
    pos = gpu_ctx.positions(st.n_unique)
    nrep = gpu_ctx.nreported().copy()
    db = oracle.ProfileDB(paths, pre)
    assert db.n == nprof and (config != "c3s" or nprof == 340)
    o, ost = bench.oracle_pipeline(oracle, db, side, seq, off, threads=os.cpu_count(), want_pos=True)
    assert st.n_unique == len(o["first"]) and np.array_equal(got["rep"], o["rep"])
    assert ss.n_past_msv == ost.n_past_msv and ss.n_pairs == ost.npairs_total
    assert ss.n_past_bias == ost.n_past_bias and ss.n_past_fwd == ost.n_past_fwd
    assert ss.n_multidomain_regions == ost.n_multidomain_regions
    assert np.array_equal(nrep, o["nrep"]) and ss.n_domains_reported == len(o["rows"])
    for k in ("start", "stop", "tlen", "left_from", "left_to", "right_from", "right_to"):
        assert np.array_equal(pos[k], o["pos"][k]), k
    for k in ("left_score10", "right_score10"):
        assert np.max(np.abs(pos[k].astype(np.int64) - o["pos"][k].astype(np.int64))) <= 1, k
    assert np.array_equal(got["keep"], o["keep"]) and int(o["keep"].sum()) == st.n_kept > 0.5 * len(o["keep"])
    assert np.array_equal(got["lo"], o["lo"]) and np.array_equal(got["hi"], o["hi"])
    assert ss.n_selected_multidomain == _winners_from_multidomain(o["rows"], side)
    if config == "c2":
        assert ss.n_multidomain_regions > 1000       # the resolver is exercised at scale


@pytest.mark.parametrize("config,scale", [("c2_small", 0.3), ("c4s", 0.01)])
def test_compact_rows_mode_gives_the_same_positions(gpu_ctx, config, scale):
    """keep_rows = 2 (rows that are printed for every possible domZ enter ItsPosition's arg-max at once; only undecided
    rows that would beat that winner are stored) == keep_rows = 1 (every row kept until domZ is known): positions,
    printed scores, reported-hit counts and the multidomain-winner count, for the sequence count and for a loose
    domz_upper; the stored rows are a small fraction; stage 2 cannot be applied twice to a compact search."""
    import synth
    from itsxpress_b200 import _lib
    seq, off, which, cfg = synth.make_config(config, scale=scale)
    gpu_ctx.load_profiles([os.path.join(HMM_DIR, f) for f in cfg["search_files"]], [cfg["left_prefix"], cfg["right_prefix"]])
    gpu_ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    res = {}
    for mode, zup in ((1, 0), (2, 0), (2, 100_000_000)):
        prm = _lib.default_params()
        prm.keep_rows, prm.domz_upper = mode, zup
        out, st = gpu_ctx.run(seq, off, prm)
        ss = gpu_ctx.search_stats()
        res[(mode, zup)] = (dict((k, v.copy()) for k, v in out.items()), gpu_ctx.positions(st.n_unique),
                            gpu_ctx.nreported().copy(), ss.n_selected_multidomain, ss.n_domains, len(gpu_ctx.hits()),
                            st.n_kept)
    ref = res[(1, 0)]
    assert ref[6] > 0.5 * (len(off) - 1) and ref[4] == res[(2, 0)][4] > 0
    for key in ((2, 0), (2, 100_000_000)):
        got = res[key]
        for k in ("rep", "keep", "lo", "hi"):
            assert np.array_equal(got[0][k], ref[0][k]), (key, k)
        for k in ref[1]:
            assert np.array_equal(got[1][k], ref[1][k]), (key, k)
        assert np.array_equal(got[2], ref[2]) and got[3] == ref[3] and got[6] == ref[6]
        assert got[5] < 0.1 * ref[5]                    # rows kept on the device: a small fraction of the table
    with pytest.raises(_lib.ItsxError):
        gpu_ctx.search_stage2()


@pytest.mark.parametrize("keep_rows", [1, 2])
def test_several_samples_in_one_pass(gpu_ctx, keep_rows):
    """the samples of an artifact batched into ONE device pass (itsx_reads_set_samples: sample id folded into
    the derep key and the class test, reported hits counted per (sample, profile)) == the reference's sequential loop
    over samples (q2_itsxpress.py:273-333): per sample the same classes, domZ, positions, keep / lo / hi and bytes.
    The samples share sequences (which must NOT merge across samples) and differ in size (so their domZ differ)."""
    import synth
    from itsxpress_b200 import _lib
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.5)
    qual = synth.make_quals(3, off)
    n = len(off) - 1
    cuts = [0, n // 8, n // 2, n]                      # three samples of very different sizes over the same unique pool
    paths = [os.path.join(HMM_DIR, cfg["hmm_file"])]
    gpu_ctx.load_profiles(paths, [cfg["left_prefix"], cfg["right_prefix"]])
    gpu_ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    prm = _lib.default_params()
    prm.keep_rows = keep_rows
    prm.domE = 1e-3            # a tighter domE than the reference's 10: domZ then really decides rows (10-20 bit domains)
    alone = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s_, q_, o_ = seq[off[a]:off[b]], qual[off[a]:off[b]], off[a:b + 1] - off[a]
        out, st = gpu_ctx.run(s_, o_, prm)
        out = {k: v.copy() for k, v in out.items()}
        nrep = gpu_ctx.nreported().copy()
        tr, _ = gpu_ctx.run_trim(s_, q_, o_, prm)
        alone.append((out, nrep, {k: v.copy() for k, v in tr.items()}, st.n_unique))
    assert len({tuple(x[1]) for x in alone}) == 3          # the three samples have different domZ vectors
    sample = np.repeat(np.arange(3, dtype=np.int32), np.diff(cuts))
    gpu_ctx.reads_upload(seq, off)
    gpu_ctx.quals_upload(qual)
    gpu_ctx.set_samples(sample, 3)
    st = gpu_ctx.run_resident(prm)
    ki, oo, os_, oq = gpu_ctx.run_fetch()
    nrep_b = gpu_ctx.nreported(3).copy()
    assert st.n_unique == sum(x[3] for x in alone)
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        out, nrep, tr, nu = alone[k]
        assert np.array_equal(nrep_b[k], nrep), k
        lo_k, hi_k = np.searchsorted(ki, a), np.searchsorted(ki, b)
        assert np.array_equal(ki[lo_k:hi_k] - a, tr["kept_index"]), k
        assert np.array_equal(os_[oo[lo_k]:oo[hi_k]], tr["out_seq"]) and np.array_equal(oq[oo[lo_k]:oo[hi_k]], tr["out_qual"]), k
        assert np.array_equal(oo[lo_k:hi_k + 1] - oo[lo_k], tr["out_off"]), k


def test_trim_set_map_drops_unmapped_reads(gpu_ctx):
    off = np.array([0, 10, 20, 30], np.int64)
    gpu_ctx.trim_set_map(np.array([0, -1, 1], np.int32), 2)
    gpu_ctx.positions_set(np.array([2, 0], np.int32), np.array([8, 4], np.int32), np.array([10, 10], np.int32))
    keep, lo, hi, nk = gpu_ctx.trim_bounds(3, mode=0, off_other=off)
    assert keep.tolist() == [1, 0, 1] and nk == 2
    assert (lo[0], hi[0], lo[2], hi[2]) == (2, 8, 0, 4)
    with pytest.raises(Exception):
        gpu_ctx.trim_bounds(3, mode=0)           # offsets are mandatory once the map is external


def test_full_size_properties(gpu_ctx):
    """BASELINE configs[1] at FULL size (1 M reads x 250 bp, 300 k uniques) through itsx_run: properties that do
    not need the oracle -- the planted class structure is recovered exactly, representatives are first
    occurrences and idempotent, abundances add up, every kept slice is a non-empty in-range slice, and a second
    run over the representatives alone returns the same boundaries."""
    import synth
    seq, off, which, cfg = synth.make_config("c2", scale=1.0)
    n = len(off) - 1
    gpu_ctx.load_profiles([os.path.join(HMM_DIR, cfg["hmm_file"])], [cfg["left_prefix"], cfg["right_prefix"]])
    gpu_ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    out, st = gpu_ctx.run(seq, off)
    rep, keep, lo, hi = out["rep"].copy(), out["keep"].copy(), out["lo"].copy(), out["hi"].copy()
    assert st.n_reads == n == 1_000_000 and st.n_unique == 300_000
    # derep == the generator's planted classes (no planted read is the reverse complement of another)
    first_of_class = np.full(which.max() + 1, n, np.int64)
    np.minimum.at(first_of_class, which, np.arange(n))
    assert np.array_equal(rep, first_of_class[which])
    assert np.all(rep <= np.arange(n)) and np.array_equal(rep[rep], rep)
    first, ab = gpu_ctx.derep_clusters(st.n_unique)
    assert int(ab.sum()) == n and np.all(np.diff(first) > 0)
    # trim: members of a class share the decision and the bounds; slices are in range and non-empty
    assert np.array_equal(keep, keep[rep]) and np.array_equal(lo, lo[rep]) and np.array_equal(hi, hi[rep])
    k = keep == 1
    assert st.n_kept == int(k.sum()) > 0.5 * n
    lens = np.diff(off)
    assert np.all(lo[k] >= 0) and np.all(hi[k] <= lens[k]) and np.all(lo[k] < hi[k])
    # idempotence: searching only the representatives gives every class the same boundaries
    idx = first.astype(np.int64)
    sub_lens = lens[idx]
    sub_off = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(sub_lens, out=sub_off[1:])
    delta = np.repeat(off[idx] - sub_off[:-1], sub_lens)
    sub = seq[delta + np.arange(int(sub_off[-1]), dtype=np.int64)]
    out2, st2 = gpu_ctx.run(sub, sub_off)
    assert st2.n_unique == len(idx)
    assert np.array_equal(out2["keep"], keep[idx]) and np.array_equal(out2["lo"], lo[idx])
    assert np.array_equal(out2["hi"], hi[idx])


def test_search_short_and_empty_sequences(gpu_ctx, oracle, fixture_reads):
    """Ragged input: empty, 1-base, shorter-than-the-model and ordinary sequences in one search; and a search
    over zero sequences."""
    b, seq, off, _ = fixture_reads
    parts = [np.zeros(0, np.uint8), seq[off[3]:off[3] + 1], seq[off[4]:off[4] + 7], seq[off[5]:off[5] + 44],
             seq[off[6]:off[7]], np.zeros(0, np.uint8), seq[off[8]:off[8] + 130], seq[off[9]:off[10]]]
    s2 = np.concatenate(parts)
    o2 = np.zeros(len(parts) + 1, np.int64)
    o2[1:] = np.cumsum([len(p) for p in parts])
    rows, st, orows, onrep, ost, side, db = _search_both(gpu_ctx, oracle, ["M.hmm"], ["3_", "4_"], s2, o2, "3_", "4_")
    assert len(orows) > 50
    _compare_search(gpu_ctx, oracle, rows, st, orows, onrep, ost, side, np.diff(o2).astype(np.int32))
    gpu_ctx.search_seqs(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(gpu_ctx.hits()) == 0


def test_derep_then_trim_ragged(gpu_ctx, oracle):
    """Empty reads and single-base reads go through derep and the whole path without special cases."""
    reads = [b"", b"A", b"", b"ACGT" * 70, b"T", b"ACGT" * 70, b"a"]
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    seq = np.frombuffer(b"".join(reads), np.uint8).copy()
    rep, strand, nu = gpu_ctx.derep(seq, off)
    orep, ostrand, onu = oracle.derep(seq, off)
    assert nu == onu and np.array_equal(rep, orep) and np.array_equal(strand, ostrand)
    gpu_ctx.load_profiles([os.path.join(HMM_DIR, "A.hmm")], ["3_", "4_"])
    gpu_ctx.set_sides_by_prefix("3_", "4_")
    out, st = gpu_ctx.run(seq, off)
    assert st.n_unique == onu and st.n_kept == 0
