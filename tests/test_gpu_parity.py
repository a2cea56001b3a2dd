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


def test_run_sharded_single_rank_equals_itsx_run(gpu_ctx, fixture_reads):
    """distributed.run_sharded with the GPU engine on one rank == itsx_run on the same reads."""
    from itsxpress_b200.distributed import Comm, GpuEngine, run_sharded
    b, seq, off, _ = fixture_reads
    paths = [os.path.join(HMM_DIR, "M.hmm")]
    gpu_ctx.load_profiles(paths, ["3_", "4_"])
    gpu_ctx.set_sides_by_prefix("3_", "4_")
    want, st = gpu_ctx.run(seq, off)
    want = {k: v.copy() for k, v in want.items()}
    got = run_sharded(GpuEngine(gpu_ctx), Comm(), seq, off, 0)
    assert got["n_unique_global"] == st.n_unique
    assert np.array_equal(got["rep"], want["rep"])
    assert np.array_equal(got["keep"], want["keep"]) and int(got["keep"].sum()) == st.n_kept
    assert np.array_equal(got["lo"], want["lo"]) and np.array_equal(got["hi"], want["hi"])


def test_run_sharded_device_single_rank_equals_itsx_run(gpu_ctx, fixture_reads):
    """the device-resident sharded driver (C ABI called with device pointers, torch index arithmetic on the GPU) on one
    rank == itsx_run == the host-orchestrated driver, including strands and the per-profile reported-hit counts."""
    from itsxpress_b200.distributed import Comm, GpuEngine, run_sharded, run_sharded_device
    b, seq, off, _ = fixture_reads
    # make some reads reverse complements of others so that strand '-' occurs
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    parts = [seq[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
    parts += [p.translate(comp)[::-1] for p in parts[:20]]
    seq2 = np.frombuffer(b"".join(parts), np.uint8).copy()
    off2 = np.zeros(len(parts) + 1, np.int64)
    off2[1:] = np.cumsum([len(p) for p in parts])
    paths = [os.path.join(HMM_DIR, "M.hmm")]
    gpu_ctx.load_profiles(paths, ["3_", "4_"])
    gpu_ctx.set_sides_by_prefix("3_", "4_")
    want, st = gpu_ctx.run(seq2, off2)
    want = {k: v.copy() for k, v in want.items()}
    host = run_sharded(GpuEngine(gpu_ctx), Comm(), seq2, off2, 0)
    got = run_sharded_device(gpu_ctx, seq2, off2, 0)
    assert got["n_unique_global"] == st.n_unique
    assert np.array_equal(got["rep"], want["rep"])
    assert np.array_equal(got["keep"], want["keep"]) and int(got["keep"].sum()) == st.n_kept
    assert np.array_equal(got["lo"], want["lo"]) and np.array_equal(got["hi"], want["hi"])
    assert np.array_equal(got["strand"], host["strand"]) and got["strand"].sum() >= 20
    assert np.array_equal(got["nreported"], host["nreported"])
    # an empty block is legal (more ranks than reads)
    empty = run_sharded_device(gpu_ctx, np.zeros(0, np.uint8), np.zeros(1, np.int64), 0)
    assert len(empty["keep"]) == 0 and empty["n_unique_global"] == 0


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
