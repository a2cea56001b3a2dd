"""GPU parity of the paired-end merge (csrc/merge.cu, itsx_merge_pairs / itsx_merge_fetch through the C ABI) against
the CPU oracle (oracle/ora_merge.c) and the committed golden vectors: decisions, merged bases and merged qualities are
byte-identical.  Reference call site: SeqSamplePairedNotInterleaved._merge_reads, itsxpress/SeqSample.py:266-365."""
import os
import subprocess
import zlib

import numpy as np
import pytest

from conftest import TD
from itsxpress_b200 import _lib
from itsxpress_b200.fastq import read_fastq

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_merge.tsv")


def _load(r1, r2):
    b1, b2 = read_fastq(os.path.join(TD, r1)), read_fastq(os.path.join(TD, r2))
    fs, fo = b1.seq_concat()
    rs, ro = b2.seq_concat()
    return b1, b2, fs, b1.qual_concat()[0], fo, rs, b2.qual_concat()[0], ro


def _compare(oracle, ctx, fs, fq, fo, rs, rq, ro, stagger, **kw):
    oml, owhy, oseq, oqual = oracle.merge_pairs(fs, fq, fo, rs, rq, ro, oracle.merge_params(allow_stagger=stagger, **kw))
    ml, why, idx, ooff, gseq, gqual = ctx.merge_pairs(fs, fq, fo, rs, rq, ro, _lib.merge_params(allow_stagger=stagger, **kw))
    assert np.array_equal(why, owhy)
    assert np.array_equal(ml, oml)
    assert np.array_equal(idx, np.flatnonzero(oml > 0))
    assert np.array_equal(np.diff(ooff), oml[idx])
    # slot layout of the oracle -> packed layout of the library
    slot = (np.asarray(fo[:-1]) + np.asarray(ro[:-1]))[idx]
    src = np.repeat(slot - ooff[:-1], oml[idx]) + np.arange(int(ooff[-1]), dtype=np.int64)
    assert np.array_equal(gseq, oseq[src])
    assert np.array_equal(gqual, oqual[src])
    st = ctx.merge_stats()
    assert st.n_pairs == len(ml) and st.n_merged == len(idx)
    assert [st.by_reason[r] for r in range(10)] == np.bincount(owhy, minlength=10).tolist()
    return ml, why, idx, ooff, gseq, gqual


@pytest.mark.parametrize("r1,r2", [("4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq"),
                                   ("high_qual_scores_R1.fastq.gz", "high_qual_scores_R2.fastq.gz")])
@pytest.mark.parametrize("stagger", [0, 1])
def test_fixture_pairs_match_oracle_and_golden(oracle, gpu_ctx, r1, r2, stagger):
    b1, b2, fs, fq, fo, rs, rq, ro = _load(r1, r2)
    ml, why, idx, ooff, gseq, gqual = _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, stagger)
    name = r1.split("_R1")[0]
    k = 0
    for line in open(GOLD):
        f = line.rstrip("\n").split("\t")
        if line.startswith("#") or f[0] != name or int(f[1]) != stagger:
            continue
        i = int(f[2])
        assert _lib.MERGE_REASONS[why[i]] == f[3] and ml[i] == int(f[4])
        if ml[i]:
            assert idx[k] == i
            assert zlib.crc32(gseq[ooff[k]:ooff[k + 1]].tobytes()) == int(f[5])
            assert zlib.crc32(gqual[ooff[k]:ooff[k + 1]].tobytes()) == int(f[6])
            k += 1
    assert k == len(idx)


@pytest.mark.parametrize("stagger", [0, 1])
def test_synthetic_pairs_match_oracle(oracle, gpu_ctx, stagger):
    """Ragged input: fragments shorter and longer than the reads (staggered / no overlap), 3'-trimmed reads, sequencing
    errors, N, lower case, empty reads."""
    import synth
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(5, 30000, frag_len=(150, 480), trim=(0, 40), lower_rate=0.05,
                                                          empty_rate=0.002)
    ml, why, *_ = _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, stagger)
    hist = np.bincount(why, minlength=10)
    assert hist[0] > 10000 and hist[5] > 0 and hist[6] > 0 and hist[8] > 0      # ok, nokmers, minscore, maxee all occur
    assert (hist[2] > 0) == (not stagger)


def test_other_read_lengths_and_options(oracle, gpu_ctx):
    import synth
    # 2 x 301 (MiSeq v3) with a tight mismatch budget and a long minimum overlap
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(9, 4000, frag_len=(300, 590), read_len=301, err_scale=3.0)
    _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, 0, maxdiffs=5, minovlen=30)
    # short reads, unequal lengths, very noisy, lax expected-error filter
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(10, 4000, frag_len=(40, 200), read_len=75, trim=(0, 30),
                                                          err_scale=5.0, n_rate=0.02)
    _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, 1, maxee=8.0)
    # qualities 0 and 1 (error probability pinned at 0.75; log-odds round to +-1e-16 there)
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(13, 4000, frag_len=(260, 480), err_scale=2.0)
    rng = np.random.default_rng(13)
    fq, rq = fq.copy(), rq.copy()
    fq[rng.random(len(fq)) < 0.15] = 33
    rq[rng.random(len(rq)) < 0.15] = 34
    fq[rng.random(len(fq)) < 0.05] = 34
    ml, why, *_ = _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, 0, maxee=200.0)
    assert (ml > 0).sum() > 1000
    # long reads (2 x 1000) exercise the multi-word planes and the shared-memory layout
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(12, 300, frag_len=(900, 1900), read_len=1000, err_scale=0.2)
    _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, 0, maxee=50.0)
    # low-complexity fragments: tandem repeats make several diagonals score -> "repeat"
    rng = np.random.default_rng(4)
    frags = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(4, 30)))) * 40 for _ in range(500)]
    flen = np.array([min(len(f), 400) for f in frags], np.int64)
    foff = np.zeros(len(frags) + 1, np.int64)
    foff[1:] = np.cumsum(flen)
    fcat = np.frombuffer(b"".join(f[:400] for f in frags), np.uint8)
    fs, fq, fo, rs, rq, ro = synth.make_pairs(21, fcat, foff, read_len=250)
    ml, why, *_ = _compare(oracle, gpu_ctx, fs, fq, fo, rs, rq, ro, 0)
    assert np.bincount(why, minlength=10)[1] > 100


def test_empty_and_degenerate_input(oracle, gpu_ctx):
    z8, z64 = np.zeros(0, np.uint8), np.zeros(1, np.int64)
    ml, why, idx, ooff, gseq, gqual = gpu_ctx.merge_pairs(z8, z8, z64, z8, z8, z64)
    assert len(ml) == 0 and len(idx) == 0 and ooff.tolist() == [0] and len(gseq) == 0
    # only empty reads; one empty mate
    fs = np.frombuffer(b"ACGTACGTACGT", np.uint8)
    fq = np.frombuffer(b"IIIIIIIIIIII", np.uint8)
    fo = np.array([0, 0, 12, 12], np.int64)
    ro = np.array([0, 12, 12, 12], np.int64)
    _compare(oracle, gpu_ctx, fs, fq, fo, fs, fq, ro, 1)
    with pytest.raises(ValueError):
        gpu_ctx.merge_pairs(fs, fq, fo, fs, fq, ro[:-1])


def test_quality_above_qmax_is_an_error(gpu_ctx):
    import synth
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(3, 100, frag_len=(300, 400))
    fq = fq.copy()
    fq[1234] = 33 + 94
    with pytest.raises(_lib.ItsxError) as e:
        gpu_ctx.merge_pairs(fs, fq, fo, rs, rq, ro)
    assert e.value.code == -7 and "qmax" in str(e.value)
    rq = rq.copy()
    rq[77] = 32
    fq[1234] = 70
    with pytest.raises(_lib.ItsxError):
        gpu_ctx.merge_pairs(fs, fq, fo, rs, rq, ro)
    # the context stays usable
    ml, *_ = gpu_ctx.merge_pairs(fs, fq, fo, rs, np.maximum(rq, 33), ro)
    assert (ml > 0).sum() > 50


def test_merge_reads_api_writes_seq_fq(oracle, tmp_path):
    """SeqSamplePairedNotInterleaved._merge_reads (reference tests/test_main_pytest.py:182-194): seq.fq holds the merged
    records in input order under R1's title, plain / .gz inputs alike, reversed_primers swaps the mates."""
    from itsxpress_b200.SeqSample import SeqSamplePairedNotInterleaved
    b1, b2, fs, fq, fo, rs, rq, ro = _load("4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq")
    oml, owhy, oseq, oqual = oracle.merge_pairs(fs, fq, fo, rs, rq, ro, oracle.merge_params(allow_stagger=True))
    want = []
    for i in np.flatnonzero(oml > 0):
        s = int(fo[i] + ro[i])
        want.append("@%s\n%s\n+\n%s\n" % (b1.title(i), oseq[s:s + oml[i]].tobytes().decode(),
                                          oqual[s:s + oml[i]].tobytes().decode()))
    for ext in ("", ".gz"):
        td = str(tmp_path / ("t" + ext.strip(".")))
        sobj = SeqSamplePairedNotInterleaved(fastq=os.path.join(TD, "4774-1-MSITS3_R1.fastq" + ext), tempdir=td,
                                             fastq2=os.path.join(TD, "4774-1-MSITS3_R2.fastq" + ext))
        sobj._merge_reads(stagger=True, threads=1)
        assert sobj.seq_file == os.path.join(td, "seq.fq")
        assert open(sobj.seq_file).read() == "".join(want)
        sobj.deduplicate(threads=1)                      # the merged file feeds the next stage as upstream
        assert os.path.getsize(sobj.rep_file) > 0
    # reversed primers: R2 is the forward read
    td = str(tmp_path / "rev")
    sobj = SeqSamplePairedNotInterleaved(fastq=os.path.join(TD, "4774-1-MSITS3_R1.fastq"), tempdir=td,
                                         fastq2=os.path.join(TD, "4774-1-MSITS3_R2.fastq"), reversed_primers=True)
    sobj._merge_reads(stagger=False, threads="4")
    rml, _, _, _ = oracle.merge_pairs(rs, rq, ro, fs, fq, fo)
    got = read_fastq(sobj.seq_file)
    assert got.n == int((rml > 0).sum()) and got.title(0) == b2.title(int(np.flatnonzero(rml > 0)[0]))


def test_merge_reads_errors(tmp_path):
    from itsxpress_b200.SeqSample import SeqSamplePairedNotInterleaved
    r1 = os.path.join(TD, "4774-1-MSITS3_R1.fastq")
    # a file that is not FASTQ: vsearch exits non-zero -> CalledProcessError after logging (SeqSample.py:351-358)
    with pytest.raises(subprocess.CalledProcessError):
        SeqSamplePairedNotInterleaved(fastq=r1, tempdir=str(tmp_path / "a"),
                                      fastq2=os.path.join(TD, "broken.fastq"))._merge_reads(1, False)
    # different record counts
    short = tmp_path / "short.fastq"
    short.write_text("".join(open(os.path.join(TD, "4774-1-MSITS3_R2.fastq")).readlines()[:400]))
    with pytest.raises(subprocess.CalledProcessError) as e:
        SeqSamplePairedNotInterleaved(fastq=r1, tempdir=str(tmp_path / "b"), fastq2=str(short))._merge_reads(1, False)
    assert b"More forward reads" in e.value.stderr
    with pytest.raises(FileNotFoundError):
        SeqSamplePairedNotInterleaved(fastq=r1, tempdir=str(tmp_path / "c"),
                                      fastq2=str(tmp_path / "nope.fastq"))._merge_reads(1, False)
    with pytest.raises(ValueError):
        SeqSamplePairedNotInterleaved(fastq=r1, tempdir=str(tmp_path / "d"), fastq2=None)._merge_reads(1, False)


def test_full_scale_property(gpu_ctx):
    """1 M pairs (BASELINE configs[4] sample scale is 781 250 pairs): size-independent properties instead of the oracle --
    every merged read of an error-free pair IS its fragment, merged lengths and the packed layout are consistent, and a
    second run gives the same bytes."""
    import synth
    n = 1_000_000
    frag, foff_, fs, fq, fo, rs, rq, ro = synth.make_pair_config(77, n, frag_len=(260, 480), err_scale=0.0, n_rate=0.0)
    ml, why, idx, ooff, gseq, gqual = gpu_ctx.merge_pairs(fs, fq, fo, rs, rq, ro)
    flen = np.diff(foff_)
    assert np.all(ml[idx] == flen[idx]) and np.array_equal(np.diff(ooff), ml[idx])
    assert len(idx) > 0.9 * n
    # nothing but the expected-error filter and the rare random second hit ("repeat", ~2e-5 of random fragments) fails
    assert set(np.unique(why).tolist()) <= {0, 1, 8} and int((why == 1).sum()) < n // 1000
    src = np.repeat(foff_[:-1][idx] - ooff[:-1], ml[idx]) + np.arange(int(ooff[-1]), dtype=np.int64)
    assert np.array_equal(gseq, frag[src])
    ml2, why2, idx2, ooff2, gseq2, gqual2 = gpu_ctx.merge_pairs(fs, fq, fo, rs, rq, ro)
    assert np.array_equal(ml, ml2) and zlib.crc32(gqual.tobytes()) == zlib.crc32(gqual2.tobytes())
    st = gpu_ctx.merge_stats()
    print("merge_kernel: %.2f ms for %d pairs (%.1f M pairs/s, %.1f GB/s in+out)" % (
        st.ms_kernel, n, n / st.ms_kernel / 1e3, (st.bytes_in + st.bytes_out) / st.ms_kernel / 1e6))


def _oracle_pipeline(oracle, r1, r2, stagger=False):
    """merge -> derep -> search (M.hmm 3_/4_) -> ItsPosition with the CPU oracle; returns what the trim needs."""
    from conftest import HMM_DIR
    b1, b2, fs, fq, fo, rs, rq, ro = _load(r1, r2)
    oml, _, oseq, oqual = oracle.merge_pairs(fs, fq, fo, rs, rq, ro, oracle.merge_params(allow_stagger=stagger))
    midx = np.flatnonzero(oml > 0)
    moff = np.zeros(len(midx) + 1, np.int64)
    moff[1:] = np.cumsum(oml[midx])
    slot = (fo[:-1] + ro[:-1])[midx]
    src = np.repeat(slot - moff[:-1], oml[midx]) + np.arange(int(moff[-1]), dtype=np.int64)
    mseq, mqual = oseq[src], oqual[src]
    rep, _, _ = oracle.derep(mseq, moff)
    uidx = np.flatnonzero(rep == np.arange(len(midx)))
    parts = [mseq[moff[i]:moff[i + 1]] for i in uidx]
    uoff = np.zeros(len(uidx) + 1, np.int64)
    uoff[1:] = np.cumsum([len(p) for p in parts])
    db = oracle.ProfileDB([os.path.join(HMM_DIR, "M.hmm")], ["3_", "4_"])
    rows, _, _ = db.search(oracle.digitize(np.concatenate(parts).tobytes()), uoff)
    side = np.array([0 if n.startswith("3_") else 1 for n in db.names], np.int8)
    pos = oracle.itspos(rows, side, np.diff(uoff).astype(np.int32))
    n = len(midx)
    s_r = np.full(n, -1, np.int32); e_r = s_r.copy(); t_r = s_r.copy()
    s_r[uidx] = pos["start"]; e_r[uidx] = pos["stop"]; t_r[uidx] = pos["tlen"]
    return b1, b2, fo, ro, midx, moff, mseq, mqual, rep, s_r, e_r, t_r


def test_cli_paired_end_to_end(tmp_path, oracle):
    """BASELINE configs[0]: `itsxpress --fastq R1 --fastq2 R2 --region ITS2` on the bundled pair sample (reference
    test_main_paired, tests/test_main_pytest.py:228-254; Metazoa profiles, F.hmm is missing from the mount), every
    stage on the GPU -- merged output and unmerged (--outfile2) output byte-identical to the oracle pipeline."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    r1n, r2n = "4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq"
    b1, b2, fo, ro, midx, moff, mseq, mqual, rep, s_r, e_r, t_r = _oracle_pipeline(oracle, r1n, r2n, stagger=True)      # the CLI's default, main.py:120-123
    # merged output: records carry R1's title
    out = str(tmp_path / "merged.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(TD, r1n), "--fastq2", os.path.join(TD, r2n),
                                             "--outfile", out, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l1.txt")]))
    keep, lo, hi = oracle.trim_bounds(moff, rep, s_r, e_r, t_r, mode=0)
    ki = np.flatnonzero(keep)
    assert len(ki) > 150
    want = []
    for k in ki:
        a, b = int(moff[k] + lo[k]), int(moff[k] + hi[k])
        want.append("@%s\n%s\n+\n%s\n" % (b1.title(int(midx[k])), mseq[a:b].tobytes().decode(), mqual[a:b].tobytes().decode()))
    assert open(out).read() == "".join(want)
    # unmerged output: R1[start:stop], R2[tlen-stop:tlen-start] of the pairs whose merged read was kept
    o1, o2 = str(tmp_path / "r1.fastq.gz"), str(tmp_path / "r2.fastq.gz")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(TD, r1n), "--fastq2", os.path.join(TD, r2n),
                                             "--outfile", o1, "--outfile2", o2, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l2.txt")]))
    for path, batch, off_all, mode in ((o1, b1, fo, 2), (o2, b2, ro, 1)):
        ln = np.diff(off_all)[midx]
        off_m = np.zeros(len(midx) + 1, np.int64)
        off_m[1:] = np.cumsum(ln)
        keep, lo, hi = oracle.trim_bounds(off_m, rep, s_r, e_r, t_r, mode=mode, off_r2=off_m)
        ki = np.flatnonzero(keep)
        assert fq._open_bytes(path) == fq.format_records(batch, midx[ki], lo[ki], hi[ki])


def test_q2_trim_pair_actions(tmp_path):
    """BASELINE configs[4] at fixture scale: the QIIME 2 actions trim-pair and trim-pair-output-unmerged on the
    reference's own paired per-sample directory (tests/test_data/paired/.../data; reference q2_itsxpress.py:156-230,
    tests/test_q2_itsxpress.py), merge included on the GPU.  Plugin and CLI both allow staggered merges by default
    (q2_itsxpress.py:161,199; main.py:120-123); output bytes equal the CLI's."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    from itsxpress_b200 import q2_itsxpress as q2
    src = os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data")
    n1, n2 = "4774-1-MSITS3_0_L001_R1_001.fastq.gz", "4774-1-MSITS3_1_L001_R2_001.fastq.gz"
    res = q2.trim_pair_output_unmerged(q2.PerSampleDir(src), region="ITS2", taxa="M")
    o1, o2 = os.path.join(str(res), n1), os.path.join(str(res), n2)
    b1, b2 = fq.read_fastq(o1), fq.read_fastq(o2)
    assert b1.n == b2.n and b1.n > 150
    assert [b1.title(i).split()[0] for i in range(b1.n)] == [b2.title(i).split()[0] for i in range(b2.n)]
    c1, c2 = str(tmp_path / "c1.fastq"), str(tmp_path / "c2.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(src, n1), "--fastq2", os.path.join(src, n2),
                                             "--outfile", c1, "--outfile2", c2, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l.txt")]))
    assert fq._open_bytes(o1) == open(c1, "rb").read() and fq._open_bytes(o2) == open(c2, "rb").read()
    man = open(os.path.join(str(res), "MANIFEST")).read().splitlines()
    assert man == ["sample-id,filename,direction", "4774-1-MSITS3,%s,forward" % n1, "4774-1-MSITS3,%s,reverse" % n2]
    # merged output
    res = q2.trim_pair(q2.PerSampleDir(src), region="ITS2", taxa="M")
    m = fq.read_fastq(os.path.join(str(res), n1))
    assert m.n == b1.n and not os.path.exists(os.path.join(str(res), n2))
    # reversed primers: the mates swap roles before the merge, so the merged reads of THIS sample come out as the
    # reverse complement and no profile matches (hmmsearch scans the given strand only): an empty, valid output
    res = q2.trim_pair(q2.PerSampleDir(src), region="ITS2", taxa="M", reversed_primers=True)
    assert fq.read_fastq(os.path.join(str(res), n1)).n == 0


def test_q2_main_sharded_single_rank_equals_action(tmp_path):
    """q2_itsxpress.main_sharded (samples dealt to ranks; here one rank) writes the same files as the plain action."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import q2_itsxpress as q2
    src = os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data")
    n1, n2 = "4774-1-MSITS3_0_L001_R1_001.fastq.gz", "4774-1-MSITS3_1_L001_R2_001.fastq.gz"
    ref = q2.trim_pair_output_unmerged(q2.PerSampleDir(src), region="ITS2", taxa="M")
    res, mine = q2.main_sharded(q2.PerSampleDir(src), str(tmp_path / "out"), region="ITS2", taxa="M", rank=0, world=1)
    assert mine == ["4774-1-MSITS3"]
    for n in (n1, n2):
        assert fq._open_bytes(os.path.join(str(res), n)) == fq._open_bytes(os.path.join(str(ref), n))
    assert open(os.path.join(str(res), "MANIFEST")).read() == open(os.path.join(str(ref), "MANIFEST")).read()


def _make_artifact(root, sizes, seed=3):
    """A paired per-sample directory of len(sizes) samples drawn from the reference's 250 fixture pairs (overlapping
    subsets, so the same sequences occur in several samples), plus one sample with zero reads."""
    import gzip
    from itsxpress_b200 import fastq as fq
    src = os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data")
    b1 = fq.read_fastq(os.path.join(src, "4774-1-MSITS3_0_L001_R1_001.fastq.gz"))
    b2 = fq.read_fastq(os.path.join(src, "4774-1-MSITS3_1_L001_R2_001.fastq.gz"))
    rng = np.random.default_rng(seed)
    os.makedirs(root)
    lines = ["sample-id,filename,direction"]
    for k, n in enumerate(sizes):
        pick = np.sort(rng.choice(b1.n, size=n, replace=False)) if n else np.zeros(0, np.int64)
        for b, tag, d in ((b1, "R1", "forward"), (b2, "R2", "reverse")):
            fn = "S%d_%d_L001_%s_001.fastq.gz" % (k, k, tag)
            z = np.zeros(len(pick), np.int32)
            with gzip.open(os.path.join(root, fn), "wb", compresslevel=1) as f:
                f.write(fq.format_records(b, pick, z, b.s_len[pick].astype(np.int32)))
            lines.append("S%d,%s,%s" % (k, fn, d))
    with open(os.path.join(root, "MANIFEST"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return root


@pytest.mark.parametrize("action", ["pair-unmerged", "pair"])
def test_q2_batched_samples_equal_the_per_sample_loop(tmp_path, action, monkeypatch):
    """all samples of an artifact in ONE device pass (segmented derep: the sample id is part of the key and
    of the class test; reported hits per (sample, profile)) write the same bytes as the reference's sequential loop over
    samples (q2_itsxpress.py:273-333) -- five samples of 17..230 pairs sharing sequences, one of them empty."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import q2_itsxpress as q2
    art = _make_artifact(str(tmp_path / "in"), [230, 17, 120, 0, 64])
    fn = q2.trim_pair_output_unmerged if action == "pair-unmerged" else q2.trim_pair
    monkeypatch.setattr(q2, "BATCH_READS", 0)
    ref = fn(q2.PerSampleDir(art), region="ITS2", taxa="M")
    monkeypatch.setattr(q2, "BATCH_READS", 4_000_000)
    one = fn(q2.PerSampleDir(art), region="ITS2", taxa="M")
    monkeypatch.setattr(q2, "BATCH_READS", 300)             # several batches
    few = fn(q2.PerSampleDir(art), region="ITS2", taxa="M")
    names = sorted(f for f in os.listdir(str(ref)) if f.endswith(".gz"))
    assert len(names) == (10 if action == "pair-unmerged" else 5)
    total = 0
    for got in (one, few):
        assert sorted(f for f in os.listdir(str(got)) if f.endswith(".gz")) == names
        for n in names:
            want = fq._open_bytes(os.path.join(str(ref), n))
            assert fq._open_bytes(os.path.join(str(got), n)) == want, n
            total += len(want)
        assert open(os.path.join(str(got), "MANIFEST")).read() == open(os.path.join(str(ref), "MANIFEST")).read()
    assert total > 100_000
