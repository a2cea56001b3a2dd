"""The command line end to end on the CPU: the product's host code (main.py, SeqSample / ItsPosition / Dedup, the FASTQ
readers and writers incl. the native gzip reader, temp-file policy, streaming) with the CPU oracle standing in for the
device (tests/oracle_context.py).  What the reference's own CLI tests do with vsearch + hmmsearch underneath
(tests/test_main_pytest.py:205-350), here with the oracle underneath; the GPU suite runs the same flows on the real library.

Checked: the trimmed records equal an independent composition of the oracle's stages; plain / .gz / .zst inputs and outputs
agree; the streamed path (chunks of a few KB) equals the whole-file path; with --keeptemp the files the reference's stages
exchange are the real vsearch output of the fixture, byte for byte (uc.txt, rep.fa); the paired command line (merged and
unmerged output, BASELINE configs[0]) and the QIIME 2 paired actions on the reference's own per-sample directory."""
import glob
import gzip
import os

import numpy as np
import pytest

from conftest import HMM_DIR, TD

from itsxpress_b200 import SeqSample
from itsxpress_b200 import _zstd
from itsxpress_b200 import fastq as fq
from itsxpress_b200 import main as cli
from oracle_context import OracleContext

MERGED = os.path.join(TD, "4774-1-MSITS3_merged.fastq")          # == ex_tmpdir/seq.fq.gz, 227 reads


@pytest.fixture()
def on_oracle(oracle, monkeypatch):
    ctx = OracleContext(oracle)
    monkeypatch.setattr(SeqSample, "get_context", lambda: ctx)
    SeqSample.reset_sessions()
    yield ctx
    SeqSample.reset_sessions()


def expected_single(oracle, path):
    """derep -> search of the uniques -> ItsPosition -> trim bounds -> records, straight from the oracle's stages"""
    b = fq.read_fastq(path)
    seq, off = b.seq_concat()
    rep, _, nu = oracle.derep(seq, off)
    first = np.flatnonzero(rep == np.arange(b.n))
    uid = np.searchsorted(first, rep).astype(np.int32)
    parts = [seq[off[i]:off[i + 1]] for i in first]
    uoff = np.zeros(nu + 1, np.int64)
    uoff[1:] = np.cumsum([len(p) for p in parts])
    db = oracle.ProfileDB([os.path.join(HMM_DIR, "M.hmm")], ["3_", "4_"])
    side = np.array([0 if n.startswith("3_") else 1 for n in db.names], np.int8)
    rows, _, _ = db.search(oracle.digitize(np.concatenate(parts).tobytes()), uoff, oracle.default_params(0, 1))
    pos = oracle.itspos(rows[rows["is_reported"] != 0], side, np.diff(uoff).astype(np.int32))
    keep, lo, hi = oracle.trim_bounds(off, uid, pos["start"], pos["stop"], pos["tlen"], mode=0)
    ki = np.flatnonzero(keep)
    return fq.format_records(b, ki, lo[ki], hi[ki]), len(ki)


def run_cli(tmp, fastq, outname, *extra):
    out = os.path.join(str(tmp), outname)
    argv = ["--fastq", fastq, "--single_end", "--outfile", out, "--region", "ITS2", "--taxa", "Metazoa",
            "--log", os.path.join(str(tmp), "log.txt"), "--tempdir", str(tmp)] + list(extra)
    cli.main(args=cli.myparser().parse_args(argv))
    return out


def test_single_end_cli_equals_the_oracles_stages_for_every_container(oracle, on_oracle, tmp_path):
    want, nkept = expected_single(oracle, MERGED)
    assert 100 < nkept <= 227
    raw = open(MERGED, "rb").read()
    inputs = {"plain": MERGED, "gz": str(tmp_path / "in.fastq.gz"), "zst": str(tmp_path / "in.fastq.zst")}
    open(inputs["gz"], "wb").write(gzip.compress(raw, 6))
    open(inputs["zst"], "wb").write(_zstd.compress(raw))
    for kind, path in inputs.items():
        for outname in ("o_%s.fastq" % kind, "o_%s.fastq.gz" % kind, "o_%s.fastq.zst" % kind):
            out = run_cli(tmp_path, path, outname)
            assert fq._open_bytes(out) == want, (kind, outname)
    assert gzip.open(str(tmp_path / "o_gz.fastq.gz"), "rb").read() == want          # any gzip reader
    assert "derep" in on_oracle.calls and "search" in on_oracle.calls and "reads_begin" not in on_oracle.calls
    assert fq.cached_count(str(tmp_path / "o_gz.fastq.gz")) == nkept            # the closing read counts need no second pass


@pytest.mark.parametrize("suffix", ["", ".gz", ".zst"])
def test_streamed_cli_equals_the_whole_file_path(oracle, on_oracle, tmp_path, monkeypatch, suffix):
    want, _ = expected_single(oracle, MERGED)
    raw = open(MERGED, "rb").read()
    src = str(tmp_path / ("in.fastq" + suffix))
    open(src, "wb").write(raw if not suffix else gzip.compress(raw) if suffix == ".gz" else _zstd.compress(raw))
    monkeypatch.setenv("ITSX_STREAM", "1")
    monkeypatch.setattr(fq, "STREAM_CHUNK_BYTES", 16_000)
    out = run_cli(tmp_path, src, "streamed.fastq" + suffix)
    assert fq._open_bytes(out) == want
    calls = on_oracle.calls
    # (the zstd decoder hands out whole 1 MiB buffers: this small file is one chunk there)
    assert calls.count("reads_append") > (5 if suffix != ".zst" else 0)
    assert calls.count("trim_gather_range") == calls.count("reads_append")
    assert "derep" not in calls and "trim_gather" not in calls
    monkeypatch.setenv("ITSX_STREAM", "0")
    on_oracle.calls.clear()
    out = run_cli(tmp_path, src, "whole.fastq" + suffix)
    assert fq._open_bytes(out) == want and "reads_append" not in on_oracle.calls


def test_keeptemp_leaves_the_files_vsearch_writes(oracle, on_oracle, tmp_path):
    """`uc.txt` and `rep.fa` of the fixture are a REAL `vsearch --fastx_uniques` output for this input: the command line
    must leave exactly those bytes behind (the reference's stages hand each other these files, SeqSample.py:93-131)."""
    run_cli(tmp_path, os.path.join(TD, "ex_tmpdir", "seq.fq.gz"), "kept.fastq", "--keeptemp")
    tmp = [d for d in glob.glob(os.path.join(str(tmp_path), "itsxpress_*")) if os.path.isdir(d)]
    assert len(tmp) == 1
    for name in ("uc.txt", "rep.fa"):
        assert open(os.path.join(tmp[0], name), "rb").read() == open(os.path.join(TD, "ex_tmpdir", name), "rb").read(), name
    dom = open(os.path.join(tmp[0], "domtbl.txt")).read().splitlines()
    rows = [ln for ln in dom if ln and not ln.startswith("#")]
    assert len(rows) > 200 and all(len(ln.split()) >= 23 for ln in rows)
    # the table on disk is what a fresh ItsPosition parses, with the same positions as the device hand-off
    SeqSample.reset_sessions()
    parsed = SeqSample.ItsPosition(os.path.join(tmp[0], "domtbl.txt"), "ITS2")
    want, _ = expected_single(oracle, MERGED)
    dd = SeqSample.Dedup(uc_file=os.path.join(tmp[0], "uc.txt"), rep_file=os.path.join(tmp[0], "rep.fa"),
                         seq_file=os.path.join(TD, "ex_tmpdir", "seq.fq.gz"))
    out = str(tmp_path / "from_files.fastq")
    dd.create_trimmed_seqs(out, gzipped=False, zstd_file=False, itspos=parsed, wri_file=True, tempdir=str(tmp_path))
    assert open(out, "rb").read() == want


def test_missing_profiles_and_broken_input_end_the_cli_with_status_1(on_oracle, tmp_path):
    with pytest.raises(SystemExit) as e:
        run_cli(tmp_path, os.path.join(TD, "broken.fastq"), "x.fastq")
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:                          # --taxa Fungi: F.hmm is not shipped
        out = os.path.join(str(tmp_path), "y.fastq")
        cli.main(args=cli.myparser().parse_args(["--fastq", MERGED, "--single_end", "--outfile", out, "--region", "ITS2",
                                                 "--taxa", "Fungi", "--log", os.path.join(str(tmp_path), "l.txt"),
                                                 "--tempdir", str(tmp_path)]))
    assert e.value.code == 1 and not os.path.exists(out)


def test_paired_cli_merged_and_unmerged_output(oracle, on_oracle, tmp_path):
    """BASELINE configs[0] (reference test_main_paired / test_main_paired_no_merge, tests/test_main_pytest.py:228-350) on the
    bundled pair sample: the expectations of tests/test_gpu_merge.py::test_cli_paired_end_to_end, host code on the oracle."""
    from test_gpu_merge import _oracle_pipeline
    r1n, r2n = "4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq"
    b1, b2, fo, ro, midx, moff, mseq, mqual, rep, s_r, e_r, t_r = _oracle_pipeline(oracle, r1n, r2n, stagger=True)
    out = str(tmp_path / "merged.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(TD, r1n), "--fastq2", os.path.join(TD, r2n),
                                             "--outfile", out, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l1.txt"), "--tempdir", str(tmp_path)]))
    keep, lo, hi = oracle.trim_bounds(moff, rep, s_r, e_r, t_r, mode=0)
    ki = np.flatnonzero(keep)
    assert len(ki) > 150
    want = []
    for k in ki:
        a, b = int(moff[k] + lo[k]), int(moff[k] + hi[k])
        want.append("@%s\n%s\n+\n%s\n" % (b1.title(int(midx[k])), mseq[a:b].tobytes().decode(), mqual[a:b].tobytes().decode()))
    assert open(out).read() == "".join(want)
    assert "merge_pairs" in on_oracle.calls
    for ext in (".fastq.gz", ".fastq"):
        o1, o2 = str(tmp_path / ("r1" + ext)), str(tmp_path / ("r2" + ext))
        gz = ".gz" if ext.endswith(".gz") else ""
        cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(TD, r1n + gz), "--fastq2", os.path.join(TD, r2n + gz),
                                                 "--outfile", o1, "--outfile2", o2, "--region", "ITS2", "--taxa", "Metazoa",
                                                 "--log", str(tmp_path / "l2.txt"), "--tempdir", str(tmp_path)]))
        for path, batch, off_all, mode in ((o1, b1, fo, 2), (o2, b2, ro, 1)):
            ln = np.diff(off_all)[midx]
            off_m = np.zeros(len(midx) + 1, np.int64)
            off_m[1:] = np.cumsum(ln)
            keep, lo, hi = oracle.trim_bounds(off_m, rep, s_r, e_r, t_r, mode=mode, off_r2=off_m)
            ki = np.flatnonzero(keep)
            assert fq._open_bytes(path) == fq.format_records(batch, midx[ki], lo[ki], hi[ki]), (ext, mode)


def test_q2_paired_actions_equal_the_cli(oracle, on_oracle, tmp_path, monkeypatch):
    """trim-pair-output-unmerged and trim-pair on the reference's paired per-sample directory (q2_itsxpress.py:156-230): the
    files and the MANIFEST the plugin would hand back, bytes equal to the command line's."""
    from itsxpress_b200 import q2_itsxpress as q2
    monkeypatch.setattr(q2, "BATCH_READS", 0)                 # the per-sample loop (the batched pass is GPU-suite territory)
    src = os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data")
    n1, n2 = "4774-1-MSITS3_0_L001_R1_001.fastq.gz", "4774-1-MSITS3_1_L001_R2_001.fastq.gz"
    res = q2.trim_pair_output_unmerged(q2.PerSampleDir(src), region="ITS2", taxa="M")
    o1, o2 = os.path.join(str(res), n1), os.path.join(str(res), n2)
    c1, c2 = str(tmp_path / "c1.fastq"), str(tmp_path / "c2.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(src, n1), "--fastq2", os.path.join(src, n2),
                                             "--outfile", c1, "--outfile2", c2, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l.txt"), "--tempdir", str(tmp_path)]))
    assert fq._open_bytes(o1) == open(c1, "rb").read() and fq._open_bytes(o2) == open(c2, "rb").read()
    assert fq.read_fastq(o1).n == fq.read_fastq(o2).n > 150
    man = open(os.path.join(str(res), "MANIFEST")).read().splitlines()
    assert man == ["sample-id,filename,direction", "4774-1-MSITS3,%s,forward" % n1, "4774-1-MSITS3,%s,reverse" % n2]
    res = q2.trim_pair(q2.PerSampleDir(src), region="ITS2", taxa="M")
    cm = str(tmp_path / "cm.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", os.path.join(src, n1), "--fastq2", os.path.join(src, n2),
                                             "--outfile", cm, "--region", "ITS2", "--taxa", "Metazoa",
                                             "--log", str(tmp_path / "l.txt"), "--tempdir", str(tmp_path)]))
    assert fq._open_bytes(os.path.join(str(res), n1)) == open(cm, "rb").read()


@pytest.mark.parametrize("action", ["pair-unmerged", "pair"])
def test_q2_batched_samples_equal_the_per_sample_loop(oracle, on_oracle, tmp_path, action, monkeypatch):
    """SURVEY 8(f3), host side: the samples of an artifact grouped into shared passes (q2_itsxpress._process_batch: reads of
    all samples concatenated, one derep + search, outputs split by sample) write the same files as the sequential loop
    (q2_itsxpress.py:273-333) -- the same artifact as tests/test_gpu_merge.py uses on the device."""
    from itsxpress_b200 import q2_itsxpress as q2
    from test_gpu_merge import _make_artifact
    art = _make_artifact(str(tmp_path / "in"), [60, 17, 40, 0, 25])
    fn = q2.trim_pair_output_unmerged if action == "pair-unmerged" else q2.trim_pair
    monkeypatch.setattr(q2, "BATCH_READS", 0)
    ref = fn(q2.PerSampleDir(art), region="ITS2", taxa="M")
    assert "set_samples" not in on_oracle.calls and "reads_upload" not in on_oracle.calls
    monkeypatch.setattr(q2, "BATCH_READS", 4_000_000)
    one = fn(q2.PerSampleDir(art), region="ITS2", taxa="M")
    assert on_oracle.calls.count("reads_upload") == 1
    monkeypatch.setattr(q2, "BATCH_READS", 90)               # several batches
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
    assert total > 20_000
