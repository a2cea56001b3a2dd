"""GPU tests through the reference-facing classes (SeqSample / ItsPosition / Dedup / CLI): the drop-in
boundary on a real B200, checked against the reference's golden files and against the CPU oracle."""
import argparse
import gzip
import os

import numpy as np
import pytest

from conftest import HMM_DIR, ROOT, TD

pytestmark = pytest.mark.gpu

UC = os.path.join(TD, "ex_tmpdir", "uc.txt")
SEQ = os.path.join(TD, "ex_tmpdir", "seq.fq.gz")
REP = os.path.join(TD, "ex_tmpdir", "rep.fa")
GOLD_SINGLE = os.path.join(TD, "singleOut", "75aea4f5-f10e-421e-91d2-feda9fe7b2e1", "data",
                           "4774-1-MSITS3_0_L001_R1_001.fastq.gz")


class GoldenPos:
    def __init__(self):
        self.tab = {}
        with open(os.path.join(ROOT, "tests", "golden", "c1_positions.tsv")) as f:
            for line in f:
                if not line.startswith("#"):
                    k, a, b, c = line.split("\t")
                    self.tab[k] = (int(a), int(b), int(c))

    def get_position(self, k):
        if k not in self.tab:
            raise KeyError
        return self.tab[k]


def test_deduplicate_writes_vsearch_files(tmp_path):
    """SeqSampleNotPaired.deduplicate on the GPU == the real vsearch output kept in the reference's fixture."""
    from itsxpress_b200.SeqSample import SeqSampleNotPaired
    s = SeqSampleNotPaired(SEQ, str(tmp_path))
    s.deduplicate(threads=1)
    assert s.uc_file == str(tmp_path / "uc.txt") and s.rep_file == str(tmp_path / "rep.fa")
    assert open(s.uc_file, "rb").read() == open(UC, "rb").read()
    assert open(s.rep_file, "rb").read() == open(REP, "rb").read()


@pytest.mark.parametrize("mode", ["plain", "gz", "zst"])
def test_create_trimmed_seqs_golden(tmp_path, mode):
    """reference test_dedup_create_trimmed_seqs{,_gzipped,_zst} (:68-161): 226 records, 42 637 bases -- and here
    the exact bytes of the QIIME 2 single-end golden."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200.SeqSample import Dedup
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    out = str(tmp_path / {"plain": "o.fastq", "gz": "o.fastq.gz", "zst": "o.fastq.zst"}[mode])
    d.create_trimmed_seqs(out, gzipped=mode == "gz", zstd_file=mode == "zst", itspos=GoldenPos(), wri_file=True,
                          tempdir=str(tmp_path))
    b = fq.read_fastq(out)
    assert b.n == 226 and int(b.s_len.sum()) == 42637
    assert fq._open_bytes(out) == gzip.open(GOLD_SINGLE, "rb").read()


def test_create_paired_trimmed_seqs_golden(tmp_path):
    """reference test_create_paired_trimmed_seqs (:378-397): filecmp with t2_r1.fq / t2_r2.fq."""
    from itsxpress_b200.SeqSample import Dedup
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ, fastq=os.path.join(TD, "4774-1-MSITS3_R1.fastq"),
              fastq2=os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    o1, o2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    d.create_paired_trimmed_seqs(o1, o2, gzipped=False, zstd_file=False, itspos=GoldenPos(), wri_file=True)
    assert open(o1, "rb").read() == open(os.path.join(TD, "t2_r1.fq"), "rb").read()
    assert open(o2, "rb").read() == open(os.path.join(TD, "t2_r2.fq"), "rb").read()


def test_trim_ccs_bulk(tmp_path):
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200.SeqSample import Dedup
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    out = str(tmp_path / "ccs.fastq")
    d.create_trimmed_seqs(out, False, False, GoldenPos(), True, str(tmp_path), trim_ccs=True)
    b = fq.read_fastq(out)
    g = fq.parse_bytes(gzip.open(GOLD_SINGLE, "rb").read())
    assert b.n == 226
    for i in (0, 17, 225):
        assert b.seq(i) == "GACAGGTACAAGAAGGA" + g.seq(i) + "TTAACCCAGTCTCCAGT"
        assert b.qual(i) == "~" * 17 + g.qual(i) + "~" * 17


def test_pipeline_objects_device_vs_files(tmp_path, oracle):
    """deduplicate -> _search -> ItsPosition -> Dedup -> create_trimmed_seqs with everything device-backed,
    against (a) the same objects rebuilt from the files alone and (b) the CPU oracle."""
    from itsxpress_b200 import SeqSample as S
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200.main import create_runtime_hmm
    s = S.SeqSampleNotPaired(SEQ, str(tmp_path))
    s.deduplicate(threads="1")
    hmm = create_runtime_hmm("Metazoa", "ITS2", str(tmp_path))
    s._search(hmmfile=hmm, threads="1")
    its = S.ItsPosition(s.dom_file, "ITS2")
    assert its._dev is not None                       # arrays came from the device
    d = S.Dedup(s.uc_file, s.rep_file, s.seq_file)
    out_dev = str(tmp_path / "dev.fastq")
    d.create_trimmed_seqs(out_dev, False, False, its, True, str(tmp_path))

    # (a) the same from files only (as the reference's tests build the objects)
    S.reset_sessions()
    its_f = S.ItsPosition(s.dom_file, "ITS2")
    assert its_f._dev is None
    d_f = S.Dedup(s.uc_file, s.rep_file, s.seq_file)
    assert len(d_f.matchdict) == 227
    out_file = str(tmp_path / "file.fastq")
    d_f.create_trimmed_seqs(out_file, False, False, its_f, True, str(tmp_path))
    assert open(out_dev, "rb").read() == open(out_file, "rb").read()
    # device positions == text-parsed positions for every representative
    for k, sid in enumerate(s._session.seq_ids):
        dev = (int(its._dev["start"][k]), int(its._dev["stop"][k]), int(its._dev["tlen"][k]))
        try:
            a, b, c = its_f.get_position(sid)
        except KeyError:
            a = b = c = None
        assert dev == tuple(-1 if v is None else v for v in (a, b, c)), sid

    # (b) oracle
    b = fq.read_fastq(SEQ)
    seq, off = b.seq_concat()
    rep, _, _ = oracle.derep(seq, off)
    idx = np.flatnonzero(rep == np.arange(b.n))
    parts = [seq[off[i]:off[i + 1]] for i in idx]
    uoff = np.zeros(len(idx) + 1, np.int64)
    uoff[1:] = np.cumsum([len(p) for p in parts])
    db = oracle.ProfileDB([os.path.join(HMM_DIR, "M.hmm")], ["3_", "4_"])
    rows, nrep, st = db.search(oracle.digitize(np.concatenate(parts).tobytes()), uoff)
    side = np.array([0 if n.startswith("3_") else 1 for n in db.names], np.int8)
    pos = oracle.itspos(rows, side, np.diff(uoff).astype(np.int32))
    s_r = np.full(b.n, -1, np.int32); e_r = s_r.copy(); t_r = s_r.copy()
    s_r[idx] = pos["start"]; e_r[idx] = pos["stop"]; t_r[idx] = pos["tlen"]
    keep, lo, hi = oracle.trim_bounds(off, rep, s_r, e_r, t_r, mode=0)
    ki = np.flatnonzero(keep)
    assert len(ki) > 100
    assert open(out_dev, "rb").read() == fq.format_records(b, ki, lo[ki], hi[ki])


def test_cli_end_to_end(tmp_path, oracle, caplog):
    """`itsxpress --fastq merged.fastq --single_end --region ITS2 --taxa Metazoa` (reference test_main_merged
    :288-314 with the taxon that is present in the mount); --keeptemp leaves the reference's temp files."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    out = str(tmp_path / "out.fastq.gz")
    args = cli.myparser().parse_args(["--fastq", os.path.join(TD, "4774-1-MSITS3_merged.fastq"), "--single_end",
                                      "--outfile", out, "--region", "ITS2", "--taxa", "Metazoa", "--keeptemp",
                                      "--tempdir", str(tmp_path / "tmp"), "--log", str(tmp_path / "log.txt")])
    import logging
    caplog.set_level(logging.INFO)
    cli.main(args=args)
    b = fq.read_fastq(out)
    assert b.n > 100
    tdirs = os.listdir(str(tmp_path / "tmp"))
    assert len(tdirs) == 1 and tdirs[0].startswith("itsxpress_")
    kept = set(os.listdir(os.path.join(str(tmp_path / "tmp"), tdirs[0])))
    assert {"uc.txt", "rep.fa", "domtbl.txt", "runtime_selected.hmm"} <= kept
    # (pytest owns the root logger, so basicConfig(filename=...) is a no-op here exactly as it is upstream)
    assert "Total number of reads in file" in caplog.text and "ITSxpress ran in" in caplog.text


def test_cli_failure_exit_code(tmp_path):
    from itsxpress_b200 import main as cli
    args = cli.myparser().parse_args(["--fastq", os.path.join(TD, "broken.fastq"), "--single_end", "--outfile",
                                      str(tmp_path / "o.fq"), "--region", "ITS2", "--taxa", "Metazoa",
                                      "--log", str(tmp_path / "log.txt")])
    with pytest.raises(SystemExit) as e:
        cli.main(args=args)
    assert e.value.code == 1


def test_search_without_profiles_fails_like_hmmsearch(tmp_path):
    """`--taxa Fungi` with F.hmm absent (or the QIIME 2 letter R -> 'Rhizaria', which misses the
    ' Rhizaria' key) yields a runtime HMM file without profiles; upstream hmmsearch exits non-zero on it
    (SeqSample.py:211-225 -> CalledProcessError -> CLI exit 1).  Never an empty output with status 0."""
    import subprocess
    from itsxpress_b200 import main as cli
    from itsxpress_b200.SeqSample import SeqSampleNotPaired
    empty = tmp_path / "runtime_selected.hmm"
    empty.write_text("")
    s = SeqSampleNotPaired(os.path.join(TD, "ex_tmpdir", "seq.fq.gz"), str(tmp_path))
    s.deduplicate(threads=1)
    with pytest.raises(subprocess.CalledProcessError):
        s._search(str(empty), threads=1)
    if not os.path.exists(os.path.join(HMM_DIR, "F.hmm")):
        args = cli.myparser().parse_args(["--fastq", os.path.join(TD, "4774-1-MSITS3_merged.fastq"), "--single_end",
                                          "--outfile", str(tmp_path / "o.fq"), "--region", "ITS2", "--taxa", "Fungi",
                                          "--log", str(tmp_path / "log.txt")])
        with pytest.raises(SystemExit) as e:
            cli.main(args=args)
        assert e.value.code == 1


def test_q2_trim_single(tmp_path):
    """q2 action trim_single on a single-end per-sample directory (layout of the reference's
    tests/test_data/singleIn artifact; the sample here is the MERGED fixture so that both ITS2 boundaries are inside
    the reads) with the taxon present in the mount; output is a Casava-1.8 directory whose file carries the input's
    name; bytes equal the CLI's output for the same sample."""
    import shutil
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    from itsxpress_b200 import q2_itsxpress as q2
    src = tmp_path / "in"
    src.mkdir()
    name = "4774-1-MSITS3_0_L001_R1_001.fastq.gz"
    shutil.copy(SEQ, str(src / name))
    (src / "MANIFEST").write_text("sample-id,filename,direction\n4774-1-MSITS3,%s,forward\n" % name)
    (src / "metadata.yml").write_text("{phred-offset: 33}\n")
    inp = q2.PerSampleDir(str(src))
    frame = inp.manifest.view(None)
    assert list(frame.columns) == ["forward"] and list(frame.index) == ["4774-1-MSITS3"]
    res = q2.trim_single(inp, region="ITS2", taxa="M")
    out = os.path.join(str(res), name)
    assert os.path.exists(out)
    b = fq.read_fastq(out)
    assert 100 < b.n <= 227
    man = open(os.path.join(str(res), "MANIFEST")).read().splitlines()
    assert man == ["sample-id,filename,direction", "4774-1-MSITS3,%s,forward" % name]
    ref = str(tmp_path / "cli.fastq")
    cli.main(args=cli.myparser().parse_args(["--fastq", SEQ, "--single_end", "--outfile", ref, "--region", "ITS2",
                                             "--taxa", "Metazoa", "--log", str(tmp_path / "l.txt")]))
    assert fq._open_bytes(out) == open(ref, "rb").read()
    # the reference's own single-end artifact holds raw R1 reads: no read spans both ITS2 boundaries
    r1 = q2.trim_single(q2.PerSampleDir(os.path.join(TD, "singleIn", "cfd0e65b-05fb-4329-9618-15ecd0aec9b3", "data")),
                        region="ITS2", taxa="M")
    assert fq.read_fastq(os.path.join(str(r1), name)).n < 20
    with pytest.raises(KeyError):
        q2._taxa_prefix_to_taxa("Z")


def test_pipeline_without_materialised_temp_files(tmp_path):
    """Large-sample mode: uc.txt / rep.fa / domtbl.txt are placeholders and ItsPosition / Dedup run from the live
    session's device arrays; the output must equal the fully materialised run."""
    from itsxpress_b200 import SeqSample as S
    from itsxpress_b200.main import create_runtime_hmm
    outs = []
    for mat in (True, False):
        d = tmp_path / ("m%d" % mat)
        d.mkdir()
        s = S.SeqSampleNotPaired(SEQ, str(d))
        s.materialize = mat
        s.deduplicate(threads=1)
        s._search(hmmfile=create_runtime_hmm("Metazoa", "ITS2", str(d)), threads=1)
        its = S.ItsPosition(s.dom_file, "ITS2")
        dd = S.Dedup(s.uc_file, s.rep_file, s.seq_file)
        out = str(d / "o.fastq")
        dd.create_trimmed_seqs(out, False, False, its, True, str(d))
        outs.append(open(out, "rb").read())
        assert (os.path.getsize(s.dom_file) > 10000) == mat
    assert outs[0] == outs[1] and len(outs[0]) > 10000


@pytest.mark.parametrize("suffix", ["", ".gz"])
def test_cli_streamed_equals_whole_file(tmp_path, monkeypatch, suffix):
    """the chunked path (ITSX_STREAM=1: itsx_reads_begin / append / end on the way in, two passes over the
    input, itsx_trim_gather_range + an appending writer on the way out) writes the same bytes as the whole-file path --
    the merged fixture through the CLI with chunks of ~20 records, plain and gzip output."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    from itsxpress_b200 import SeqSample
    src = os.path.join(TD, "4774-1-MSITS3_merged.fastq")
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ITSX_STREAM", mode)
        monkeypatch.setattr(fq, "STREAM_CHUNK_BYTES", 16_000)
        out = str(tmp_path / ("o%s.fastq%s" % (mode, suffix)))
        cli.main(args=cli.myparser().parse_args(["--fastq", src, "--single_end", "--outfile", out, "--region", "ITS2",
                                                 "--taxa", "Metazoa", "--log", str(tmp_path / "l.txt")]))
        outs[mode] = fq._open_bytes(out)
    assert outs["0"] == outs["1"] and outs["0"].count(b"\n") // 4 > 150


def test_cli_compact_row_mode_equals_full_table(tmp_path, monkeypatch):
    """Above 200 000 uniques the search keeps only undecided rows (itsx_search_params.keep_rows = 2), which needs the
    boundary sides DURING the search: _search reads them off the runtime HMM's two name prefixes and ItsPosition then
    takes the positions without a second selection pass.  Forced here on the fixture (ITSX_KEEP_ROWS=2, temp files not
    materialised): same output bytes as the default path."""
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    src = os.path.join(TD, "4774-1-MSITS3_merged.fastq")
    outs = {}
    from itsxpress_b200 import SeqSample
    monkeypatch.setattr(SeqSample, "TEMP_FILE_POLICY", "never")
    for mode in ("1", "2"):
        monkeypatch.setenv("ITSX_KEEP_ROWS", mode)
        out = str(tmp_path / ("o%s.fastq" % mode))
        cli.main(args=cli.myparser().parse_args(["--fastq", src, "--single_end", "--outfile", out, "--region", "ITS2",
                                                 "--taxa", "Metazoa", "--log", str(tmp_path / "l.txt")]))
        outs[mode] = open(out, "rb").read()
    assert outs["1"] == outs["2"] and outs["1"].count(b"\n") // 4 > 150
