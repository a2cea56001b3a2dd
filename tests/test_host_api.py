"""CPU tests of the reference-facing host layer (itsxpress_b200/SeqSample.py, main.py, fastq.py) and of the
C-ABI's shape.  They mirror the reference's own unit tests (tests/test_main_pytest.py, cited per test);
nothing here launches a kernel."""
import ctypes
import ctypes as C
import gzip
import os
import re
import shutil
import tempfile

import numpy as np
import pytest

from conftest import HMM_DIR, ROOT, TD

from itsxpress_b200 import fastq as fq
from itsxpress_b200 import main as cli
from itsxpress_b200 import q2_itsxpress as q2
from itsxpress_b200.SeqSample import Dedup, ItsPosition


UC = os.path.join(TD, "ex_tmpdir", "uc.txt")
SEQ = os.path.join(TD, "ex_tmpdir", "seq.fq.gz")
REP = os.path.join(TD, "ex_tmpdir", "rep.fa")


# ---- C ABI -------------------------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "itsx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itsx_[A-Za-z0-9_]+)\s*\(", text)))


def test_abi_exports_every_declared_symbol():
    from itsxpress_b200 import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 38
    for sym in declared:
        assert hasattr(L, sym), sym
    assert set(_lib.SYMBOLS) == set(declared)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product refuses to run (ITSX_ENODEV) instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from itsxpress_b200 import _lib
    with pytest.raises(_lib.ItsxError) as e:
        _lib.Context(0)
    assert e.value.code == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "itsxpress_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") or f.endswith(".cu") or f.endswith(".cpp") or f.endswith(".h"):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


# ---- FASTQ reader / writer (Biopython semantics) --------------------------------------
def test_fastq_roundtrip_bytes():
    raw = gzip.open(SEQ, "rb").read()
    b = fq.parse_bytes(raw)
    assert b.n == 227
    assert fq.format_records(b, np.arange(b.n), np.zeros(b.n, np.int64), b.s_len) == raw
    assert "".join(r.format("fastq") for r in fq.iter_records(SEQ)).encode() == raw


def test_broken_fastq_raises_valueerror():
    # reference tests/test_main_pytest.py:20-29
    with pytest.raises(ValueError):
        fq.read_fastq(os.path.join(TD, "broken.fastq"))
    with pytest.raises(ValueError):
        cli._check_fastqs(os.path.join(TD, "broken.fastq"))


def test_check_fastqs_empty(tmp_path):
    # reference :403-409
    p = tmp_path / "empty.fastq"
    p.write_text("")
    cli._check_fastqs(str(p))


def test_check_fastqs_reads_only_the_head_of_compressed_files(tmp_path):
    from itsxpress_b200 import _zstd
    raw = gzip.open(SEQ, "rb").read()
    good, broken = str(tmp_path / "a.fastq.zst"), str(tmp_path / "b.fastq.zst")
    open(good, "wb").write(_zstd.compress(raw))
    open(broken, "wb").write(_zstd.compress(open(os.path.join(TD, "broken.fastq"), "rb").read()))
    cli._check_fastqs(good)
    cli._check_fastqs(SEQ)
    with pytest.raises(ValueError):
        cli._check_fastqs(broken)
    with pytest.raises(Exception):
        cli._check_fastqs(str(tmp_path / "missing.fastq.zst"))


def test_zstd_and_gzip_writers(tmp_path):
    raw = gzip.open(SEQ, "rb").read()
    fq.write_compressed(str(tmp_path / "a.gz"), raw, gzipped=True)
    fq.write_compressed(str(tmp_path / "a.zst"), raw, zstd_file=True)
    fq.write_compressed(str(tmp_path / "a.fq"), raw)
    assert gzip.open(str(tmp_path / "a.gz"), "rb").read() == raw
    assert fq.read_fastq(str(tmp_path / "a.zst")).n == 227
    assert open(str(tmp_path / "a.fq"), "rb").read() == raw


def test_zstd_decoder_flushes_held_back_blocks_and_rejects_truncation():
    """the streaming decoder may consume all input while a decoded block still waits for output space
    (multi-frame files, blocks that do not align to the 1 MiB output buffer); a cut-off frame raises like pyzstd."""
    from itsxpress_b200 import _zstd
    rng = np.random.default_rng(5)
    a = rng.integers(65, 70, (1 << 20) + 12345, dtype=np.uint8).tobytes()          # > 1 MiB, compressible
    b = rng.integers(65, 70, 700_001, dtype=np.uint8).tobytes()
    blob = _zstd.compress(a, level=1) + _zstd.compress(b, level=19)                  # two frames back to back
    assert _zstd.decompress(blob) == a + b
    with pytest.raises(ValueError):
        _zstd.decompress(blob[:-7])
    # large buffers: several frames compressed concurrently, decoded frame-parallel or by the streaming decoder alike
    big = bytes(rng.integers(65, 75, 3_000_000, dtype=np.uint8))
    old_frame = _zstd.FRAME_BYTES
    _zstd.FRAME_BYTES = 700_000
    try:
        blob = _zstd.compress(big, threads=4)
    finally:
        _zstd.FRAME_BYTES = old_frame
    src = C.create_string_buffer(blob, len(blob))
    assert len(_zstd._frames(_zstd._lib(), C.addressof(src), len(blob))) == 5
    assert _zstd.decompress(blob, threads=4) == big and _zstd.decompress(blob, threads=1) == big
    assert _zstd._decompress_streaming(blob) == big and _zstd.compress(big, threads=1) != blob
    with pytest.raises(ValueError):
        _zstd.decompress(blob[:len(blob) // 2], threads=4)


def test_fastq_cut_is_behind_the_last_line_whose_number_is_a_multiple_of_four():
    rng = np.random.default_rng(9)
    L = fq._native()

    def cut(b):
        a = np.frombuffer(b, np.uint8)
        return int(L.itsx_fastq_cut(fq._vp(a), a.size)) if a.size else 0

    def want(b):
        nl = [i for i, c in enumerate(b) if c == 10]
        return nl[len(nl) // 4 * 4 - 1] + 1 if len(nl) >= 4 else 0
    assert cut(b"") == 0 and cut(b"@a\nAC\n+\n") == 0 and cut(b"@a\nAC\n+\nII\n") == 11 and cut(b"@a\nAC\n+\nII\n@b") == 11
    for _ in range(200):
        n = int(rng.integers(1, 3000))
        b = bytes(rng.choice(np.frombuffer(b"ACGT@+I\n\n", np.uint8), n))      # newline-rich, '@' anywhere
        assert cut(b) == want(b), b
    big = bytes(rng.choice(np.frombuffer(b"ACGTIIIIIIIIIIIIIIIIIIIIIII\n", np.uint8), 40_000_000))    # several threads
    a = np.frombuffer(big, np.uint8)
    nl = np.flatnonzero(a == 10)
    assert cut(big) == int(nl[len(nl) // 4 * 4 - 1]) + 1


def test_stream_fastq_chunks_equal_whole_file(tmp_path):
    """the chunked reader cuts at record boundaries found by counting lines (a quality line may start with
    '@' or '+'), over plain / multi-member gzip / multi-frame zstd input and for chunk sizes from a few records to the whole
    file; the appending writer produces valid plain / multi-member gzip / multi-frame zstd streams."""
    from itsxpress_b200 import _zstd
    raw = gzip.open(SEQ, "rb").read()
    raw += b"@tricky one\nACGTAC\n+\n@+@+@+\n@tricky two\nAC\n+tricky two\n+@\n"
    whole = fq.parse_bytes(raw)
    assert whole.n == 229
    paths = {"plain": str(tmp_path / "a.fastq"), "gz": str(tmp_path / "a.fastq.gz"), "zst": str(tmp_path / "a.fastq.zst")}
    open(paths["plain"], "wb").write(raw)
    with open(paths["gz"], "wb") as f:
        f.write(gzip.compress(raw[:50_001]) + gzip.compress(raw[50_001:]))            # two members, cut mid-record
    open(paths["zst"], "wb").write(_zstd.compress(raw[:70_001]) + _zstd.compress(raw[70_001:]))
    want_seq, want_titles = whole.seq_concat()[0], [whole.title(i) for i in range(whole.n)]
    for kind, path in paths.items():
        for chunk in (900, 7777, 65_536, 1 << 30):
            seqs, titles, nchunks = [], [], 0
            for b in fq.stream_fastq(path, chunk):
                seqs.append(b.seq_concat()[0])
                titles += [b.title(i) for i in range(b.n)]
                nchunks += 1
            assert titles == want_titles and np.array_equal(np.concatenate(seqs), want_seq), (kind, chunk)
            assert nchunks > 50 if (chunk == 900 and kind != "zst") else nchunks >= 1
    for kw, name in ((dict(gzipped=True), "o.gz"), (dict(zstd_file=True), "o.zst"), (dict(), "o.fq")):
        w = fq.ChunkWriter(str(tmp_path / name), **kw)
        w.write(raw[:30_000], 10)
        w.write(b"", 0)
        w.write(raw[30_000:], 219)
        w.close()
        assert fq._open_bytes(str(tmp_path / name)) == raw and fq.cached_count(str(tmp_path / name)) == 229
    w = fq.ChunkWriter(str(tmp_path / "e.gz"), gzipped=True)
    w.close()
    assert gzip.open(str(tmp_path / "e.gz"), "rb").read() == b""
    with pytest.raises(ValueError):
        list(fq.stream_fastq(_write(tmp_path, "bad.fastq", raw[:-9]), 5000))             # truncated last record


def _write(tmp_path, name, data):
    p = str(tmp_path / name)
    open(p, "wb").write(data)
    return p


def test_q2_batches_group_small_samples_only(tmp_path):
    """consecutive small samples share a device pass up to the read budget; a sample that alone takes more
    than an eighth of the budget goes alone (per-sample path with read-ahead); order is preserved."""
    from collections import namedtuple
    Row = namedtuple("Row", "Index forward reverse")
    rows = []
    for k, nbytes in enumerate([1000, 2000, 400_000, 1500, 1500, 1500, 90_000, 10]):
        p = str(tmp_path / ("s%d.fastq" % k))
        open(p, "wb").write(b"x" * nbytes)
        rows.append(Row("S%d" % k, p, None))
    groups = q2._batches(rows, False, 1200)          # budget in reads; plain files: ~250 bytes per read
    names = [[r.Index for r in g] for g in groups]
    assert names == [["S0", "S1"], ["S2"], ["S3", "S4", "S5"], ["S6"], ["S7"]]
    assert [r for g in groups for r in g] == rows


def test_counter_based_generator_blocks_are_consistent():
    """synth_big: any block of the global sample can be produced on its own (what lets 8 ranks build configs[3] without
    any of them holding it), the abundance law is the stated one, reads carry the boundary motifs' length range."""
    import synth_big
    S = synth_big.BigSample("c4", scale=0.00005)
    seq, qual, off = S.block(0, S.N)
    a, q, o = S.block(1234, 3456)
    lo, hi = int(off[1234]), int(off[3456])
    assert bytes(a.numpy()) == bytes(seq.numpy()[lo:hi]) and bytes(q.numpy()) == bytes(qual.numpy()[lo:hi])
    assert np.array_equal(o.numpy(), off.numpy()[1234:3457] - lo)
    offn, seqn = off.numpy(), seq.numpy()
    reads = [seqn[offn[i]:offn[i + 1]].tobytes() for i in range(S.N)]
    from collections import Counter
    mult = Counter(Counter(reads).values())
    assert mult == {1: S.U - (S.N - S.U), 2: S.N - S.U} and len(set(reads)) == S.U
    lens = np.diff(offn)
    assert lens.min() >= 330 and lens.max() <= 441 and set(seqn.tolist()) <= set(b"ACGTN")
    Z = synth_big.BigSample("c3", scale=0.0003)
    zs, _, zo = Z.block(0, Z.N, want_qual=False)
    zr = [zs.numpy()[zo[i]:zo[i + 1]].tobytes() for i in range(Z.N)]
    assert len(set(zr)) == Z.U and max(Counter(zr).values()) > 50          # Zipf: one unique dominates


# ---- Dedup / ItsPosition from files -------------------------------------------------------------------------
def test_dedup_parse():
    # reference test_dedup :49-65
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    assert len(d.matchdict) == 227
    assert d.matchdict["M02696:28:000000000-ATWK5:1:1101:11740:1800"] == "M02696:28:000000000-ATWK5:1:1101:10899:1561"
    assert d.matchdict["M02696:28:000000000-ATWK5:1:1101:10899:1561"] == "M02696:28:000000000-ATWK5:1:1101:10899:1561"


def _golden_domtbl(path):
    """A domtbl with the rows the reference's fixture must have contained for two sequences
    (tests/test_main_pytest.py:36-46) plus a losing and a tying sibling row."""
    from itsxpress_b200.host import DOMTBL_HEADER
    row = "%-20s -          %5d %-20s -             45 %9.2g %6.1f   0.0   1   1 %9.2g %9.2g %6.1f   0.0     1    45 %5d %5d %5d %5d 0.90 -\n"
    a, b = "M02696:28:000000000-ATWK5:1:1101:19331:3209", "M02696:28:000000000-ATWK5:1:1101:23011:4341"
    rows = [
        (a, 341, "3_End_5_8S_fungi_a", 51.0, 85, 128),
        (a, 341, "3_End_5_8S_fungi_b", 52.2, 84, 128),
        (a, 341, "3_End_5_8S_fungi_c", 52.2, 80, 120),        # tie: the first 52.2 row wins
        (a, 341, "4_Start_LSU_fungi_a", 59.1, 282, 326),
        (a, 341, "4_Start_LSU_fungi_b", 40.0, 283, 326),
        (b, 385, "4_Start_LSU_fungi_a", 34.0, 327, 370),
    ]
    with open(path, "w") as f:
        f.write(DOMTBL_HEADER)
        for t, tlen, q, sc, i, j in rows:
            f.write(row % (t, tlen, q, 1e-10, sc, 1e-10, 1e-10, sc, i, j, i, j))
        f.write("#\n# [ok]\n")


def test_its_position_parse(tmp_path):
    # reference test_its_position_init :32-46
    p = str(tmp_path / "domtbl.txt")
    _golden_domtbl(p)
    its = ItsPosition(p, "ITS2")
    exp1 = {"tlen": 341, "right": {"score": 59.1, "to_pos": 326, "from_pos": 282},
            "left": {"score": 52.2, "to_pos": 128, "from_pos": 84}}
    exp2 = {"tlen": 385, "right": {"score": 34.0, "to_pos": 370, "from_pos": 327}}
    assert its.ddict["M02696:28:000000000-ATWK5:1:1101:19331:3209"] == exp1
    assert its.ddict["M02696:28:000000000-ATWK5:1:1101:23011:4341"] == exp2
    assert its.get_position("M02696:28:000000000-ATWK5:1:1101:19331:3209") == (128, 281, 341)
    assert its.get_position("M02696:28:000000000-ATWK5:1:1101:23011:4341") == (None, 326, 385)
    with pytest.raises(KeyError):
        its.get_position("never-seen")


def test_domtbl_emitter_is_parsed_back():
    """host.write_domtbl -> ItsPosition.parse returns the rows' own values (columns 0, 2, 3, 13, 19, 20)."""
    from itsxpress_b200 import _lib, host
    rows = np.zeros(3, _lib.ROW_DTYPE)
    rows["seq"] = [0, 0, 1]
    rows["prof"] = [0, 1, 1]
    rows["ienv"] = [84, 282, 300]
    rows["jenv"] = [128, 326, 344]
    rows["tlen"] = [341, 341, 400]
    rows["bitscore"] = [52.24, 59.06, 33.96]
    rows["seq_score"] = [52.0, 59.0, 33.0]
    rows["lnP"] = rows["seq_lnP"] = [-40.0, -44.0, -20.0]
    text = host.write_domtbl(rows, ["s0", "s1"], ["3_left", "4_right"], [45, 45], 2, np.array([1, 2]))
    with tempfile.NamedTemporaryFile("wb", suffix=".txt", delete=False) as f:
        f.write(text)
    try:
        its = ItsPosition(f.name, "ITS2")
        assert its.get_position("s0") == (128, 281, 341)
        assert its.ddict["s0"]["left"]["score"] == 52.2 and its.ddict["s0"]["right"]["score"] == 59.1
        assert its.get_position("s1") == (None, 299, 400)
        for line in text.decode().splitlines():
            if not line.startswith("#"):
                assert len(line.split()) == 23
    finally:
        os.unlink(f.name)


# ---- record generators with duck-typed itspos (reference :412-509) -------------------------------------------
class _Mock:
    def __init__(self, pos):
        self.pos = pos

    def get_position(self, seq_id):
        return self.pos


def test_native_ids_uc_and_rep_fasta_equal_the_python_ones():
    """record.id = title.split(None, 1)[0] from the native label scanner; uc.txt / rep.fa from the native emitters equal the
    Python reference emitters byte for byte (tricky titles, both strands, clusters of every size, 80-column wrapping), and
    titles with non-ASCII bytes take the Python path."""
    from itsxpress_b200 import host
    rng = np.random.default_rng(23)
    titles = ["r%d desc %d" % (i, i) for i in range(400)]
    titles[3], titles[4], titles[5], titles[6], titles[7] = " lead space", "\ttab first", "id\ttab", "", "onlyid"
    recs, seqs = [], []
    for i, t in enumerate(titles):
        L = int(rng.choice([1, 79, 80, 81, 160, 161, 250]))
        s = "".join(rng.choice(list("ACGT"), L))
        seqs.append(s)
        recs.append("@%s\n%s\n+\n%s\n" % (t, s, "I" * L))
    b = fq.parse_bytes("".join(recs).encode())
    want_ids = [(t.split(None, 1) or [""])[0] for t in titles]
    assert b.ids() == want_ids
    lab_off, lab_len = b.labels()
    assert [b.buf[o:o + l].tobytes().decode() for o, l in zip(lab_off.tolist(), lab_len.tolist())] == want_ids
    rep = np.arange(b.n, dtype=np.int32)
    for i in range(b.n):                                   # clusters of many sizes; representatives are first occurrences
        if i % 3 and i > 10:
            rep[i] = rep[int(rng.integers(0, i))]
    strand = rng.integers(0, 2, b.n).astype(np.uint8)
    order = host.cluster_order(rep, b.ids())
    assert host.write_uc(rep, strand, b.ids(), b.s_len, order, batch=b) == host.write_uc_py(rep, strand, b.ids(), b.s_len, order)
    assert host.write_uc(rep, None, b.ids(), b.s_len, order, batch=b) == host.write_uc_py(rep, None, b.ids(), b.s_len, order)
    assert host.write_rep_fasta(b, order, b.ids()) == host.write_rep_fasta_py(b, order, b.ids())
    assert host.write_rep_fasta(b, order[:0], b.ids()) == b"" and host.write_uc(rep[:0], None, [], b.s_len[:0], order[:0], batch=fq.parse_bytes(b"")) == b""
    odd = fq.parse_bytes("@r\u00e9sum\u00e9 1\nACGT\n+\nIIII\n@plain\nAC\n+\nII\n".encode("latin-1"))
    r2 = np.array([0, 1], np.int32)
    assert host.write_uc(r2, None, odd.ids(), odd.s_len, np.array([0, 1]), batch=odd) == host.write_uc_py(r2, None, odd.ids(), odd.s_len, [0, 1])
    assert host.write_rep_fasta(odd, np.array([0, 1]), odd.ids()) == host.write_rep_fasta_py(odd, [0, 1], odd.ids())


def test_native_domtbl_formatter_equals_the_python_one():
    """csrc/fastq_host.cpp::itsx_domtbl_format against host.write_domtbl_py, byte for byte: label padding (shorter than, equal
    to and longer than the 20-column field), E-values from 1e-300 to 1e+5 and 0, scores that overflow their field width,
    several domains per hit (running index / count), rows out of hit order, many rows (several threads)."""
    from itsxpress_b200 import _lib, host
    rng = np.random.default_rng(17)
    nseq, nprof, n = 300, 7, 60000
    seq_ids = ["s%d" % i + "x" * int(rng.integers(0, 30)) for i in range(nseq)]
    seq_ids[0], seq_ids[1] = "a" * 20, "b" * 19
    prof_names = ["3_left_%d" % i + "y" * (5 * i) for i in range(nprof)]
    M = rng.integers(20, 46, nprof).astype(np.int32)
    nrep = rng.integers(0, 1000, nprof).astype(np.int32)
    rows = np.zeros(n, dtype=_lib.ROW_DTYPE)
    rows["seq"] = rng.integers(0, nseq, n)
    rows["prof"] = rng.integers(0, nprof, n)
    rows["ienv"] = rng.integers(1, 400, n)
    rows["jenv"] = rows["ienv"] + rng.integers(0, 60, n)
    rows["tlen"] = rng.integers(50, 100000, n)
    rows["bitscore"] = (rng.normal(30, 40, n)).astype(np.float32)
    rows["seq_score"] = (rng.normal(30, 4000, n)).astype(np.float32)
    rows["lnP"] = -rng.exponential(30, n) * rng.choice([0.01, 1, 25], n)
    rows["seq_lnP"] = -rng.exponential(30, n) * rng.choice([0.01, 1, 25], n)
    rows["lnP"][:5] = [-1e4, 0.0, -745.0, -709.0, 11.0]
    for k in (0, 1, 17, 2000, n):
        a = host.write_domtbl(rows[:k], seq_ids, prof_names, M, nseq, nrep)
        assert a == host.write_domtbl_py(rows[:k], seq_ids, prof_names, M, nseq, nrep), k
    # labels that are not ASCII take the Python path (padding counts characters)
    odd = ["\ufffdid"] + seq_ids[1:]
    assert host.write_domtbl(rows[:50], odd, prof_names, M, nseq, nrep) == host.write_domtbl_py(rows[:50], odd, prof_names, M, nseq, nrep)


def _dedup_seq1():
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    d.matchdict = {"seq1": "seq1"}
    return d


def test_coordinate_zero_not_falsy():
    d = _dedup_seq1()
    recs = [fq.Record("ATCG" * 50, id="seq1", description="")]
    out = list(d._get_trimmed_seq_generator(iter(recs), _Mock((0, 10, 100)), wri_file=True))
    assert len(out) == 1 and str(out[0].seq) == "ATCGATCGAT"


def test_generator_exhaustion_wri_file_false():
    d = _dedup_seq1()
    recs = [fq.Record("ATCG" * 50, id="seq1", description="")]
    out = list(d._get_trimmed_seq_generator(iter(recs), _Mock((5, 15, 100)), wri_file=False))
    assert len(out) == 1 and len(out[0].seq) == 10


def test_trim_ccs_stitching():
    d = _dedup_seq1()
    rec = fq.Record("N" * 50, id="seq1", description="", quals=[40] * 50)
    out = list(d._get_trimmed_seq_generator(iter([rec]), _Mock((5, 15, 100)), wri_file=True, trim_ccs=True))
    assert len(out) == 1
    assert str(out[0].seq) == "GACAGGTACAAGAAGGA" + "N" * 10 + "TTAACCCAGTCTCCAGT"
    assert out[0].letter_annotations["phred_quality"] == [93] * 17 + [40] * 10 + [93] * 17


def test_generators_drop_unmapped_and_inverted():
    d = _dedup_seq1()
    recs = [fq.Record("ACGT" * 10, id="other"), fq.Record("ACGT" * 10, id="seq1")]
    assert list(d._get_trimmed_seq_generator(iter(recs), _Mock((10, 10, 40)), True)) == []
    assert list(d._get_trimmed_seq_generator(iter(recs), _Mock((None, 10, 40)), True)) == []

    class Raises:
        def get_position(self, s):
            raise KeyError

    assert list(d._get_trimmed_seq_generator(iter(recs), Raises(), True)) == []


class _GoldenPos:
    """get_position from the recovered golden table (= what the missing domtbl.txt fixture implied)."""

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


def test_single_generator_golden_counts():
    # reference test_dedup_create_trimmed_seqs :68-98 -> 226 records, 42 637 bases
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    out = list(d._get_trimmed_seq_generator(fq.iter_records(SEQ), _GoldenPos(), wri_file=True))
    assert len(out) == 226
    assert sum(len(r) for r in out) == 42637


def test_paired_generator_golden_bytes():
    # reference test_get_paired_seq_generator :350-375 and test_create_paired_trimmed_seqs :378-397
    r1, r2 = os.path.join(TD, "4774-1-MSITS3_R1.fastq"), os.path.join(TD, "4774-1-MSITS3_R2.fastq")
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ, fastq=r1, fastq2=r2)
    g1, g2 = d._get_paired_seq_generator(zip(fq.iter_records(r1), fq.iter_records(r2)), _GoldenPos(), wri_file=True)
    a, b = list(g1), list(g2)
    assert len(a) == 226 and len(b) == 226
    assert "".join(r.format("fastq") for r in a).encode() == open(os.path.join(TD, "t2_r1.fq"), "rb").read()
    assert "".join(r.format("fastq") for r in b).encode() == open(os.path.join(TD, "t2_r2.fq"), "rb").read()


def test_paired_mixed_compression_rejected():
    d = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ, fastq=os.path.join(TD, "4774-1-MSITS3_R1.fastq.gz"),
              fastq2=os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    with pytest.raises(ValueError):
        d.create_paired_trimmed_seqs("/tmp/x1", "/tmp/x2", False, False, _GoldenPos(), True)
    d2 = Dedup(uc_file=UC, rep_file=REP, seq_file=SEQ)
    with pytest.raises(ValueError):
        d2.create_paired_trimmed_seqs("/tmp/x1", "/tmp/x2", False, False, _GoldenPos(), True)


# ---- CLI plumbing --------------------------------------------------------------------------------------------
def test_is_paired():
    # reference :196-204
    assert cli._is_paired("fastq1.fq", "fastq2.fq", False)
    assert not cli._is_paired("fastq1.fq", None, True)
    assert not cli._is_paired("fastq1.fq", None, False)
    with pytest.raises(AssertionError):
        cli._is_paired(None, None, False)


def test_myparser():
    # reference :207-223
    p = cli.myparser()
    a = p.parse_args(["--fastq", "test.fastq", "--outfile", "test.out", "--region", "ITS2", "--taxa", "Fungi"])
    assert a.fastq == "test.fastq" and a.outfile == "test.out" and a.region == "ITS2" and a.taxa == "Fungi"
    assert a.threads == 1 and a.cluster_id == 1.0 and a.log == "ITSxpress.log" and a.allow_staggered_reads is True
    assert not a.keeptemp and not a.single_end and not a.reversed_primers and not a.trim_ccs
    assert p.parse_args(["-f", "a", "-o", "b", "--region", "ALL", "--taxa", " Rhizaria"]).taxa == " Rhizaria"
    with pytest.raises(SystemExit):
        p.parse_args(["-f", "a", "-o", "b", "--region", "ITS2", "--cluster_id", "0.9"])
    with pytest.raises(SystemExit):
        p.parse_args(["-f", "a", "-o", "b", "--region", "ITS3"])


def test_create_runtime_hmm():
    # reference :512-569 (Metazoa stands in for Fungi: F.hmm is missing from the mount)
    tmp = tempfile.mkdtemp()
    try:
        def names(path):
            return [l[6:].strip() for l in open(path) if l.startswith("NAME  ")]
        for region, ok, bad in (("ITS2", ("3_", "4_"), ("1_", "2_")), ("ITS1", ("1_", "2_"), ("3_", "4_")),
                                ("ALL", ("1_", "4_"), ("2_", "3_"))):
            n = names(cli.create_runtime_hmm("Metazoa", region, tmp))
            assert n and all(x.startswith(ok) and not x.startswith(bad) for x in n)
        assert len(names(cli.create_runtime_hmm("Metazoa", "ITS2", tmp))) == 256
        n_all = names(cli.create_runtime_hmm("All", "ITS2", tmp))
        assert len(n_all) == 814 and all(x.startswith(("3_", "4_")) for x in n_all)
        # a taxon whose file is absent gives an empty profile file (reference main.py:214-215)
        assert names(cli.create_runtime_hmm("Fungi", "ITS2", tmp)) == [] or os.path.exists(os.path.join(HMM_DIR, "F.hmm"))
    finally:
        shutil.rmtree(tmp)


def test_suffix_dispatch():
    assert cli._suffix_flags("a.fq.gz") == (True, False)
    assert cli._suffix_flags("a.fq.zst", "b.zst") == (False, True)
    assert cli._suffix_flags("a.fq.gz", "b.fq") == (False, False)
    assert cli._suffix_flags("a.fastq") == (False, False)


def test_names_look_paired():
    assert cli._names_look_paired("r1 1:N:0:1", "r1 2:N:0:1")
    assert cli._names_look_paired("r1/1", "r1/2")
    assert not cli._names_look_paired("r1 1:N:0:1", "r2 1:N:0:1")


def test_q2_directory_formats():
    """MANIFEST reader of the per-sample directory format (the reference's fixtures under tests/test_data)."""
    from itsxpress_b200 import q2_itsxpress as q2
    p = q2.PerSampleDir(os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data"))
    f = p.manifest.view(None)
    assert list(f.columns) == ["forward", "reverse"] and list(f.index) == ["4774-1-MSITS3"]
    assert f.loc["4774-1-MSITS3", "forward"].endswith("4774-1-MSITS3_0_L001_R1_001.fastq.gz")
    assert f.loc["4774-1-MSITS3", "reverse"].endswith("4774-1-MSITS3_1_L001_R2_001.fastq.gz")
    assert q2._taxa_prefix_to_taxa("F") == "Fungi" and q2._taxa_prefix_to_taxa("ALL") == "All"
    assert q2._taxa_prefix_to_taxa("R") == "Rhizaria"          # upstream quirk: not the " Rhizaria" key


def test_native_fastq_matches_numpy_reference():
    """csrc/fastq_host.cpp (scanner, packer, formatter) against the independent numpy implementation."""
    raw = gzip.open(SEQ, "rb").read()
    for data in (raw, raw.replace(b"\n", b"\r\n"), raw + b"\n\n", raw[:-1],
                 raw.replace(b"\n+\n", b"\n+M02696:28:000000000-ATWK5:1:1101:21090:1000 1:N:0:108\n", 1)):
        a, b = fq.parse_bytes(data), fq._parse_bytes_numpy(data)
        assert a.n == b.n == 227
        for f in ("t_off", "t_len", "s_off", "s_len", "q_off"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    a = fq.parse_bytes(raw)
    s1, o1 = fq._gather(a.buf, a.s_off, a.s_len)
    s2, o2 = fq._gather_numpy(a.buf, a.s_off, a.s_len)
    assert np.array_equal(s1, s2) and np.array_equal(o1, o2)
    q1, _ = a.qual_concat()
    ki = np.arange(0, a.n, 3)
    oo = np.zeros(len(ki) + 1, np.int64)
    oo[1:] = np.cumsum(a.s_len[ki])
    seq_k, _ = fq._gather(a.buf, a.s_off[ki], a.s_len[ki])
    qual_k, _ = fq._gather(a.buf, a.q_off[ki], a.s_len[ki])
    for pre, suf in ((None, None), ((b"GAC", b"~~~"), (b"TT", b"~~"))):
        assert fq.format_gathered(a, ki, oo, seq_k, qual_k, pre, suf) == \
            fq._format_gathered_numpy(a, ki, oo, seq_k, qual_k, pre, suf)
    for bad in (raw[:200], raw.replace(b"\n+\n", b"\n-\n", 1), b"@x\nACGT\n+\nII\n", b"@x\nAC\n+\nI\x01\n",
                raw.replace(b"\n+\n", b"\n+other\n", 1)):
        with pytest.raises(ValueError):
            fq.parse_bytes(bad)
        with pytest.raises(ValueError):
            fq._parse_bytes_numpy(bad)
    assert fq.parse_bytes(b"").n == 0


def test_batch_of_gathered_equals_reparse():
    """The batch _merge_reads hands to deduplicate() is what scanning the written seq.fq would give."""
    from itsxpress_b200 import fastq as fq
    b1 = fq.read_fastq(os.path.join(TD, "4774-1-MSITS3_R1.fastq"))
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(b1.n, 180, replace=False)).astype(np.int32)
    lens = rng.integers(1, 400, len(idx))
    off = np.zeros(len(idx) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(off[-1]))
    qual = rng.integers(33, 74, int(off[-1])).astype(np.uint8)
    data = fq.format_gathered(b1, idx, off, seq, qual)
    a, b = fq.parse_bytes(data), fq.batch_of_gathered(data, b1, idx, off)
    for k in ("t_off", "t_len", "s_off", "s_len", "q_off"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.n == b.n and np.array_equal(b.seq_concat()[0], seq) and np.array_equal(b.qual_concat()[0], qual)
    assert fq.batch_of_gathered(b"", b1, np.zeros(0, np.int32), np.zeros(1, np.int64)).n == 0


def test_cached_read_counts_and_parallel_gzip(tmp_path):
    """Counts of files this process scanned / wrote are reused only while size and mtime are unchanged; the parallel
    gzip writer's multi-member stream inflates to the input bytes with any reader."""
    import subprocess
    import time
    from itsxpress_b200 import main as cli
    src = os.path.join(TD, "4774-1-MSITS3_R1.fastq")
    b = fq.read_fastq(src)
    assert fq.cached_count(src) == b.n == 250
    text = open(src, "rb").read() * 40                       # 4 MB members -> several of them
    out = str(tmp_path / "o.fastq.gz")
    fq.write_compressed(out, text, gzipped=True, n_records=250 * 40)
    assert fq.cached_count(out) == 10000
    assert gzip.open(out, "rb").read() == text
    assert subprocess.run(["gzip", "-dc", out], stdout=subprocess.PIPE).stdout == text
    plain = str(tmp_path / "p.fastq")
    fq.write_compressed(plain, text[:len(text) // 40], n_records=250)
    assert fq.cached_count(plain) == 250
    time.sleep(0.01)
    with open(plain, "ab") as f:                             # modified behind our back: the cache must not answer
        f.write(open(src, "rb").read())
    assert fq.cached_count(plain) is None
    assert fq.cached_count(str(tmp_path / "missing.fastq")) is None
    import logging
    records = []
    h = logging.Handler()
    h.emit = lambda r: records.append(r.getMessage())
    logging.getLogger().addHandler(h)
    old = logging.getLogger().level
    logging.getLogger().setLevel(logging.INFO)
    try:
        cli._check_total_reads(plain, out)
    finally:
        logging.getLogger().removeHandler(h)
        logging.getLogger().setLevel(old)
    assert any(m.endswith("p.fastq is 500.") for m in records) and any(m.endswith("o.fastq.gz is 10000.") for m in records)
    a, c = fq.read_fastq_many([src, os.path.join(TD, "4774-1-MSITS3_R2.fastq.gz")])
    assert a.n == c.n == 250 and a.title(0).split()[0] == c.title(0).split()[0]


def test_q2_shell_entry_single_process(tmp_path, monkeypatch):
    """`python -m itsxpress_b200.q2_itsxpress --in DIR --out DIR ...` with one process: every sample goes through the
    per-sample pipeline (stubbed here: the real one needs a B200) with the options of the chosen plugin action."""
    from itsxpress_b200 import q2_itsxpress as q2
    src = os.path.join(TD, "paired", "445cf54a-bf06-4852-8010-13a60fa1598c", "data")
    seen = []

    def stub(sample, results, tempdir, threads, taxa, region, paired_in, paired_out, reversed_primers, stagger, *rest):
        seen.append((sample.Index, taxa, region, paired_in, paired_out, reversed_primers, stagger))
        shutil.copy(sample.forward, os.path.join(str(results), os.path.basename(sample.forward)))

    monkeypatch.setattr(q2, "_process_sample", stub)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    out = str(tmp_path / "o")
    assert q2.cli(["--in", src, "--out", out, "--region", "ITS2", "--taxa", "M", "--mode", "pair", "--no-staggered"]) == 0
    assert seen == [("4774-1-MSITS3", "Metazoa", "ITS2", True, False, False, False)]
    assert open(os.path.join(out, "MANIFEST")).read().splitlines()[1].endswith(",forward")
    with pytest.raises(SystemExit):
        q2.cli(["--in", src, "--out", out, "--region", "ITS9"])


def test_fastq_read_ahead(tmp_path):
    """fq.prefetch reads files in the background; read_fastq collects the result once, falls back to an ordinary read
    when the file changed in between, and raises errors at the point of use (not in prefetch)."""
    import time
    r1, r2 = os.path.join(TD, "4774-1-MSITS3_R1.fastq.gz"), os.path.join(TD, "4774-1-MSITS3_R2.fastq")
    want1, want2 = fq._read_fastq_now(r1), fq._read_fastq_now(r2)
    fq.prefetch([r1, r2, None])
    assert len(fq._PREFETCH) == 2
    a, b = fq.read_fastq_many([r1, r2])
    assert not fq._PREFETCH
    assert np.array_equal(a.buf, want1.buf) and np.array_equal(b.s_off, want2.s_off) and a.n == b.n == 250
    assert fq.read_fastq(r1).n == 250                       # nothing pending: ordinary read
    # a file that changes between read-ahead and use is read again
    p = str(tmp_path / "x.fastq")
    shutil.copy(r2, p)
    fq.prefetch([p])
    time.sleep(0.05)
    with open(p, "ab") as f:
        f.write(open(r2, "rb").read())
    assert fq.read_fastq(p).n == 500
    # errors surface where the reference raises them
    fq.prefetch([os.path.join(TD, "broken.fastq"), str(tmp_path / "missing.fastq")])
    with pytest.raises(ValueError):
        fq.read_fastq(os.path.join(TD, "broken.fastq"))
    with pytest.raises(FileNotFoundError):
        fq.read_fastq(str(tmp_path / "missing.fastq"))
    fq.prefetch([r1])
    fq.drop_prefetched()
    assert not fq._PREFETCH


def test_q2_loop_reads_next_sample_ahead(tmp_path, monkeypatch):
    """The per-sample loops of the QIIME 2 driver start the next sample's read while the current one is processed."""
    from itsxpress_b200 import q2_itsxpress as q2
    src = tmp_path / "in"
    src.mkdir()
    lines = ["sample-id,filename,direction"]
    for k in range(3):
        for m, tag, d in ((1, "R1", "forward"), (2, "R2", "reverse")):
            fn = "S%d_%d_L001_%s_001.fastq.gz" % (k, k, tag)
            shutil.copy(os.path.join(TD, "4774-1-MSITS3_R%d.fastq.gz" % m), str(src / fn))
            lines.append("S%d,%s,%s" % (k, fn, d))
    (src / "MANIFEST").write_text("\n".join(lines) + "\n")
    pending = []

    def stub(sample, results, *rest):
        pending.append(sorted(os.path.basename(p) for p in fq._PREFETCH))
        b1, b2 = fq.read_fastq_many([sample.forward, sample.reverse])       # what _merge_reads does
        assert b1.n == b2.n == 250
        shutil.copy(sample.forward, os.path.join(str(results), os.path.basename(sample.forward)))

    monkeypatch.setattr(q2, "_process_sample", stub)
    monkeypatch.setattr(q2, "BATCH_READS", 0)          # the per-sample loop (small samples would share a device pass)
    monkeypatch.setattr(fq, "READ_AHEAD", 1)
    q2.main_sharded(q2.PerSampleDir(str(src)), str(tmp_path / "o"), region="ITS2", taxa="M", rank=0, world=1)
    # at the start of a sample its own files (read ahead during the previous one) and the next sample's are pending
    assert [len(p) for p in pending] == [2, 4, 2] and not fq._PREFETCH
    assert pending[0] == ["S1_1_L001_R1_001.fastq.gz", "S1_1_L001_R2_001.fastq.gz"]
    # two samples ahead: both followers are being read when the first sample starts; a byte budget of zero keeps one
    del pending[:]
    monkeypatch.setattr(fq, "READ_AHEAD", 2)
    q2.main_sharded(q2.PerSampleDir(str(src)), str(tmp_path / "o2"), region="ITS2", taxa="M", rank=0, world=1)
    assert [len(p) for p in pending] == [4, 4, 2] and not fq._PREFETCH
    del pending[:]
    monkeypatch.setattr(q2, "READ_AHEAD_BYTES", 0)
    q2.main_sharded(q2.PerSampleDir(str(src)), str(tmp_path / "o3"), region="ITS2", taxa="M", rank=0, world=1)
    assert [len(p) for p in pending] == [2, 4, 2] and not fq._PREFETCH
