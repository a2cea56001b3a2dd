"""The native gzip reader (itsxpress_b200/csrc/inflate_host.cpp) against zlib: every block type, multi-member files,
header extras, buffers that have to grow, damaged streams (which must end up as the gzip module's own errors), reads into
buffers of any size, and ONE deflate stream inflated on several cores (chunking shrunk so that small inputs take that
path; block starts that are none; members ending inside chunks)."""
import gzip
import io
import os
import zlib

import numpy as np
import pytest

from itsxpress_b200 import fastq as fq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TD = os.path.join(ROOT, "tests", "test_data")


def _member(data, level=6, **kw):
    buf = io.BytesIO()
    with gzip.GzipFile(fileobj=buf, mode="wb", compresslevel=level, **kw) as g:
        g.write(data)
    return buf.getvalue()


def _payloads():
    rng = np.random.default_rng(3)
    with open(os.path.join(TD, "4774-1-MSITS3_merged.fastq"), "rb") as f:
        fastq = f.read()
    return {
        "empty": b"", "one byte": b"A", "short": b"ACGTACGTACGT\n", "run": b"I" * 300000,
        "acgt": bytes(rng.integers(65, 69, 200000, dtype=np.uint8)),
        "random": bytes(rng.integers(0, 256, 150000, dtype=np.uint8)),
        "skewed": bytes(np.minimum(rng.geometric(0.02, 250000), 255).astype(np.uint8)),
        "fastq": fastq, "fastq x12": fastq * 12,
    }


SMALL = (512, 2048, 3000)          # chunk_min, chunk_max, par_min of the multi-core mode for small inputs


def test_gunzip_equals_zlib_on_every_level_and_block_type():
    for name, data in _payloads().items():
        for level in (0, 1, 3, 6, 9):                  # 0: stored blocks; tiny inputs: fixed Huffman codes
            comp = _member(data, level)
            assert fq.gunzip(comp, 1).tobytes() == data, (name, level)
            assert fq.gunzip(comp, 4, SMALL).tobytes() == data, (name, level, "4 threads")
        # fixed Huffman on a long input (zlib's Z_FIXED strategy)
        co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
        comp = co.compress(data) + co.flush()
        assert fq.gunzip(comp, 1).tobytes() == data and fq.gunzip(comp, 3, SMALL).tobytes() == data, name
        # many small blocks (sync flushes put empty stored blocks between them)
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        parts = [co.compress(data[i:i + 7000]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(data), 7000)]
        comp = b"".join(parts) + co.flush()
        assert fq.gunzip(comp, 1).tobytes() == data and fq.gunzip(comp, 8, SMALL).tobytes() == data, name


def test_gunzip_multi_member_header_fields_and_growth():
    p = _payloads()
    data = p["fastq"]
    members = [_member(data[i:i + 20000], 6) for i in range(0, len(data), 20000)]
    for th, tune in ((1, None), (4, SMALL), (8, (256, 256, 0))):
        assert fq.gunzip(b"".join(members), th, tune).tobytes() == data       # the ISIZE guess is far too small: the buffer grows
        assert fq.gunzip(b"".join(members) + b"\0" * 37, th, tune).tobytes() == data          # zero padding behind the last member
    with_name = _member(data, 6, filename="reads.fastq", mtime=12345)
    assert fq.gunzip(with_name, 1).tobytes() == data
    # FEXTRA + FCOMMENT + FHCRC written by hand around a raw deflate stream
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    hdr = bytes([0x1f, 0x8b, 8, 4 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"extra" + b"a comment\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")
    trailer = zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")
    for th, tune in ((1, None), (4, SMALL)):
        assert fq.gunzip(hdr + body + trailer, th, tune).tobytes() == data
        assert fq.gunzip(hdr + body + trailer + members[0], th, tune).tobytes() == data + data[:20000]
    # real files of the reference's test data
    for name in ("4774-1-MSITS3_R1.fastq.gz", "4774-1-MSITS3_R2.fastq.gz"):
        path = os.path.join(TD, name)
        with open(path, "rb") as f:
            comp = f.read()
        assert fq.gunzip(comp).tobytes() == gzip.open(path, "rb").read()
        assert fq.gunzip(comp, 4, SMALL).tobytes() == gzip.open(path, "rb").read()


def _read_all(comp, threads, tune, caps, contiguous=False):
    r = fq.GzReader(comp, threads, tune)
    parts, i = [], 0
    if contiguous:
        out, done = np.empty(1 << 22, np.uint8), 0
        while True:
            k = r.readinto(out[:min(out.size, done + caps[i % len(caps)])], done, done)
            i += 1
            assert k >= 0
            if k == 0:
                break
            done += k
        stats = (r.stat(0), r.stat(1))
        r.close()
        return out[:done].tobytes(), stats
    while True:
        out = np.empty(caps[i % len(caps)], np.uint8)
        i += 1
        k = r.readinto(out)
        assert k >= 0
        if k == 0:
            break
        parts.append(out[:k].tobytes())
    stats = (r.stat(0), r.stat(1))
    r.close()
    return b"".join(parts), stats


def test_reader_fills_buffers_of_any_size():
    data = _payloads()["fastq x12"][:700000]
    for level in (1, 6):
        comp = _member(data, level)
        for threads, tune in ((1, None), (3, SMALL), (8, (256, 512, 0))):
            for caps in ([1, 7, 1], [100, 900, 5000], [100000], [65535, 65536, 65537], [300000, 1, 1 << 20]):
                got, _ = _read_all(comp, threads, tune, caps)
                assert got == data, (level, threads, caps)
                got, _ = _read_all(comp, threads, tune, caps, contiguous=True)
                assert got == data, (level, threads, caps, "contiguous")
    r = fq.GzReader(b"", 1)
    assert r.readinto(np.empty(10, np.uint8)) == 0
    r.close()


def test_one_stream_on_several_cores_survives_block_starts_that_are_none():
    """A stored block whose payload is the middle of another deflate stream: chunks that start inside it find perfectly
    valid dynamic block headers that are not block starts of THIS stream.  The predecessor runs past them, the stitching
    drops them and decodes again up to the next real start."""
    fastq = _payloads()["fastq x12"]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)                   # a new block (and header) about every 700 bytes
    decoy = b"".join(co.compress(fastq[i:i + 2000]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(400000, 700000, 2000))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body, data = [], []
    for r in range(8):
        piece = fastq[r * 30000:(r + 1) * 30000]                   # real blocks, small ones: ends byte-aligned, not final
        body += [co.compress(piece[i:i + 2500]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(piece), 2500)]
        body.append(co.flush(zlib.Z_FULL_FLUSH))                   # what follows the inserted block must not reach across it
        payload = decoy[1000 + 7000 * r:1000 + 7000 * r + 5000]    # one stored block full of block headers of another stream
        body.append(b"\0" + len(payload).to_bytes(2, "little") + (len(payload) ^ 0xffff).to_bytes(2, "little") + payload)
        data += [piece, payload]
    body.append(co.compress(fastq[240000:300000]) + co.flush())
    data = b"".join(data) + fastq[240000:300000]
    comp = (bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3]) + b"".join(body) +
            zlib.crc32(data).to_bytes(4, "little") + len(data).to_bytes(4, "little"))
    assert gzip.decompress(comp) == data
    bridged = 0
    for threads, tune in ((4, (2048, 2048, 0)), (8, (1024, 4096, 0)), (16, (512, 512, 0)), (3, (8192, 8192, 0))):
        got, (batches, bridges) = _read_all(comp, threads, tune, [1 << 22])
        assert got == data and batches >= 1, (threads, tune)
        bridged += bridges
        assert fq.gunzip(comp, threads, tune).tobytes() == data
    assert bridged > 0                                             # the decoys were found, refused and decoded over
    # damage inside a chunk of the multi-core mode is still an error (and then the gzip module's)
    bad = bytearray(comp)
    bad[len(bad) // 2] ^= 0x10
    with pytest.raises((gzip.BadGzipFile, zlib.error, EOFError)):
        fq.gunzip(bytes(bad), 4, (2048, 2048, 0))


def test_many_small_members_start_chunks_at_member_headers():
    """BGZF (bgzip, some sequencers): members of <= 64 KB with an extra field and ONE final block each -- there is no
    non-final block header to find, so chunks of the multi-core mode start at member headers."""
    data = _payloads()["fastq x12"][:900000]
    members = []
    for i in range(0, len(data), 30000):
        piece = data[i:i + 30000]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(piece) + co.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        members.append(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0]) + bsize.to_bytes(2, "little") + body +
                       zlib.crc32(piece).to_bytes(4, "little") + len(piece).to_bytes(4, "little"))
    comp = b"".join(members) + bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])   # BGZF EOF
    assert gzip.decompress(comp) == data
    for threads, tune in ((1, None), (4, (4096, 16384, 0)), (8, (2048, 2048, 0))):
        got, (batches, _) = _read_all(comp, threads, tune, [1 << 21])
        assert got == data and (batches > 0) == (threads > 1), (threads, tune)
        assert fq.gunzip(comp, threads, tune).tobytes() == data
    bad = bytearray(comp)
    bad[len(members[0]) + len(members[1]) - 3] ^= 1                  # ISIZE of the second member
    with pytest.raises(gzip.BadGzipFile):
        fq.gunzip(bytes(bad), 4, (4096, 16384, 0))


def test_damaged_streams_raise_what_the_gzip_module_raises():
    data = _payloads()["fastq"]
    comp = bytearray(_member(data, 6))
    for th, tune in ((1, None), (4, SMALL)):
        with pytest.raises(EOFError):
            fq.gunzip(bytes(comp[:len(comp) // 2]), th, tune)         # truncated
        bad = bytearray(comp)
        bad[-5] ^= 0xff                                               # CRC
        with pytest.raises(gzip.BadGzipFile):
            fq.gunzip(bytes(bad), th, tune)
        bad = bytearray(comp)
        bad[len(bad) // 2] ^= 0x55                                    # somewhere in the deflate stream
        with pytest.raises((gzip.BadGzipFile, zlib.error, EOFError)):
            fq.gunzip(bytes(bad), th, tune)
        with pytest.raises(gzip.BadGzipFile):
            fq.gunzip(b"this is not a gzip file at all, just text\n", th, tune)


def test_readers_use_the_native_gunzip(tmp_path):
    path = os.path.join(TD, "4774-1-MSITS3_R1.fastq.gz")
    b = fq.read_fastq(path)
    assert b.n == 250 and b.buf.tobytes() == gzip.open(path, "rb").read()
    # the chunked reader: whole file, damaged file (zlib takes over where the native reader stopped and raises)
    data = _payloads()["fastq x12"]
    good = str(tmp_path / "good.fastq.gz")
    with open(good, "wb") as f:
        f.write(_member(data[:300000], 6) + _member(data[300000:], 1) + b"\0" * 9)
    assert b"".join(fq._raw_blocks(good, 50000)) == data
    comp = bytearray(_member(data, 6))
    cut = str(tmp_path / "cut.fastq.gz")
    with open(cut, "wb") as f:
        f.write(comp[:len(comp) // 2])
    with pytest.raises(EOFError):
        b"".join(fq._raw_blocks(cut, 50000))
    comp[len(comp) // 2] ^= 0x55
    bad = str(tmp_path / "bad.fastq.gz")
    with open(bad, "wb") as f:
        f.write(comp)
    with pytest.raises((zlib.error, EOFError)):
        b"".join(fq._raw_blocks(bad, 50000))


def test_crc_paths_agree_without_carry_less_multiplication():
    """CRC-32 runs by PCLMULQDQ folding where the CPU has it (checked against the table-driven register at start-up); without
    it, slice-by-8 on a thread that follows the decoder.  ITSX_NO_CLMUL forces the second path (fresh process)."""
    import subprocess
    import sys
    code = ("import sys, gzip, numpy as np; sys.path.insert(0, %r); from itsxpress_b200 import fastq as fq; "
            "rng = np.random.default_rng(3); "
            "data = bytes(rng.integers(33, 75, 12_000_000, dtype=np.uint8)); comp = gzip.compress(data, 1); "
            "assert len(comp) > (3 << 20); "
            "assert fq.gunzip(comp, 1).tobytes() == data and fq.gunzip(comp, 4).tobytes() == data; "
            "bad = bytearray(comp); bad[-6] ^= 1; "
            "import pytest; "
            "pytest.raises(gzip.BadGzipFile, fq.gunzip, bytes(bad), 1); pytest.raises(gzip.BadGzipFile, fq.gunzip, bytes(bad), 4); "
            "print('ok')" % ROOT)
    for env in ({}, {"ITSX_NO_CLMUL": "1"}):
        r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]


def test_the_bench_writes_its_gz_input_as_one_member():
    """bench.py's gz-in / gz-out leg feeds the command line ONE gzip member, as a sequencer writes it (pieces compressed
    concurrently the way pigz does it): a valid stream for zlib and for the native reader."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from cli_e2e import gzip_one_member
    data = _payloads()["fastq x12"]
    comp = gzip_one_member(data, piece=100_000)
    d = zlib.decompressobj(31)
    assert d.decompress(comp) == data and d.eof and d.unused_data == b""          # one member, nothing behind it
    assert fq.gunzip(comp, 1).tobytes() == data and fq.gunzip(comp, 4, SMALL).tobytes() == data
    assert gzip.decompress(gzip_one_member(b"")) == b""
