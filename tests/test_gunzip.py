"""The native gzip reader (itsxpress_b200/csrc/inflate_host.cpp) against zlib: every block type, multi-member files,
header extras, buffers that have to grow, damaged streams (which must end up as the gzip module's own errors)."""
import ctypes as C
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


def test_gunzip_equals_zlib_on_every_level_and_block_type():
    for name, data in _payloads().items():
        for level in (0, 1, 3, 6, 9):                  # 0: stored blocks; tiny inputs: fixed Huffman codes
            comp = _member(data, level)
            assert fq.gunzip(comp).tobytes() == data, (name, level)
        # fixed Huffman on a long input (zlib's Z_FIXED strategy)
        co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
        assert fq.gunzip(co.compress(data) + co.flush()).tobytes() == data, name
        # many small blocks (sync flushes put empty stored blocks between them)
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        parts = [co.compress(data[i:i + 7000]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(data), 7000)]
        assert fq.gunzip(b"".join(parts) + co.flush()).tobytes() == data, name


def test_gunzip_multi_member_header_fields_and_growth():
    p = _payloads()
    data = p["fastq"]
    members = [_member(data[i:i + 20000], 6) for i in range(0, len(data), 20000)]
    assert fq.gunzip(b"".join(members)).tobytes() == data                     # the ISIZE guess is far too small: the buffer grows
    with_name = _member(data, 6, filename="reads.fastq", mtime=12345)
    assert fq.gunzip(with_name).tobytes() == data
    # FEXTRA + FCOMMENT + FHCRC written by hand around a raw deflate stream
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    hdr = bytes([0x1f, 0x8b, 8, 4 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"extra" + b"a comment\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")
    trailer = zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")
    assert fq.gunzip(hdr + body + trailer).tobytes() == data
    assert fq.gunzip(hdr + body + trailer + members[0]).tobytes() == data + data[:20000]
    # the library call itself: members that do not fit are reported at a member boundary
    L = fq._native()
    comp = np.frombuffer(b"".join(members[:3]), np.uint8)
    out = np.empty(45000, np.uint8)
    ui, uo = C.c_int64(), C.c_int64()
    rc = L.itsx_gunzip(C.c_void_p(comp.ctypes.data), comp.size, C.c_void_p(out.ctypes.data), out.size, C.byref(ui), C.byref(uo))
    assert rc == 1 and uo.value == 40000 and ui.value == len(members[0]) + len(members[1])
    rc = L.itsx_gunzip(C.c_void_p(comp.ctypes.data), comp.size, C.c_void_p(out.ctypes.data), 1000, C.byref(ui), C.byref(uo))
    assert rc == 1 and ui.value == 0 and uo.value == 0
    # real files of the reference's test data
    for name in ("4774-1-MSITS3_R1.fastq.gz", "4774-1-MSITS3_R2.fastq.gz"):
        path = os.path.join(TD, name)
        with open(path, "rb") as f:
            assert fq.gunzip(f.read()).tobytes() == gzip.open(path, "rb").read()


def test_damaged_streams_raise_what_the_gzip_module_raises():
    data = _payloads()["fastq"]
    comp = bytearray(_member(data, 6))
    with pytest.raises(EOFError):
        fq.gunzip(bytes(comp[:len(comp) // 2]))                   # truncated
    bad = bytearray(comp)
    bad[-5] ^= 0xff                                               # CRC
    with pytest.raises(gzip.BadGzipFile):
        fq.gunzip(bytes(bad))
    bad = bytearray(comp)
    bad[len(bad) // 2] ^= 0x55                                    # somewhere in the deflate stream
    with pytest.raises((gzip.BadGzipFile, zlib.error, EOFError)):
        fq.gunzip(bytes(bad))
    with pytest.raises(gzip.BadGzipFile):
        fq.gunzip(b"this is not a gzip file at all, just text\n")


def test_reader_uses_the_native_gunzip(tmp_path):
    path = os.path.join(TD, "4774-1-MSITS3_R1.fastq.gz")
    b = fq.read_fastq(path)
    assert b.n == 250 and b.buf.tobytes() == gzip.open(path, "rb").read()
