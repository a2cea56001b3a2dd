"""FASTQ ingest / emit in structure-of-arrays form.

Replaces the Biopython ``SeqIO.parse(..., "fastq")`` / ``SeqIO.write(..., "fastq")`` calls on the
reference's hot path (itsxpress/SeqSample.py:746-757, 912-945) with a vectorised numpy scanner that
produces the flat byte buffers the C-ABI consumes (sequence bytes + offsets, quality bytes, titles).
Semantics kept from Biopython: title = line after '@' right-stripped, id =
first whitespace-delimited token, '+' line may repeat the title, len(seq) == len(qual) and quality
characters in ASCII 33..126, otherwise ``ValueError``; output is ``@title\\nseq\\n+\\nqual\\n``.
"""
import gzip
import io
import os
import threading

import numpy as np


class FastqBatch:
    """All records of one FASTQ file.

    buf      : uint8 array, the decompressed file
    t_off/t_len : title start (after '@') and right-stripped length
    s_off/s_len : sequence start / length (q_off: quality start; same length)
    """

    __slots__ = ("buf", "t_off", "t_len", "s_off", "s_len", "q_off", "n", "_ids", "_labels")

    def __init__(self, buf, t_off, t_len, s_off, s_len, q_off):
        self.buf, self.t_off, self.t_len = buf, t_off, t_len
        self.s_off, self.s_len, self.q_off = s_off, s_len, q_off
        self.n = len(t_off)
        self._ids = None
        self._labels = None

    # -- flat views for the device path -------------------------------------------------------
    def seq_concat(self):
        """(bytes uint8[total], off int64[n+1]) with the sequences packed back to back."""
        return _gather(self.buf, self.s_off, self.s_len)

    def qual_concat(self):
        return _gather(self.buf, self.q_off, self.s_len)

    def title(self, i):
        o = int(self.t_off[i])
        return self.buf[o:o + int(self.t_len[i])].tobytes().decode("ascii", "replace")

    def labels(self):
        """(offset into buf int64[n], length int32[n]) of every record's id -- the first whitespace-delimited token of
        its title, Biopython's record.id -- found by the native scanner; cached."""
        if self._labels is None:
            lab_off, lab_len = np.empty(self.n, np.int64), np.empty(self.n, np.int32)
            if self.n:
                t_off = np.ascontiguousarray(self.t_off, dtype=np.int64)
                t_len = np.ascontiguousarray(self.t_len, dtype=np.int64)
                buf = np.ascontiguousarray(self.buf)
                if _native().itsx_fastq_labels(_vp(buf), _vp(t_off), _vp(t_len), self.n, _vp(lab_off), _vp(lab_len)) < 0:
                    raise ValueError("itsx_fastq_labels")
            self._labels = (lab_off, lab_len)
        return self._labels

    def ids(self):
        """First whitespace-delimited token of every title (Biopython's record.id) as str; cached."""
        if self._ids is None:
            lab_off, lab_len = self.labels()
            packed, off = _gather(self.buf, lab_off, lab_len)
            text = packed.tobytes().decode("ascii", "replace")          # one character per byte, valid or not
            o = off.tolist()
            self._ids = [text[o[k]:o[k + 1]] for k in range(self.n)]
        return self._ids

    def seq(self, i):
        o = int(self.s_off[i])
        return self.buf[o:o + int(self.s_len[i])].tobytes().decode("ascii")

    def qual(self, i):
        o = int(self.q_off[i])
        return self.buf[o:o + int(self.s_len[i])].tobytes().decode("ascii")


def _native():
    from . import _lib
    return _lib.lib()


def _vp(a):
    import ctypes
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _gather(buf, off, length):
    """Segments buf[off[i] : off[i] + length[i]] packed back to back (native, multi-threaded memcpy)."""
    n = len(off)
    off = np.ascontiguousarray(off, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    buf = np.ascontiguousarray(buf)
    out_off = np.zeros(n + 1, dtype=np.int64)
    if n == 0:
        return np.zeros(0, np.uint8), out_off
    L = _native()
    total = L.itsx_bytes_gather(_vp(buf), _vp(off), _vp(length), n, None, _vp(out_off))
    out = np.empty(total, np.uint8)
    L.itsx_bytes_gather(_vp(buf), _vp(off), _vp(length), n, _vp(out), _vp(out_off))
    return out, out_off


def _gather_numpy(buf, off, length):
    n = len(off)
    out_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(length, out=out_off[1:])
    total = int(out_off[-1])
    if n == 0:
        return np.zeros(0, np.uint8), out_off
    # index = repeat(start - out_start) + arange
    delta = np.repeat(off.astype(np.int64) - out_off[:-1], length)
    idx = delta + np.arange(total, dtype=np.int64)
    return buf[idx], out_off


def host_share():
    """Host cores this process may use: all of them divided by the ranks that share the box."""
    try:
        cores = len(os.sched_getaffinity(0))          # what this process may run on (containers, taskset)
    except (AttributeError, OSError):
        cores = os.cpu_count() or 2
    return max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))


class GzReader:
    """gzread-like reader over the bytes of a gzip file (csrc/inflate_host.cpp).  ``threads`` > 1 inflates a single
    deflate stream on several cores; every member's CRC-32 and ISIZE are checked."""

    def __init__(self, comp, threads=1, _tune=None):
        import ctypes as C
        self._a = comp if isinstance(comp, np.ndarray) else np.frombuffer(comp, dtype=np.uint8)     # keeps the input alive
        self._L = _native()
        self._h = self._L.itsx_gz_open(C.c_void_p(self._a.ctypes.data if self._a.size else 0), self._a.size, int(threads))
        if not self._h:
            raise MemoryError("itsx_gz_open")
        if _tune:
            self._L.itsx_gz_tune(self._h, *[int(v) for v in _tune])

    def readinto(self, out, at=0, hist=0):
        """Fill out[at:] (uint8 array); bytes written, 0 at the end of the stream, < 0 for a malformed stream.  ``hist``:
        how many bytes in front of out[at] are the output of the preceding calls (0 for a fresh buffer)."""
        import ctypes as C
        return int(self._L.itsx_gz_read(self._h, C.c_void_p(out.ctypes.data + at), out.size - at, int(hist)))

    def stat(self, what):
        return int(self._L.itsx_gz_stat(self._h, what))

    def close(self):
        if self._h:
            self._L.itsx_gz_close(self._h)
            self._h = None
        self._a = None

    __del__ = close


def gunzip(comp, threads=None, _tune=None):
    """A gzip file's bytes inflated by the native reader -> uint8 array.  ``threads`` None: this process's share of the
    host cores (one deflate stream is then inflated on several of them).  Anything the reader does not accept goes to
    the gzip module, which inflates it or raises what the reference would have raised (BadGzipFile, EOFError,
    zlib.error)."""
    a = comp if isinstance(comp, np.ndarray) else np.frombuffer(comp, dtype=np.uint8)
    n = a.size
    if n >= 18:
        isize = int.from_bytes(a[-4:].tobytes(), "little")
        # one member (every Casava / Illumina file): ISIZE is the answer unless it wrapped; several: start from a guess
        cap = isize + 64 if isize >= n // 2 else 4 * n + (1 << 16)
        out = np.empty(max(cap, 1 << 16), np.uint8)
        r = GzReader(a, host_share() if threads is None else threads, _tune)
        try:
            done = 0
            while True:
                if done == out.size:                        # grow; deflate cannot expand beyond 1032 : 1
                    if out.size > 1032 * n + (1 << 16):
                        break
                    grown = np.empty(out.size + max(out.size // 2, 1 << 20), np.uint8)
                    grown[:done] = out
                    out = grown
                k = r.readinto(out, done, done)
                if k < 0:
                    break
                if k == 0:
                    return out[:done]
                done += k
        finally:
            r.close()
    return np.frombuffer(gzip.decompress(bytes(comp)), dtype=np.uint8)


def _open_buffer(path, threads=None):
    """Decompressed content of ``path`` as a bytes-like object (a uint8 array for .gz: no copy behind the inflater).
    ``threads``: host cores the inflate of a .gz may use (None: this process's share)."""
    if path.endswith(".gz"):
        import mmap
        with open(path, "rb") as f:
            if os.fstat(f.fileno()).st_size < (1 << 20):
                return gunzip(f.read(), threads)
            # mapped, not read: the inflating threads pull the pages in themselves (the map lives only for this call)
            try:
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            except (OSError, ValueError):          # a file system that cannot map: read it
                return gunzip(f.read(), threads)
            try:
                a = np.frombuffer(mm, dtype=np.uint8)
                try:
                    return gunzip(a, threads)
                finally:
                    del a
            finally:
                try:
                    mm.close()
                except BufferError:          # (a reference to the view survived somewhere: the GC unmaps it later)
                    pass
    if path.endswith(".zst"):
        from . import _zstd
        with open(path, "rb") as f:
            return _zstd.decompress(f.read(), as_array=True)
    with open(path, "rb") as f:
        return f.read()


def _open_bytes(path):
    data = _open_buffer(path)
    return data if isinstance(data, bytes) else data.tobytes()


def parse_bytes(data):
    """Parse a whole (decompressed) 4-line FASTQ byte string into a FastqBatch (native scanner,
    csrc/fastq_host.cpp); ValueError on malformed input with Biopython's wording."""
    buf = np.frombuffer(data, dtype=np.uint8)
    z = np.zeros(0, np.int64)
    if buf.size == 0:
        return FastqBatch(buf, z, z, z, z, z)
    L = _native()
    n = L.itsx_fastq_index(_vp(buf), buf.size, 0, None, None, None, None, None)
    if n < 0:
        raise ValueError(L.itsx_host_last_error().decode())
    t_off, s_off, q_off = (np.empty(n, np.int64) for _ in range(3))
    t_len, s_len = (np.empty(n, np.int32) for _ in range(2))
    if n:
        rc = L.itsx_fastq_index(_vp(buf), buf.size, n, _vp(t_off), _vp(t_len), _vp(s_off), _vp(s_len), _vp(q_off))
        if rc < 0:
            raise ValueError(L.itsx_host_last_error().decode())
    return FastqBatch(buf, t_off, t_len.astype(np.int64), s_off, s_len.astype(np.int64), q_off)


def _parse_bytes_numpy(data):
    """The same scanner in numpy (kept as an independent cross-check of the native one in the tests)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    if buf.size == 0:
        z = np.zeros(0, np.int64)
        return FastqBatch(buf, z, z, z, z, z)
    nl = np.flatnonzero(buf == 10)
    if buf[-1] != 10:
        nl = np.append(nl, buf.size)
    starts = np.empty(len(nl), dtype=np.int64)
    starts[0] = 0
    starts[1:] = nl[:-1] + 1
    ends = nl.astype(np.int64)
    # strip trailing '\r'
    cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
    ends = ends - cr
    # drop trailing blank lines
    nlines = len(starts)
    while nlines > 0 and ends[nlines - 1] == starts[nlines - 1]:
        nlines -= 1
    starts, ends = starts[:nlines], ends[:nlines]
    if nlines % 4 != 0:
        raise ValueError("FASTQ is truncated or not in 4-line format")
    t0, s0, p0, q0 = starts[0::4], starts[1::4], starts[2::4], starts[3::4]
    t1, s1, p1, q1 = ends[0::4], ends[1::4], ends[2::4], ends[3::4]
    if np.any(t1 == t0) or np.any(buf[t0] != ord("@")):
        raise ValueError("Records in Fastq files should start with '@' character")
    if np.any(p1 == p0) or np.any(buf[p0] != ord("+")):
        raise ValueError("Expected '+' line in FASTQ record")
    slen = s1 - s0
    if np.any((q1 - q0) != slen):
        raise ValueError("Lengths of sequence and quality values differs")
    # '+' line may carry the title again; it must then be identical
    rep = np.flatnonzero((p1 - p0) > 1)
    for i in rep.tolist():
        if buf[p0[i] + 1:p1[i]].tobytes().rstrip() != buf[t0[i] + 1:t1[i]].tobytes().rstrip():
            raise ValueError("Sequence and quality captions differ.")
    # right-strip titles (Biopython strips the title line)
    tl = t1 - (t0 + 1)
    if len(t0):
        last = buf[np.maximum(t1 - 1, 0)]
        ws = (tl > 0) & ((last == 32) | (last == 9))
        while np.any(ws):
            tl = tl - ws
            last = buf[np.maximum(t0 + tl, 0)]
            ws = (tl > 0) & ((last == 32) | (last == 9))
    # quality range check (ASCII 33..126)
    batch = FastqBatch(buf, t0 + 1, tl, s0, slen, q0)
    if len(t0):
        q, _ = _gather_numpy(buf, q0, slen)
        if q.size and (q.min() < 33 or q.max() > 126):
            raise ValueError("Invalid character in quality string")
    return batch


# record counts of files this process has just scanned or written, keyed by path and validated against size + mtime:
# lets the closing "Total number of reads in file ..." lines (main.py:417-438 upstream) skip a second pass over
# gigabytes it has already counted
_COUNTS = {}


def note_count(path, n):
    try:
        st = os.stat(path)
        _COUNTS[os.path.abspath(path)] = (st.st_size, st.st_mtime_ns, int(n))
    except OSError:
        pass


def cached_count(path):
    """Record count of ``path`` if this process scanned / wrote exactly this file (same size and mtime), else None."""
    hit = _COUNTS.get(os.path.abspath(path))
    if hit is None:
        return None
    try:
        st = os.stat(path)
    except OSError:
        return None
    return hit[2] if (st.st_size, st.st_mtime_ns) == hit[:2] else None


def _read_fastq_now(path, threads=None):
    batch = parse_bytes(_open_buffer(path, threads))
    note_count(path, batch.n)
    return batch


# how many samples a multi-sample driver reads ahead (two files each, one thread per file: a single-member .gz inflates on
# one core at ~200 MB/s, so one sample ahead leaves the other cores idle and the inflate bounds the artifact); the
# default leaves about half the cores of the process's share to it, ITSX_READ_AHEAD overrides
def _default_read_ahead():
    cores = os.cpu_count() or 2
    share = max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))
    return max(1, min(4, share // 4))


READ_AHEAD = int(os.environ.get("ITSX_READ_AHEAD", "0")) or _default_read_ahead()
# files being read ahead on background threads: abspath -> (Future[FastqBatch], size, mtime_ns)
_PREFETCH = {}
_PREFETCH_POOL = None


def prefetch(paths):
    """Start reading / inflating / scanning these files in the background (zlib and the native scanner release the
    GIL).  A multi-sample driver calls it with the NEXT sample's files while the current sample is on the GPU;
    ``read_fastq`` then picks the result up.  Errors are not raised here: a failed read-ahead is dropped and the
    ordinary read raises in the caller's context."""
    global _PREFETCH_POOL
    from concurrent.futures import ThreadPoolExecutor
    if _PREFETCH_POOL is None:
        _PREFETCH_POOL = ThreadPoolExecutor(2 * READ_AHEAD, thread_name_prefix="itsx-readahead")
    for p in paths:
        if not p:
            continue
        key = os.path.abspath(p)
        if key in _PREFETCH:
            continue
        try:
            st = os.stat(p)
        except OSError:
            continue
        # 2 x READ_AHEAD files are in flight: each inflates on its part of the cores (one core each on a small share, where
        # the single-core decoder does the same work in half the core-seconds)
        _PREFETCH[key] = (_PREFETCH_POOL.submit(_read_fastq_now, p, max(1, host_share() // (2 * READ_AHEAD))), st.st_size, st.st_mtime_ns)


def drop_prefetched():
    """Forget read-ahead results nobody collected (a driver calls it when it is done or gives up)."""
    for fut, _, _ in list(_PREFETCH.values()):
        fut.cancel()
    _PREFETCH.clear()


def read_fastq(path, threads=None):
    """The file as a FastqBatch.  ``threads``: host cores for the inflate of a .gz (None: this process's share, or the
    read-ahead's per-file part of it while other files are being read ahead)."""
    if threads is None and _PREFETCH:
        threads = max(1, host_share() // (2 * READ_AHEAD))
    hit = _PREFETCH.pop(os.path.abspath(path), None)
    if hit is not None:
        fut, size, mtime = hit
        try:
            st = os.stat(path)
            if (st.st_size, st.st_mtime_ns) == (size, mtime):
                return fut.result()
        except Exception:
            pass                        # fall through: the ordinary read reports the problem where it is expected
    return _read_fastq_now(path, threads)


def read_fastq_many(paths):
    """The files read, decompressed and scanned concurrently (the native inflater and scanner release the GIL), each
    on its part of this process's cores: R1 and R2 of a pair arrive in the time of the slower one."""
    paths = list(paths)
    if len(paths) < 2:
        return [read_fastq(p) for p in paths]
    from concurrent.futures import ThreadPoolExecutor
    per_file = None if _PREFETCH else max(1, host_share() // len(paths))
    with ThreadPoolExecutor(len(paths)) as ex:
        return list(ex.map(lambda p: read_fastq(p, per_file), paths))


def write_records(fh, batch, keep_idx, lo, hi):
    """Write records ``keep_idx`` of ``batch`` sliced to [lo:hi) as 4-line FASTQ to binary handle fh."""
    fh.write(format_records(batch, keep_idx, lo, hi))


def format_records(batch, keep_idx, lo, hi):
    """Vectorised FASTQ text assembly: returns bytes of '@title\\nseq\\n+\\nqual\\n' for each kept record."""
    keep_idx = np.asarray(keep_idx, dtype=np.int64)
    n = len(keep_idx)
    if n == 0:
        return b""
    lo = np.asarray(lo, dtype=np.int64)
    hi = np.asarray(hi, dtype=np.int64)
    tl = batch.t_len[keep_idx].astype(np.int64)
    sl = hi - lo
    rec_len = 1 + tl + 1 + sl + 1 + 2 + sl + 1
    rec_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rec_len, out=rec_off[1:])
    out = np.empty(int(rec_off[-1]), dtype=np.uint8)
    base = rec_off[:-1]
    out[base] = ord("@")
    _scatter(out, base + 1, batch.buf, batch.t_off[keep_idx], tl)
    p = base + 1 + tl
    out[p] = 10
    _scatter(out, p + 1, batch.buf, batch.s_off[keep_idx] + lo, sl)
    p = p + 1 + sl
    out[p] = 10
    out[p + 1] = ord("+")
    out[p + 2] = 10
    _scatter(out, p + 3, batch.buf, batch.q_off[keep_idx] + lo, sl)
    out[p + 3 + sl] = 10
    return out.tobytes()


def _scatter(dst, dst_off, src, src_off, length):
    total = int(length.sum())
    if total == 0:
        return
    ar = np.arange(total, dtype=np.int64)
    cs = np.zeros(len(length) + 1, dtype=np.int64)
    np.cumsum(length, out=cs[1:])
    within = ar - np.repeat(cs[:-1], length)
    dst[np.repeat(dst_off, length) + within] = src[np.repeat(src_off.astype(np.int64), length) + within]


def format_gathered(batch, keep_idx, out_off, out_seq, out_qual, prefix=None, suffix=None, as_array=False):
    """FASTQ text from slices already gathered back to back on the device (itsx_trim_gather):
    record t = title of batch[keep_idx[t]], bases out_seq[out_off[t]:out_off[t+1]], same for qualities.
    prefix / suffix: (bases, quals) byte strings stitched to every record (--trim-ccs, SeqSample.py:601-622).
    Native, multi-threaded (csrc/fastq_host.cpp).  ``as_array``: the uint8 array the formatter filled instead of a bytes
    copy of it (the writers, the GPU gzip stage and the scanner take either)."""
    n = len(keep_idx)
    if n == 0:
        return np.zeros(0, np.uint8) if as_array else b""
    ki = np.ascontiguousarray(keep_idx, dtype=np.int32)
    oo = np.ascontiguousarray(out_off, dtype=np.int64)
    os_ = np.ascontiguousarray(out_seq, dtype=np.uint8)
    oq = np.ascontiguousarray(out_qual, dtype=np.uint8)
    t_off = np.ascontiguousarray(batch.t_off, dtype=np.int64)
    t_len = np.ascontiguousarray(batch.t_len, dtype=np.int32)
    buf = np.ascontiguousarray(batch.buf)
    pre_s, pre_q = prefix if prefix else (b"", b"")
    suf_s, suf_q = suffix if suffix else (b"", b"")
    L = _native()
    args = [_vp(buf), _vp(t_off), _vp(t_len), _vp(ki), n, _vp(oo), _vp(os_), _vp(oq), pre_s, pre_q, len(pre_s),
            suf_s, suf_q, len(suf_s)]
    total = L.itsx_fastq_format(*args, None)
    dst = np.empty(total, np.uint8)
    L.itsx_fastq_format(*args, _vp(dst))
    return dst if as_array else dst.tobytes()


def batch_of_gathered(data, batch, keep_idx, out_off):
    """The FastqBatch ``parse_bytes`` would build from ``data = format_gathered(batch, keep_idx, out_off, ...)``
    (no prefix / suffix), computed from the record lengths instead of scanning the text again."""
    keep_idx = np.asarray(keep_idx, dtype=np.int64)
    tl = np.asarray(batch.t_len, dtype=np.int64)[keep_idx]
    sl = np.diff(np.asarray(out_off, dtype=np.int64))
    base = np.zeros(len(keep_idx), np.int64)
    if len(keep_idx) > 1:
        np.cumsum((tl + 2 * sl + 6)[:-1], out=base[1:])            # '@' title '\n' seq '\n+\n' qual '\n'
    s_off = base + 2 + tl
    return FastqBatch(np.frombuffer(data, dtype=np.uint8), base + 1, tl, s_off, sl, s_off + sl + 3)


def _format_gathered_numpy(batch, keep_idx, out_off, out_seq, out_qual, prefix=None, suffix=None):
    """numpy version of format_gathered (cross-check in the tests)."""
    keep_idx = np.asarray(keep_idx, dtype=np.int64)
    n = len(keep_idx)
    if n == 0:
        return b""
    out_off = np.asarray(out_off, dtype=np.int64)
    pre_s, pre_q = prefix if prefix else (b"", b"")
    suf_s, suf_q = suffix if suffix else (b"", b"")
    lp, ls = len(pre_s), len(suf_s)
    tl = batch.t_len[keep_idx].astype(np.int64)
    sl = out_off[1:] - out_off[:-1]
    full = sl + lp + ls
    rec_len = 1 + tl + 1 + full + 1 + 2 + full + 1
    rec_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rec_len, out=rec_off[1:])
    out = np.empty(int(rec_off[-1]), dtype=np.uint8)
    base = rec_off[:-1]
    out[base] = ord("@")
    _scatter(out, base + 1, batch.buf, batch.t_off[keep_idx], tl)
    p = base + 1 + tl
    out[p] = 10
    p = p + 1
    for src, pre, suf in ((out_seq, pre_s, suf_s), (out_qual, pre_q, suf_q)):
        if lp:
            out[(p[:, None] + np.arange(lp)[None, :]).ravel()] = np.tile(np.frombuffer(pre, np.uint8), n)
        _scatter(out, p + lp, src, out_off[:-1], sl)
        if ls:
            out[((p + lp + sl)[:, None] + np.arange(ls)[None, :]).ravel()] = np.tile(np.frombuffer(suf, np.uint8), n)
        q = p + full
        out[q] = 10
        if src is out_seq:
            out[q + 1] = ord("+")
            out[q + 2] = 10
            p = q + 3
    return out.tobytes()


class Record:
    """The slice of Biopython's SeqRecord that the reference's generators touch (SeqSample.py:586-670,
    814-864): ``id``, ``name``, ``description``, ``seq`` (a str), ``letter_annotations['phred_quality']``,
    ``len()``, slicing, and ``format('fastq')``."""

    __slots__ = ("id", "name", "description", "seq", "letter_annotations")

    def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>",
                 quals=None):
        self.seq = str(seq)
        self.id, self.name, self.description = id, name, description
        self.letter_annotations = {}
        if quals is not None:
            self.letter_annotations["phred_quality"] = list(quals)

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, idx):
        if not isinstance(idx, slice):
            return self.seq[idx]
        r = Record(self.seq[idx], self.id, self.name, self.description)
        if "phred_quality" in self.letter_annotations:
            r.letter_annotations["phred_quality"] = self.letter_annotations["phred_quality"][idx]
        return r

    def format(self, fmt="fastq"):
        if fmt != "fastq":
            raise ValueError("only the fastq format is supported")
        q = self.letter_annotations.get("phred_quality")
        if q is None:
            raise ValueError("No suitable quality scores found in letter_annotations of SeqRecord")
        title = self.description if self.description and self.description.split(None, 1)[0] == self.id else \
            (self.id if not self.description or self.description == "<unknown description>" else
             "%s %s" % (self.id, self.description))
        return "@%s\n%s\n+\n%s\n" % (title, self.seq, "".join(chr(min(int(v), 93) + 33) for v in q))


def iter_records(source):
    """Records of a FASTQ path (plain / .gz / .zst), text handle or bytes, like SeqIO.parse(.., 'fastq')."""
    if isinstance(source, (bytes, bytearray)):
        data = bytes(source)
    elif isinstance(source, str):
        data = _open_bytes(source)
    else:
        data = source.read()
        if isinstance(data, str):
            data = data.encode()
    b = parse_bytes(data)
    for i in range(b.n):
        title = b.title(i)
        rid = title.split(None, 1)[0] if title.split() else ""
        yield Record(b.seq(i), id=rid, name=rid, description=title,
                     quals=[c - 33 for c in b.buf[int(b.q_off[i]):int(b.q_off[i]) + int(b.s_len[i])].tolist()])


GZIP_LEVEL = int(os.environ.get("ITSX_GZIP_LEVEL", "6"))     # upstream writes level 9 on one core (gzip.open default)
# who compresses .gz output: "gpu" = itsx_gzip_compress (deflate.cu; a context of its own, so that a writer thread
# never shares a stream with the search), "host" = zlib members on every host core.  The inflated bytes are the same.
GZIP_BACKEND = os.environ.get("ITSX_GZIP", "gpu")
_GZ_CTX = None
_GZ_LOCK = threading.Lock()


def gzip_members(data):
    """``data`` as a valid multi-member gzip stream (a buffer file.write() accepts)."""
    global _GZ_CTX
    if GZIP_BACKEND == "gpu":
        from . import _lib
        with _GZ_LOCK:
            if _GZ_CTX is None:
                _GZ_CTX = _lib.Context(int(os.environ.get("ITSX_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
            return _GZ_CTX.gzip_compress(data)
    if GZIP_BACKEND != "host":
        raise ValueError("ITSX_GZIP must be 'gpu' or 'host', not %r" % (GZIP_BACKEND,))
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def member(chunk):
        co = zlib.compressobj(GZIP_LEVEL, zlib.DEFLATED, 31)
        return co.compress(chunk) + co.flush()
    step = 4 << 20
    chunks = [data[i:i + step] for i in range(0, len(data), step)] or [b""]
    with ThreadPoolExecutor(max(1, os.cpu_count() or 1)) as ex:
        return b"".join(ex.map(member, chunks))


def write_compressed(path, data, gzipped=False, zstd_file=False, threads=None, n_records=None):
    """Write FASTQ bytes plain, as gzip (independent members -- a valid gzip stream whose DECOMPRESSED bytes are what
    parity is judged on; gzip headers carry mtime, so compressed bytes are not reproducible even reference-vs-reference)
    or as one zstd frame.  The members come from the GPU (gzip_members); on the host zlib level 6 runs at ~9 MB/s per
    core on amplicon FASTQ (level 9: ~4 MB/s, 1.5 % smaller), which is what used to bound every .gz output."""
    if gzipped:
        with open(path, "wb") as f:
            f.write(gzip_members(data))
    elif zstd_file:
        from . import _zstd
        with open(path, "wb") as f:
            f.write(_zstd.compress(data))
    else:
        with open(path, "wb") as f:
            f.write(data)
    if n_records is not None:
        note_count(path, n_records)


# ---- chunked reader / writer: a file never sits whole in host memory -------------------------------
STREAM_CHUNK_BYTES = int(os.environ.get("ITSX_STREAM_CHUNK_BYTES", str(256 << 20)))


def _gz_blocks_zlib(path, block, skip=0):
    """Inflated blocks of a gzip file through zlib (multi-member aware, zero padding between / behind members skipped as
    the gzip module does), the first ``skip`` bytes left out."""
    import zlib
    with open(path, "rb") as f:
        d, fresh = zlib.decompressobj(31), True
        while True:
            raw = f.read(max(block // 4, 1 << 16))
            if not raw:
                if not fresh and not d.eof:
                    raise EOFError("Compressed file ended before the end-of-stream marker was reached")
                return
            while raw:
                if fresh:
                    raw = raw.lstrip(b"\0")
                    if not raw:
                        break
                    fresh = False
                out = d.decompress(raw, block)          # at most `block` bytes per call: memory stays bounded
                if out:
                    if skip >= len(out):
                        skip -= len(out)
                    else:
                        yield out[skip:]
                        skip = 0
                if d.eof:                               # next gzip member
                    raw = d.unused_data
                    d, fresh = zlib.decompressobj(31), True
                else:
                    raw = d.unconsumed_tail


def _gz_blocks(path, block):
    """Inflated blocks of about ``block`` bytes of a gzip file: the file is mapped (never read whole) and inflated by
    the native reader on this process's share of the cores.  A stream the reader does not accept is handed to zlib from
    where the reader stopped, which inflates it or raises what the reference would have raised."""
    import mmap
    size = os.path.getsize(path)
    if size == 0:
        return
    done = 0
    with open(path, "rb") as f:
        try:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        except (OSError, ValueError):              # a file system that cannot map: zlib streams it
            mm = None
        if mm is None:
            yield from _gz_blocks_zlib(path, block)
            return
        try:
            a = np.frombuffer(mm, dtype=np.uint8)
            r = GzReader(a, host_share())
            try:
                while True:
                    out = np.empty(max(int(block), 1), np.uint8)
                    k = r.readinto(out)
                    if k < 0:
                        break
                    if k == 0:
                        return
                    done += k
                    yield out[:k].tobytes() if k < out.size else out.tobytes()
            finally:
                r.close()
                del a
        finally:
            try:
                mm.close()
            except BufferError:
                pass
    yield from _gz_blocks_zlib(path, block, skip=done)


def _raw_blocks(path, block):
    """Decompressed bytes of ``path`` in blocks of about ``block`` bytes (plain, multi-member gzip, zstd)."""
    if path.endswith(".gz"):
        yield from _gz_blocks(path, block)
    elif path.endswith(".zst"):
        from . import _zstd
        with open(path, "rb") as f:
            for out in _zstd.decompress_stream(f, max(block // 4, 1 << 16)):
                yield out
    else:
        with open(path, "rb") as f:
            while True:
                raw = f.read(block)
                if not raw:
                    return
                yield raw


def stream_fastq(path, chunk_bytes=None):
    """Yield FastqBatch chunks of ``path`` in file order, each holding whole records and about ``chunk_bytes`` of
    text.  Record boundaries are found by COUNTING lines from the start of the file (a quality line may begin with
    '@', so no local pattern identifies a title line); the next block is read / inflated on a background thread while
    the caller works on the current chunk."""
    from concurrent.futures import ThreadPoolExecutor
    chunk_bytes = chunk_bytes or STREAM_CHUNK_BYTES
    blocks = _raw_blocks(path, chunk_bytes)
    parts, have = [], 0                    # text not yet handed out: the carry of the last cut + the blocks since
    pool = ThreadPoolExecutor(1, thread_name_prefix="itsx-stream")
    try:
        nxt = pool.submit(next, blocks, None)
        while True:
            raw = nxt.result()
            if raw is not None:
                nxt = pool.submit(next, blocks, None)
                parts.append(raw)
                have += len(raw)
                if have < chunk_bytes:
                    continue
            data = parts[0] if len(parts) == 1 else b"".join(parts)
            if raw is None:
                if data.strip():
                    yield parse_bytes(data)
                return
            # cut behind the last line whose number is a multiple of four (lines counted on the host threads)
            buf = np.frombuffer(data, np.uint8)
            cut = int(_native().itsx_fastq_cut(_vp(buf), buf.size))
            if cut == 0:
                parts, have = [data], len(data)
                continue
            parts, have = [data[cut:]], len(data) - cut
            yield parse_bytes(memoryview(data)[:cut])
    finally:
        pool.shutdown(wait=False, cancel_futures=True)


class ChunkWriter:
    """Appends FASTQ text chunk by chunk: plain, gzip (every chunk becomes independent members compressed on all cores:
    a valid multi-member stream) or zstd (one frame per chunk: a valid multi-frame stream)."""

    def __init__(self, path, gzipped=False, zstd_file=False):
        self.path, self.gz, self.zst = path, gzipped, zstd_file
        self.f = open(path, "wb")
        self.n = 0

    def write(self, data, n_records=0):
        self.n += int(n_records)
        if len(data) == 0:
            return
        if self.gz:
            self.f.write(gzip_members(data))
        elif self.zst:
            from . import _zstd
            self.f.write(_zstd.compress(data))
        else:
            self.f.write(data)

    def close(self):
        if self.gz and self.f.tell() == 0:
            self.f.write(gzip_members(b""))          # an empty but valid gzip file
        self.f.close()
        note_count(self.path, self.n)
