"""Minimal zstd codec over the system libzstd (ctypes), standing in for the reference's `pyzstd`
dependency (itsxpress/SeqSample.py:8, used at :727-729, :770-773, :916-921, :933-940).  Whole-buffer
compress / decompress only -- that is all the FASTQ reader/writer needs."""
import ctypes as C

_L = None


def _lib():
    global _L
    if _L is None:
        try:
            L = C.CDLL("libzstd.so.1")
        except OSError as e:           # same failure class the reference shows when zstd support is missing
            raise ModuleNotFoundError("zstd support needs libzstd.so.1") from e
        L.ZSTD_compressBound.restype = C.c_size_t
        L.ZSTD_compressBound.argtypes = [C.c_size_t]
        L.ZSTD_compress.restype = C.c_size_t
        L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ZSTD_isError.restype = C.c_uint
        L.ZSTD_isError.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.restype = C.c_char_p
        L.ZSTD_getErrorName.argtypes = [C.c_size_t]
        L.ZSTD_createDStream.restype = C.c_void_p
        L.ZSTD_freeDStream.argtypes = [C.c_void_p]
        L.ZSTD_initDStream.argtypes = [C.c_void_p]
        L.ZSTD_initDStream.restype = C.c_size_t
        L.ZSTD_decompressStream.restype = C.c_size_t
        L.ZSTD_decompressStream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ZSTD_DStreamOutSize.restype = C.c_size_t
        _L = L
    return _L


class _Buf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def compress(data, level=3):
    L = _lib()
    data = bytes(data)
    cap = L.ZSTD_compressBound(len(data))
    out = C.create_string_buffer(cap)
    n = L.ZSTD_compress(out, cap, data, len(data), level)
    if L.ZSTD_isError(n):
        raise ValueError("zstd: " + L.ZSTD_getErrorName(n).decode())
    return out.raw[:n]


def decompress(data):
    """Streaming decode (handles multi-frame files and frames without a content size)."""
    L = _lib()
    data = bytes(data)
    ds = L.ZSTD_createDStream()
    L.ZSTD_initDStream(ds)
    try:
        src = C.create_string_buffer(data, len(data))
        ib = _Buf(C.cast(src, C.c_void_p), len(data), 0)
        osz = max(L.ZSTD_DStreamOutSize(), 1 << 20)
        dst = C.create_string_buffer(osz)
        parts = []
        rc, full = 0, False
        # keep calling while input remains OR the last call filled the output buffer: the decoder can hold a decoded
        # block back after it has consumed all of the input (multi-frame / pzstd / `zstd -T` files)
        while ib.pos < ib.size or full:
            ob = _Buf(C.cast(dst, C.c_void_p), osz, 0)
            rc = L.ZSTD_decompressStream(ds, C.byref(ob), C.byref(ib))
            if L.ZSTD_isError(rc):
                raise ValueError("zstd: " + L.ZSTD_getErrorName(rc).decode())
            parts.append(dst.raw[:ob.pos])
            full = ob.pos == osz
        if rc != 0:         # the last frame was not finished: pyzstd raises here too
            raise ValueError("zstd: truncated input")
        return b"".join(parts)
    finally:
        L.ZSTD_freeDStream(ds)


def decompress_stream(fh, block=1 << 20):
    """Generator over the decoded bytes of a zstd file object, read in blocks (multi-frame files included)."""
    L = _lib()
    ds = L.ZSTD_createDStream()
    L.ZSTD_initDStream(ds)
    try:
        osz = max(L.ZSTD_DStreamOutSize(), 1 << 20)
        dst = C.create_string_buffer(osz)
        rc = 0
        while True:
            raw = fh.read(block)
            if not raw:
                break
            src = C.create_string_buffer(raw, len(raw))
            ib = _Buf(C.cast(src, C.c_void_p), len(raw), 0)
            full = False
            while ib.pos < ib.size or full:
                ob = _Buf(C.cast(dst, C.c_void_p), osz, 0)
                rc = L.ZSTD_decompressStream(ds, C.byref(ob), C.byref(ib))
                if L.ZSTD_isError(rc):
                    raise ValueError("zstd: " + L.ZSTD_getErrorName(rc).decode())
                if ob.pos:
                    yield dst.raw[:ob.pos]
                full = ob.pos == osz
        if rc != 0:
            raise ValueError("zstd: truncated input")
    finally:
        L.ZSTD_freeDStream(ds)
