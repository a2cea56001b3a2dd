"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.
Restates: vsearch --fastx_uniques (SeqSample.py:106-116), hmmsearch (SeqSample.py:191-209),
ItsPosition (SeqSample.py:380-498) and the trim rules (SeqSample.py:564-884).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OraPair(C.Structure):
    _fields_ = [
        ("msv_xJ", C.c_int32), ("msv_overflow", C.c_int32), ("usc", C.c_float), ("nullsc", C.c_float),
        ("filtersc", C.c_float), ("fwdsc", C.c_float), ("P_msv", C.c_double), ("P_bias", C.c_double),
        ("P_fwd", C.c_double), ("pass_msv", C.c_int32), ("pass_bias", C.c_int32), ("pass_fwd", C.c_int32),
        ("bcksc", C.c_float), ("nregions", C.c_int32), ("nmultidomain", C.c_int32), ("ndom", C.c_int32),
        ("reported", C.c_int32), ("seq_score", C.c_float), ("pre_score", C.c_float), ("lnP", C.c_double),
    ]


DOM_DTYPE = np.dtype([
    ("seq", "<i4"), ("prof", "<i4"), ("ienv", "<i4"), ("jenv", "<i4"), ("envsc", "<f4"),
    ("domcorrection", "<f4"), ("bitscore", "<f4"), ("dombias", "<f4"), ("lnP", "<f8"),
    ("dom_idx", "<i4"), ("is_multidomain", "<i4"), ("is_reported", "<i4"), ("pad", "<i4"),
])


class OraDom(C.Structure):
    _fields_ = [
        ("seq", C.c_int32), ("prof", C.c_int32), ("ienv", C.c_int32), ("jenv", C.c_int32),
        ("envsc", C.c_float), ("domcorrection", C.c_float), ("bitscore", C.c_float), ("dombias", C.c_float),
        ("lnP", C.c_double), ("dom_idx", C.c_int32), ("is_multidomain", C.c_int32),
        ("is_reported", C.c_int32), ("pad", C.c_int32),
    ]


class OraParams(C.Structure):
    _fields_ = [("T", C.c_float), ("F1", C.c_double), ("F2", C.c_double), ("F3", C.c_double),
                ("domE", C.c_double), ("nthreads", C.c_int), ("resolve_multidomain", C.c_int)]


class OraStats(C.Structure):
    _fields_ = [("npairs_total", C.c_int64), ("n_past_msv", C.c_int64), ("n_past_bias", C.c_int64),
                ("n_past_fwd", C.c_int64), ("n_reported_pairs", C.c_int64),
                ("n_multidomain_regions", C.c_int64), ("ndom_total", C.c_int64),
                ("ndom_reported", C.c_int64), ("msv_cells", C.c_double), ("fwd_cells", C.c_double),
                ("bck_cells", C.c_double), ("env_cells", C.c_double)]


class OraMergeParams(C.Structure):
    _fields_ = [("maxdiffs", C.c_int32), ("allow_stagger", C.c_int32), ("qmax", C.c_int32), ("minovlen", C.c_int32),
                ("qmaxout", C.c_int32), ("qminout", C.c_int32), ("ascii", C.c_int32), ("pad", C.c_int32),
                ("maxee", C.c_double), ("maxdiffpct", C.c_double)]


MERGE_REASONS = ("ok", "repeat", "staggered", "maxdiffs", "maxdiffpct", "nokmers", "minscore", "minovlen", "maxee",
                 "badqual")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ora_hmm.c", "ora_derep.c", "ora_merge.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    L.ora_db_load.restype = vp
    L.ora_db_load.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
    L.ora_db_append.restype = C.c_int
    L.ora_db_append.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
    L.ora_db_free.argtypes = [vp]
    L.ora_db_count.argtypes = [vp]
    L.ora_db_name.restype = C.c_char_p
    L.ora_db_name.argtypes = [vp, C.c_int]
    L.ora_db_M.argtypes = [vp, C.c_int]
    L.ora_db_evparam.argtypes = [vp, C.c_int, vp]
    L.ora_db_raw.argtypes = [vp, C.c_int, vp, vp, vp]
    L.ora_db_msv.argtypes = [vp, C.c_int, vp, vp]
    L.ora_digitize.argtypes = [C.c_char_p, i64, vp]
    L.ora_default_params.argtypes = [C.POINTER(OraParams)]
    L.ora_pair_run.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(OraParams), C.POINTER(OraPair), vp, C.c_int]
    L.ora_forward_parser.argtypes = [vp, C.c_int, vp, C.c_int] + [vp] * 7
    L.ora_domain_decoding.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp]
    for fn in ("ora_msv_score", "ora_viterbi_filter"):
        getattr(L, fn).restype = f32
        getattr(L, fn).argtypes = [vp, C.c_int, vp, C.c_int, vp]
    for fn in ("ora_forward_score", "ora_backward_score", "ora_bias_filtersc"):
        getattr(L, fn).restype = f32
        getattr(L, fn).argtypes = [vp, C.c_int, vp, C.c_int]
    L.ora_nullsc.restype = f32
    L.ora_nullsc.argtypes = [C.c_int]
    L.ora_flogsum.restype = f32
    L.ora_flogsum.argtypes = [f32, f32]
    L.ora_search.restype = i64
    L.ora_search.argtypes = [vp, vp, vp, i64, C.POINTER(OraParams), C.POINTER(vp), vp, C.POINTER(OraStats)]
    L.ora_free.argtypes = [vp]
    L.ora_itspos.argtypes = [vp, i64, vp, vp, i64] + [vp] * 9
    L.ora_score10.restype = i32
    L.ora_score10.argtypes = [f32]
    L.ora_derep.restype = i64
    L.ora_derep.argtypes = [vp, vp, i64, vp, vp]
    L.ora_trim_bounds.restype = i64
    L.ora_trim_bounds.argtypes = [vp, i64, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.ora_merge_default_params.argtypes = [C.POINTER(OraMergeParams)]
    L.ora_merge_pairs.restype = i64
    L.ora_merge_pairs.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.POINTER(OraMergeParams), vp, vp, vp, vp, C.c_int]
    L.ora_merge_tables.argtypes = [C.POINTER(OraMergeParams), vp, vp, vp, vp, vp]
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def digitize(seq):
    """ASCII (bytes/str) -> uint8 residue codes; raises ValueError on illegal characters."""
    if isinstance(seq, str):
        seq = seq.encode()
    out = np.empty(len(seq), dtype=np.uint8)
    bad = lib().ora_digitize(seq, len(seq), _ptr(out))
    if bad:
        raise ValueError("illegal residue character in sequence")
    return out


class ProfileDB:
    """Profiles selected by name prefix from one or more HMMER3/f files, in file order
    (= create_runtime_hmm, main.py:176-231)."""

    def __init__(self, paths, prefixes=None):
        if isinstance(paths, (str, bytes)):
            paths = [paths]
        L = lib()
        self._h = L.ora_db_load(None, None, 0)
        pre = [p.encode() for p in (prefixes or [])]
        arr = (C.c_char_p * max(1, len(pre)))(*pre) if pre else None
        for p in paths:
            if not os.path.exists(p):
                continue
            if L.ora_db_append(self._h, p.encode(), arr, len(pre)) < 0:
                raise IOError("cannot read " + p)
        self.n = L.ora_db_count(self._h)
        self.names = [L.ora_db_name(self._h, i).decode() for i in range(self.n)]
        self.M = [L.ora_db_M(self._h, i) for i in range(self.n)]

    def __del__(self):
        try:
            lib().ora_db_free(self._h)
        except Exception:
            pass

    def evparam(self, p):
        out = np.zeros(6, dtype=np.float32)
        lib().ora_db_evparam(self._h, p, _ptr(out))
        return out

    def raw(self, p):
        M = self.M[p]
        mat = np.zeros((M + 1, 4), np.float32)
        t = np.zeros((M + 1, 7), np.float32)
        compo = np.zeros(4, np.float32)
        lib().ora_db_raw(self._h, p, _ptr(mat), _ptr(t), _ptr(compo))
        return mat, t, compo

    def msv_profile(self, p):
        M = self.M[p]
        cost = np.zeros((M + 1, 16), np.uint8)
        sc = np.zeros(4, np.int32)
        lib().ora_db_msv(self._h, p, _ptr(cost), _ptr(sc))
        return cost, dict(bias=int(sc[0]), base=int(sc[1]), tbm=int(sc[2]), tec=int(sc[3]))

    # --- single-pair stage functions -----------------------------------------
    def msv_score(self, p, dsq):
        ov = C.c_int(0)
        sc = lib().ora_msv_score(self._h, p, _ptr(dsq), len(dsq), C.byref(ov))
        return sc, bool(ov.value)

    def viterbi_filter(self, p, dsq):
        """16-bit Viterbi filter score in nats and its overflow flag (not on the reference's path: F1 == F2)."""
        ov = C.c_int()
        sc = lib().ora_viterbi_filter(self._h, p, _ptr(dsq), len(dsq), C.byref(ov))
        return float(sc), bool(ov.value)

    def forward_score(self, p, dsq):
        return lib().ora_forward_score(self._h, p, _ptr(dsq), len(dsq))

    def backward_score(self, p, dsq):
        return lib().ora_backward_score(self._h, p, _ptr(dsq), len(dsq))

    def bias_filtersc(self, p, dsq):
        return lib().ora_bias_filtersc(self._h, p, _ptr(dsq), len(dsq))

    def forward_parser(self, p, dsq):
        n = len(dsq) + 1
        arrs = [np.zeros(n, np.float32) for _ in range(6)]
        sc = np.zeros(1, np.float32)
        lib().ora_forward_parser(self._h, p, _ptr(dsq), len(dsq), *[_ptr(a) for a in arrs], _ptr(sc))
        return dict(zip("ENJBCS", arrs)), float(sc[0])

    def domain_decoding(self, p, dsq):
        n = len(dsq) + 1
        b, e, m = (np.zeros(n, np.float32) for _ in range(3))
        lib().ora_domain_decoding(self._h, p, _ptr(dsq), len(dsq), _ptr(b), _ptr(e), _ptr(m))
        return b, e, m

    def pair_run(self, p, dsq, params=None, domcap=64):
        prm = params or default_params()
        pr = OraPair()
        doms = np.zeros(domcap, dtype=DOM_DTYPE)
        n = lib().ora_pair_run(self._h, p, _ptr(dsq), len(dsq), C.byref(prm), C.byref(pr), _ptr(doms), domcap)
        return pr, doms[:n].copy()

    # --- whole search -----------------------------------------------------------
    def search(self, codes, off, params=None):
        """codes: uint8 concatenated residue codes; off: int64[nseq+1].  Returns (rows, nreported, stats)."""
        prm = params or default_params()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        nseq = len(off) - 1
        rows_p = C.c_void_p()
        nrep = np.zeros(max(1, self.n), np.int32)
        st = OraStats()
        n = lib().ora_search(self._h, _ptr(codes), _ptr(off), nseq, C.byref(prm), C.byref(rows_p), _ptr(nrep),
                             C.byref(st))
        if n > 0:
            buf = (C.c_char * (n * DOM_DTYPE.itemsize)).from_address(rows_p.value)
            rows = np.frombuffer(buf, dtype=DOM_DTYPE).copy()
        else:
            rows = np.zeros(0, dtype=DOM_DTYPE)
        lib().ora_free(rows_p)
        return rows, nrep[: self.n], st


def default_params(nthreads=0, resolve_multidomain=1):
    prm = OraParams()
    lib().ora_default_params(C.byref(prm))
    prm.nthreads = nthreads
    prm.resolve_multidomain = resolve_multidomain
    return prm


def rng_state0(seed=42):
    L = lib()
    L.ora_rng_state0.restype = C.c_uint32
    L.ora_rng_state0.argtypes = [C.c_uint32]
    return int(L.ora_rng_state0(seed))


def rng_jump(x, k):
    L = lib()
    L.ora_rng_jump.restype = C.c_uint32
    L.ora_rng_jump.argtypes = [C.c_uint32, C.c_uint64]
    return int(L.ora_rng_jump(x, k))


def score10(bits):
    return lib().ora_score10(float(np.float32(bits)))


def itspos(rows, side_of_profile, seqlen):
    """ItsPosition restatement: returns dict of int32 arrays (start/stop/tlen = -1 for None)."""
    rows = np.ascontiguousarray(rows, dtype=DOM_DTYPE)
    side = np.ascontiguousarray(side_of_profile, dtype=np.int8)
    seqlen = np.ascontiguousarray(seqlen, dtype=np.int32)
    n = len(seqlen)
    names = ["start", "stop", "tlen", "left_score10", "left_from", "left_to", "right_score10", "right_from",
             "right_to"]
    out = {k: np.zeros(n, np.int32) for k in names}
    lib().ora_itspos(_ptr(rows), len(rows), _ptr(side), _ptr(seqlen), n, *[_ptr(out[k]) for k in names])
    return out


def derep(seq_concat, off):
    """vsearch --fastx_uniques --strand both restatement.  Returns (rep_index int32, strand uint8, nclusters)."""
    seq = np.frombuffer(seq_concat, dtype=np.uint8) if isinstance(seq_concat, (bytes, bytearray)) else seq_concat
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    rep = np.zeros(n, np.int32)
    strand = np.zeros(n, np.uint8)
    nc = lib().ora_derep(_ptr(seq), _ptr(off), n, _ptr(rep), _ptr(strand))
    return rep, strand, int(nc)


def trim_bounds(off, rep_index, start, stop, tlen, mode=0, off_r2=None):
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    rep_index = np.ascontiguousarray(rep_index, dtype=np.int32)
    start, stop, tlen = (np.ascontiguousarray(a, dtype=np.int32) for a in (start, stop, tlen))
    keep = np.zeros(n, np.uint8)
    lo = np.zeros(n, np.int32)
    hi = np.zeros(n, np.int32)
    o2 = np.ascontiguousarray(off_r2, dtype=np.int64) if off_r2 is not None else off
    lib().ora_trim_bounds(_ptr(off), n, _ptr(rep_index), _ptr(start), _ptr(stop), _ptr(tlen), mode, _ptr(o2),
                          _ptr(keep), _ptr(lo), _ptr(hi))
    return keep, lo, hi


def merge_params(allow_stagger=False, **kw):
    p = OraMergeParams()
    lib().ora_merge_default_params(C.byref(p))
    p.allow_stagger = int(bool(allow_stagger))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def merge_pairs(fseq, fqual, foff, rseq, rqual, roff, params=None, nthreads=0):
    """`vsearch --fastq_mergepairs` restatement (ora_merge.c).  Returns (merged_len int32[n], reason uint8[n],
    slot_seq, slot_qual): pair i's merged read is slot_*[foff[i] + roff[i] :][: merged_len[i]].  Raises ValueError
    where vsearch stops with a fatal error (quality value outside [0, qmax])."""
    a = [np.ascontiguousarray(x, dtype=np.uint8) for x in (fseq, fqual, rseq, rqual)]
    foff = np.ascontiguousarray(foff, dtype=np.int64)
    roff = np.ascontiguousarray(roff, dtype=np.int64)
    n = len(foff) - 1
    params = params or merge_params()
    mlen = np.zeros(n, np.int32)
    reason = np.zeros(n, np.uint8)
    cap = int(foff[-1] + roff[-1]) + 1
    oseq = np.zeros(cap, np.uint8)
    oqual = np.zeros(cap, np.uint8)
    r = lib().ora_merge_pairs(_ptr(a[0]), _ptr(a[1]), _ptr(foff), _ptr(a[2]), _ptr(a[3]), _ptr(roff), n,
                              C.byref(params), _ptr(mlen), _ptr(reason), _ptr(oseq), _ptr(oqual), nthreads)
    if r < 0:
        raise ValueError("FASTQ quality value outside [0, qmax]")
    return mlen, reason, oseq, oqual


def merge_tables(params=None):
    params = params or merge_params()
    match = np.zeros((94, 94)); mism = np.zeros((94, 94)); q2p = np.zeros(94)
    same = np.zeros((94, 94), np.uint8); diff = np.zeros((94, 94), np.uint8)
    lib().ora_merge_tables(C.byref(params), _ptr(match), _ptr(mism), _ptr(same), _ptr(diff), _ptr(q2p))
    return match, mism, same, diff, q2p
