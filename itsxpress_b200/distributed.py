"""One sample sharded over the G GPUs of a box (one process per GPU, torch.distributed for the plumbing).

The reference runs one vsearch and one hmmsearch process per sample (itsxpress/SeqSample.py:106-116,
191-209); both are global operations (first occurrence of a class; hmmsearch's per-profile domZ), so a
sharded run needs exactly three exchanges:

  1. dereplication, hash-partitioned:  every rank dereplicates its block of reads exactly (itsx_derep);
     each LOCAL unique goes to the owner rank  key64 % G  with its global read index and its bases through
     one all-to-all; the owner dereplicates what it received in global-index order (itsx_derep again, so
     every class is verified base by base) and returns the class representative's global index.
  2. profile search: the owner searches the classes it owns (profiles replicated); the per-profile count of
     reported hits -- hmmsearch's domZ -- is all-reduced between stage 1 (scores, -T) and stage 2 (domE,
     ItsPosition arg-max).
  3. the per-class (representative index, start, stop, tlen) table is all-gathered; every rank then trims
     its own block (itsx_trim_set_map + itsx_positions_set + itsx_trim_bounds).  Rank order = input order.

The compute engine is duck-typed: `GpuEngine` (libitsx_b200, the product) below; the CPU tests drive the same
orchestration over gloo with an oracle-backed engine that lives in tests/.
"""
import numpy as np


# ---- communication ---------------------------------------------------------------------------------------
class Comm:
    """Variable-size exchanges over torch.distributed (nccl: tensors staged on the rank's GPU; gloo: CPU)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        backend = dist.get_backend(group) if self.on else "none"
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def all_to_all(self, arr, send_counts):
        """arr: 1-D numpy array laid out by destination rank; send_counts[G].  Returns (recv, recv_counts)."""
        send_counts = np.asarray(send_counts, np.int64)
        if self.world == 1:
            return arr.copy(), send_counts.copy()
        rc = self.torch.empty(self.world, dtype=self.torch.int64, device=self.device)
        self.dist.all_to_all_single(rc, self._t(send_counts), group=self.group)
        recv_counts = rc.cpu().numpy()
        view = np.ascontiguousarray(arr)
        k = view.dtype.itemsize                                   # exchanged as raw bytes
        t_in = self._t(view.view(np.uint8).reshape(-1))
        out = self.torch.empty(int(recv_counts.sum()) * k, dtype=self.torch.uint8, device=self.device)
        self.dist.all_to_all_single(out, t_in, [int(c) * k for c in recv_counts], [int(c) * k for c in send_counts],
                                    group=self.group)
        return out.cpu().numpy().view(view.dtype), recv_counts

    def all_reduce_sum(self, arr):
        if self.world == 1:
            return arr.copy()
        t = self._t(arr)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def all_gather_v(self, arr):
        """Concatenation of every rank's 1-D array (rank order) and the per-rank counts."""
        if self.world == 1:
            return arr.copy(), np.array([len(arr)], np.int64)
        n = self.torch.tensor([len(arr)], dtype=self.torch.int64, device=self.device)
        ns = [self.torch.empty_like(n) for _ in range(self.world)]
        self.dist.all_gather(ns, n, group=self.group)
        counts = np.array([int(x.item()) for x in ns], np.int64)
        m = int(counts.max()) if len(counts) else 0
        pad = np.zeros(m, arr.dtype)
        pad[:len(arr)] = arr
        t = self._t(pad)
        outs = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        parts = [o.cpu().numpy()[:c] for o, c in zip(outs, counts)]
        return np.concatenate(parts) if parts else arr.copy(), counts


# ---- the product engine --------------------------------------------------------------------------------------
class GpuEngine:
    """The five operations the sharded driver needs, on libitsx_b200 (no CPU path)."""

    def __init__(self, ctx, params=None):
        self.ctx, self.params = ctx, params

    def derep(self, seq, off):
        rep, strand, nu = self.ctx.derep(seq, off)
        first, _ = self.ctx.derep_clusters(nu) if nu else (np.zeros(0, np.int32), None)
        keys = self.ctx.derep_unique_keys(nu)
        return rep, strand, first, keys

    def search_stage1(self, seq, off):
        self.ctx.search_seqs_stage1(seq, off, self.params)
        return self.ctx.nreported().astype(np.int64)

    def search_stage2(self, nreported_global, nseq):
        self.ctx.nreported_set(nreported_global.astype(np.int32))
        self.ctx.search_stage2()
        p = self.ctx.positions(nseq)
        return p["start"], p["stop"], p["tlen"]

    def trim_bounds(self, uid, n_unique, start, stop, tlen, off, mode=0):
        self.ctx.trim_set_map(uid, n_unique)
        self.ctx.positions_set(start, stop, tlen)
        keep, lo, hi, nk = self.ctx.trim_bounds(len(uid), mode=mode, off_other=off)
        return keep, lo, hi


# ---- helpers ---------------------------------------------------------------------------------------------------
def block_range(n, rank, world):
    """[lo, hi) of rank's block when n items are dealt in contiguous, near-equal blocks."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _gather_segments(seq, off, idx):
    """Bases of reads ``idx`` packed back to back -> (bytes, lengths); multi-threaded memcpy in libitsx_b200
    (itsx_bytes_gather), host side only."""
    from .fastq import _gather
    idx = np.asarray(idx, dtype=np.int64)
    lens = (off[1:] - off[:-1])[idx]
    if len(idx) == 0:
        return np.zeros(0, np.uint8), lens
    out, _ = _gather(seq, off[idx], lens)
    return out, lens


# ---- the sharded hot path --------------------------------------------------------------------------------------
def run_sharded(engine, comm, seq, off, first_global_index):
    """Hot path for this rank's block of one sample.

    seq/off: this rank's reads (ASCII back to back, int64 offsets); first_global_index: global index of its
    first read.  Returns dict(rep=global representative index per local read (int64), strand, keep, lo, hi,
    n_unique_global, n_owned).  Results are identical to the single-GPU path on the concatenated input.
    """
    G, me = comm.world, comm.rank
    off = np.ascontiguousarray(off, np.int64)
    seq = np.ascontiguousarray(seq, np.uint8)
    n = len(off) - 1

    # 1a. local exact derep
    rep_l, strand_l, first_l, keys = engine.derep(seq, off)
    nu_l = len(first_l)
    uid_l = np.searchsorted(first_l, rep_l).astype(np.int64) if nu_l else np.zeros(0, np.int64)
    # 1b. local uniques -> owner = key % G, grouped by destination, ascending global index inside a group
    owner = (keys % np.uint64(G)).astype(np.int64) if nu_l else np.zeros(0, np.int64)
    order = np.argsort(owner, kind="stable")                     # first_l ascending => gidx ascending per group
    send_counts = np.bincount(owner, minlength=G).astype(np.int64)
    gidx_send = (first_l[order].astype(np.int64) + first_global_index)
    bases, lens = _gather_segments(seq, off, first_l[order])
    byte_counts = np.zeros(G, np.int64)
    np.add.at(byte_counts, owner[order], lens)
    gidx_recv, recv_counts = comm.all_to_all(gidx_send, send_counts)
    lens_recv, _ = comm.all_to_all(lens.astype(np.int64), send_counts)
    bases_recv, _ = comm.all_to_all(bases, byte_counts)
    # 1c. owner: exact derep of the received uniques in global-index order (first occurrence = smallest index)
    m = len(gidx_recv)
    o_in = np.zeros(m + 1, np.int64)
    np.cumsum(lens_recv, out=o_in[1:])
    by_g = np.argsort(gidx_recv, kind="stable")
    bases_sorted, _ = _gather_segments(bases_recv, o_in, by_g)
    o_sorted = np.zeros(m + 1, np.int64)
    np.cumsum(lens_recv[by_g], out=o_sorted[1:])
    rep_o, strand_o, first_o, _ = engine.derep(bases_sorted, o_sorted)
    g_sorted = gidx_recv[by_g]
    rep_g_sorted = g_sorted[rep_o] if m else np.zeros(0, np.int64)
    # back to arrival order, then home through the inverse all-to-all
    rep_g_arrival = np.empty(m, np.int64)
    rep_g_arrival[by_g] = rep_g_sorted
    strand_arrival = np.empty(m, np.int64)
    strand_arrival[by_g] = strand_o.astype(np.int64) if m else np.zeros(0, np.int64)
    rep_back, _ = comm.all_to_all(rep_g_arrival, recv_counts)
    strand_back, _ = comm.all_to_all(strand_arrival, recv_counts)
    rep_of_unique = np.empty(nu_l, np.int64)
    rep_of_unique[order] = rep_back
    strand_of_unique = np.empty(nu_l, np.int64)
    strand_of_unique[order] = strand_back
    rep_global = rep_of_unique[uid_l] if n else np.zeros(0, np.int64)
    strand = (strand_l.astype(np.int64) ^ strand_of_unique[uid_l]).astype(np.uint8) if n else np.zeros(0, np.uint8)

    # 2. the owner searches the classes it owns; domZ is global
    own_bases, own_lens = _gather_segments(bases_sorted, o_sorted, first_o)
    own_off = np.zeros(len(first_o) + 1, np.int64)
    np.cumsum(own_lens, out=own_off[1:])
    own_gidx = g_sorted[first_o] if len(first_o) else np.zeros(0, np.int64)
    nrep_local = engine.search_stage1(own_bases, own_off)
    nrep_global = comm.all_reduce_sum(nrep_local)
    start_o, stop_o, tlen_o = engine.search_stage2(nrep_global, len(first_o))

    # 3. all-gather the class table, trim the local block
    table = np.stack([own_gidx, start_o.astype(np.int64), stop_o.astype(np.int64), tlen_o.astype(np.int64)],
                     axis=1).reshape(-1) if len(first_o) else np.zeros(0, np.int64)
    flat, _ = comm.all_gather_v(table)
    tab = flat.reshape(-1, 4)
    tab = tab[np.argsort(tab[:, 0], kind="stable")]
    uid_g = np.searchsorted(tab[:, 0], rep_global).astype(np.int32) if n else np.zeros(0, np.int32)
    keep, lo, hi = engine.trim_bounds(uid_g, len(tab), tab[:, 1].astype(np.int32), tab[:, 2].astype(np.int32),
                                      tab[:, 3].astype(np.int32), off)
    return dict(rep=rep_global, strand=strand, keep=keep, lo=lo, hi=hi, n_unique_global=len(tab),
                n_owned=len(first_o), nreported=nrep_global)


# ---- the sharded hot path, device resident -----------------------------------------------------------------------
def _umod(keys_i64, G):
    """(uint64 key) % G on an int64 tensor that holds the key's bit pattern (torch has no uint64 arithmetic)."""
    hi = (keys_i64 >> 32) & 0xFFFFFFFF
    lo = keys_i64 & 0xFFFFFFFF
    return ((hi % G) * ((1 << 32) % G) + lo % G) % G


def _gather_segments_dev(torch, src, starts, lens):
    """src[starts[i] : starts[i] + lens[i]] packed back to back, on the device."""
    total = int(lens.sum().item()) if lens.numel() else 0
    out_off = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=src.device)
    if lens.numel():
        torch.cumsum(lens, 0, out=out_off[1:])
    if total == 0:
        return torch.zeros(0, dtype=src.dtype, device=src.device), out_off
    delta = torch.repeat_interleave(starts - out_off[:-1], lens)
    idx = delta + torch.arange(total, dtype=torch.int64, device=src.device)
    return src[idx], out_off


class DeviceComm:
    """Variable-size exchanges on device tensors (NCCL over NVLink); only the split sizes visit the host."""

    def __init__(self, torch, dist, group, device):
        self.torch, self.dist, self.group, self.device = torch, dist, group, device
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def counts(self, send_counts):
        """send_counts: int64 device tensor [G] -> recv_counts (python lists of both)."""
        s = [int(x) for x in send_counts.tolist()]
        if self.world == 1:
            return s, list(s)
        rc = self.torch.empty_like(send_counts)
        self.dist.all_to_all_single(rc, send_counts.contiguous(), group=self.group)
        return s, [int(x) for x in rc.tolist()]

    def all_to_all(self, t, send, recv):
        if self.world == 1:
            return t.clone()
        out = self.torch.empty(sum(recv), dtype=t.dtype, device=self.device)
        self.dist.all_to_all_single(out, t.contiguous(), recv, send, group=self.group)
        return out

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, t):
        """t: [n, k] int64 -> concatenation over ranks (rank order)."""
        if self.world == 1:
            return t
        torch = self.torch
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=self.device)
        ns = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(ns, n, group=self.group)
        counts = [int(x) for x in ns.tolist()]
        m = max(counts) if counts else 0
        pad = torch.zeros((m, t.shape[1]), dtype=t.dtype, device=self.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * m, t.shape[1]), dtype=t.dtype, device=self.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        return torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)])


def run_sharded_device(ctx, seq, off, first_global_index, params=None, group=None):
    """Same contract and results as run_sharded(GpuEngine(ctx), ...), but every intermediate stays in HBM: the C ABI is
    called with device pointers (torch tensors), the regrouping is torch index arithmetic on the GPU, and the three
    exchanges are NCCL collectives on device buffers.  Host traffic: the rank's block of reads in, keep/lo/hi/rep
    out, and the split sizes of the collectives."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._lib import lib, SearchParams  # noqa: F401
    L, h = lib(), ctx._h
    dev = torch.device("cuda", int(ctx.device))
    comm = DeviceComm(torch, dist, group, dev)
    G = comm.world
    i64, i32, u8 = torch.int64, torch.int32, torch.uint8
    off = np.ascontiguousarray(off, np.int64)
    seq = np.ascontiguousarray(seq, np.uint8)
    n = len(off) - 1
    prm = C.byref(params) if params is not None else None

    def P(t):
        return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None

    def fence():
        """torch's stream -> the library's stream: the library copies from these tensors on its own stream (and
        synchronises it before it returns, which orders the other direction)."""
        torch.cuda.current_stream(dev).synchronize()

    def derep(seq_d, off_d, cnt):
        rep = torch.empty(cnt, dtype=i32, device=dev)
        strand = torch.empty(cnt, dtype=u8, device=dev)
        nu = C.c_int64()
        if cnt == 0:
            off_d = torch.zeros(1, dtype=i64, device=dev)
        fence()
        ctx._chk(L.itsx_derep(h, P(seq_d), C.c_void_p(off_d.data_ptr()), cnt, P(rep), P(strand), C.byref(nu)))
        first = torch.empty(int(nu.value), dtype=i32, device=dev)
        if nu.value:
            ctx._chk(L.itsx_derep_clusters(h, P(first), None))
        return rep, strand, first

    # 1a. local exact derep
    seq_d = torch.from_numpy(seq).to(dev)
    off_d = torch.from_numpy(off).to(dev)
    rep_l, strand_l, first_l = derep(seq_d, off_d, n)
    nu_l = first_l.numel()
    keys = torch.empty(nu_l, dtype=i64, device=dev)
    if nu_l:
        ctx._chk(L.itsx_derep_unique_keys(h, P(keys)))
    first_l64 = first_l.to(i64)
    uid_l = torch.searchsorted(first_l64, rep_l.to(i64)) if nu_l else torch.zeros(0, dtype=i64, device=dev)
    # 1b. local uniques -> owner = key % G, grouped by destination, ascending global index inside a group
    owner = _umod(keys, G) if nu_l else torch.zeros(0, dtype=i64, device=dev)
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=G).to(i64)
    first_o = first_l64[order]
    gidx_send = first_o + int(first_global_index)
    lens_all = off_d[1:] - off_d[:-1]
    lens = lens_all[first_o]
    bases, _ = _gather_segments_dev(torch, seq_d, off_d[first_o], lens)
    byte_counts = torch.zeros(G, dtype=i64, device=dev)
    if nu_l:
        byte_counts.index_add_(0, owner[order], lens)
    sc, rc = comm.counts(send_counts)
    bsc, brc = comm.counts(byte_counts)
    gidx_recv = comm.all_to_all(gidx_send, sc, rc)
    lens_recv = comm.all_to_all(lens, sc, rc)
    bases_recv = comm.all_to_all(bases, bsc, brc)
    # 1c. owner: exact derep of the received uniques in global-index order (first occurrence = smallest index)
    m = gidx_recv.numel()
    o_in = torch.zeros(m + 1, dtype=i64, device=dev)
    if m:
        torch.cumsum(lens_recv, 0, out=o_in[1:])
    by_g = torch.argsort(gidx_recv, stable=True)
    bases_sorted, o_sorted = _gather_segments_dev(torch, bases_recv, o_in[:-1][by_g], lens_recv[by_g])
    rep_o, strand_o, first_own = derep(bases_sorted, o_sorted, m)     # the owned classes are now ctx's search set
    g_sorted = gidx_recv[by_g]
    rep_g_arrival = torch.empty(m, dtype=i64, device=dev)
    strand_arrival = torch.empty(m, dtype=i64, device=dev)
    if m:
        rep_g_arrival[by_g] = g_sorted[rep_o.to(i64)]
        strand_arrival[by_g] = strand_o.to(i64)
    rep_back = comm.all_to_all(rep_g_arrival, rc, sc)
    strand_back = comm.all_to_all(strand_arrival, rc, sc)
    rep_of_unique = torch.empty(nu_l, dtype=i64, device=dev)
    strand_of_unique = torch.empty(nu_l, dtype=i64, device=dev)
    if nu_l:
        rep_of_unique[order] = rep_back
        strand_of_unique[order] = strand_back
    rep_global = rep_of_unique[uid_l] if n else torch.zeros(0, dtype=i64, device=dev)
    strand = (strand_l.to(i64) ^ strand_of_unique[uid_l]).to(u8) if n else torch.zeros(0, dtype=u8, device=dev)

    # 2. the owner searches the classes it owns (resident since the second derep); domZ is global
    n_own = first_own.numel()
    ctx._chk(L.itsx_search_stage1(h, prm))
    nrep = torch.from_numpy(ctx.nreported().astype(np.int64)).to(dev)
    nrep = comm.all_reduce_sum(nrep)
    nrep_h = nrep.cpu().numpy()
    ctx.nreported_set(nrep_h.astype(np.int32))
    ctx._chk(L.itsx_search_stage2(h))
    start = torch.full((n_own,), -1, dtype=i32, device=dev)
    stop = torch.full((n_own,), -1, dtype=i32, device=dev)
    tlen = torch.full((n_own,), -1, dtype=i32, device=dev)
    if n_own:
        ctx._chk(L.itsx_positions(h, P(start), P(stop), P(tlen), None, None, None, None, None, None))
    own_gidx = g_sorted[first_own.to(i64)] if n_own else torch.zeros(0, dtype=i64, device=dev)

    # 3. all-gather the class table, trim the local block
    table = torch.stack([own_gidx, start.to(i64), stop.to(i64), tlen.to(i64)], dim=1) if n_own else \
        torch.zeros((0, 4), dtype=i64, device=dev)
    tab = comm.all_gather_rows(table)
    tab = tab[torch.argsort(tab[:, 0], stable=True)]
    nt = tab.shape[0]
    uid_g = torch.searchsorted(tab[:, 0].contiguous(), rep_global).to(i32) if n else torch.zeros(0, dtype=i32, device=dev)
    keep = torch.empty(n, dtype=u8, device=dev)
    lo = torch.empty(n, dtype=i32, device=dev)
    hi = torch.empty(n, dtype=i32, device=dev)
    if n:
        t1, t2, t3 = (tab[:, k].to(i32).contiguous() for k in (1, 2, 3))
        fence()
        ctx._chk(L.itsx_trim_set_map(h, P(uid_g), n, nt))
        ctx._chk(L.itsx_positions_set(h, P(t1), P(t2), P(t3), nt))
        nk = C.c_int64()
        ctx._chk(L.itsx_trim_bounds(h, 0, C.c_void_p(off_d.data_ptr()), n, P(keep), P(lo), P(hi), C.byref(nk)))
    return dict(rep=rep_global.cpu().numpy(), strand=strand.cpu().numpy(), keep=keep.cpu().numpy(), lo=lo.cpu().numpy(),
                hi=hi.cpu().numpy(), n_unique_global=int(nt), n_owned=int(n_own), nreported=nrep_h)
