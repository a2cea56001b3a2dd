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
