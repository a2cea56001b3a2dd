"""Counter-based synthetic amplicon generator for the BASELINE configurations that do not fit one process's Python
loops (configs[2]: 10 M reads, configs[3]: 100 M reads): any rank can produce any block [lo, hi) of the SAME global
sample without generating the rest, on the GPU (torch) or on the CPU.

Same template as synth.py / SURVEY.md 8(d):
    [5' flank 5-25][45-mer sampled from a LEFT-boundary profile][random spacer][45-mer sampled from a RIGHT-boundary
    profile][3' flank 5-25], total length drawn per unique; 0.5 % of the uniques carry one N.
Everything is a pure function of (seed, unique id, position): a 32-bit integer mix evaluated on int64 tensors.
Which unique a read is:
    exact_twice (configs[3]): slot p = affine permutation of the read index; p < U is unique p, p >= U is a second
                  copy of unique p - U  (80 M singletons + 10 M uniques seen exactly twice at full size);
    zipf        (configs[2]): unique u has 1 + floor(E w_u) copies (w ~ 1/rank, E = N - U; the remainder goes to the
                  most abundant uniques); slot p -> u through the cumulative counts.
"""
import math
import os

import numpy as np

import synth

BIG_CONFIGS = {
    # BASELINE.json configs[3]: 100 M reads, 90 % unique (hmmsearch-bound), ITS2
    "c4": dict(n_reads=100_000_000, n_unique=90_000_000, length=(330, 441), length_law="uniform", law="exact_twice",
               hmm_file="M.hmm", left_prefix="3_", right_prefix="4_", region="ITS2",
               taxa="Metazoa (stand-in for Fungi: F.hmm missing)", seed=4 * 1_000_003),
    # BASELINE.json configs[2]: 10 M merged reads ~450 bp, --region ALL --taxa All; 40 % unique
    "c3": dict(n_reads=10_000_000, n_unique=4_000_000, length=(380, 520), length_law="normal450", law="zipf",
               hmm_file="M.hmm", left_prefix="1_", right_prefix="4_", region="ALL", taxa="All", search_files="ALL",
               seed=3 * 1_000_003),
}


def config(name, scale=1.0):
    cfg = dict(BIG_CONFIGS[name])
    if scale != 1.0:
        cfg["n_reads"] = max(1000, int(cfg["n_reads"] * scale))
        cfg["n_unique"] = max(300, int(cfg["n_unique"] * scale))
    if cfg.get("search_files") == "ALL":
        cfg["search_files"] = sorted(f for f in os.listdir(synth.HMM_DIR)
                                     if f.endswith(".hmm") and os.path.getsize(os.path.join(synth.HMM_DIR, f)))
    else:
        cfg["search_files"] = [cfg["hmm_file"]]
    return cfg


def _mix(torch, x):
    """32-bit avalanche (murmur3 finaliser) on int64 tensors holding values < 2^32."""
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & m
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & m
    x = x ^ (x >> 16)
    return x


def _h(torch, u, salt):
    return _mix(torch, (u * 0x9E3779B1 + salt * 0x7F4A7C15 + 0x165667B1) & 0xFFFFFFFF)


def _h2(torch, u, pos, salt):
    return _mix(torch, (_h(torch, u, salt) + pos * 0x27D4EB2F) & 0xFFFFFFFF)


def _coprime_multiplier(n):
    a = int(n * 0.6180339887) | 1
    while math.gcd(a, n) != 1:
        a += 2
    return a


class BigSample:
    """One global sample; block(lo, hi) materialises reads lo..hi-1 (bases, qualities, offsets)."""

    def __init__(self, name, scale=1.0, device="cpu"):
        import torch
        self.torch, self.device = torch, torch.device(device)
        self.cfg = cfg = config(name, scale)
        self.N, self.U, self.seed = cfg["n_reads"], cfg["n_unique"], cfg["seed"]
        path = os.path.join(synth.HMM_DIR, cfg["hmm_file"])
        tabs = []
        for pre in (cfg["left_prefix"], cfg["right_prefix"]):
            profs = [p for _, p in synth.read_match_emissions(path, [pre]) if p.shape[0] == 45]
            cdf = np.cumsum(np.stack(profs), axis=2)                     # [nprof, 45, 4]
            cdf[:, :, 3] = 1.0
            tabs.append(torch.from_numpy((cdf * 65536.0).astype(np.int64)).to(self.device))
        self.lcdf, self.rcdf = tabs
        self.mult = _coprime_multiplier(self.N)
        self.add = (self.seed * 2654435761) % self.N
        self.cum = None
        if cfg["law"] == "zipf":
            w = 1.0 / np.arange(1, self.U + 1)
            w /= w.sum()
            E = self.N - self.U
            cnt = 1 + np.floor(E * w).astype(np.int64)
            cnt[:int(self.N - cnt.sum())] += 1                             # the remainder: one more for the top ranks
            assert cnt.sum() == self.N
            self.cum = torch.from_numpy(np.cumsum(cnt)).to(self.device)    # slot p belongs to the first u with cum > p

    def unique_of(self, idx):
        t = self.torch
        p = (idx * self.mult + self.add) % self.N
        if self.cum is not None:
            u = t.searchsorted(self.cum, p, right=True)
            # abundant uniques would otherwise be runs of neighbouring slots: scatter the ids
            return u
        return t.where(p < self.U, p, p - self.U)

    def lengths(self, u):
        t = self.torch
        lo, hi = self.cfg["length"]
        h = _h(t, u, 11)
        if self.cfg["length_law"] == "uniform":
            return lo + h % (hi - lo + 1)
        # ~N(450, 30) clipped: mean of four uniforms on [-52, 52] has sd 30
        s = (h & 0xFF) + ((h >> 8) & 0xFF) + ((h >> 16) & 0xFF) + ((h >> 24) & 0xFF)      # 0..1020, sd 147.8
        return t.clamp(450 + ((s - 510) * 30 * 1000 // 147800), lo, hi)

    def block(self, lo, hi, chunk=200_000, want_qual=True, pinned=False):
        """Returns (seq uint8[total], qual uint8[total] or None, off int64[n+1]) as torch tensors on self.device."""
        t = self.torch
        seqs, quals, lens_all = [], [], []
        for c0 in range(lo, hi, chunk):
            c1 = min(hi, c0 + chunk)
            idx = t.arange(c0, c1, dtype=t.int64, device=self.device)
            u = self.unique_of(idx)
            L = self.lengths(u)
            Lmax = int(L.max().item())
            pos = t.arange(Lmax, dtype=t.int64, device=self.device)[None, :]
            uu = u[:, None]
            f5 = (5 + _h(t, u, 21) % 21)[:, None]
            f3 = (5 + _h(t, u, 22) % 21)[:, None]
            Lc = L[:, None]
            lp = (_h(t, u, 31) % self.lcdf.shape[0])[:, None]
            rp = (_h(t, u, 32) % self.rcdf.shape[0])[:, None]
            r = _h2(t, uu, pos, 41)
            base = r & 3                                                    # flanks and spacer: uniform ACGT
            u16 = (r >> 8) & 0xFFFF
            kl = pos - f5                                                   # node of the left motif at this column
            inl = (kl >= 0) & (kl < 45)
            cl = self.lcdf[lp.expand(-1, Lmax), kl.clamp(0, 44)]            # [n, Lmax, 4]
            bl = (u16[:, :, None] >= cl).sum(2).clamp(max=3)
            kr = pos - (Lc - f3 - 45)
            inr = (kr >= 0) & (kr < 45)
            cr = self.rcdf[rp.expand(-1, Lmax), kr.clamp(0, 44)]
            br = (u16[:, :, None] >= cr).sum(2).clamp(max=3)
            base = t.where(inl, bl, t.where(inr, br, base))
            letters = t.tensor([65, 67, 71, 84], dtype=t.uint8, device=self.device)[base]
            hn = _h(t, u, 51)
            isn = ((hn % 1000) < 5)[:, None] & (pos == ((hn >> 10) % L)[:, None])
            letters = t.where(isn, t.full_like(letters, 78), letters)
            keep = pos < Lc
            seqs.append(letters[keep])
            if want_qual:
                rq = _h2(t, idx[:, None], pos, 61)
                q = 38 - (20 * pos * pos) // (Lc * Lc) + (rq & 7) - 4
                quals.append((q.clamp(2, 41) + 33).to(t.uint8)[keep])
            lens_all.append(L)
        lens = t.cat(lens_all) if lens_all else t.zeros(0, dtype=t.int64, device=self.device)
        off = t.zeros(len(lens) + 1, dtype=t.int64, device=self.device)
        if len(lens):
            t.cumsum(lens, 0, out=off[1:])
        seq = t.cat(seqs) if seqs else t.zeros(0, dtype=t.uint8, device=self.device)
        qual = (t.cat(quals) if quals else t.zeros(0, dtype=t.uint8, device=self.device)) if want_qual else None
        return seq, qual, off
